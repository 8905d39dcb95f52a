#!/bin/bash
# round-2 late check: full GPU tests + default bench on the l2-prefetch build, then the band-schedule sweep
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-r2z}
bash tools/gpu_full.sh $tag
echo "== band schedules"; timeout 300 python tools/e2e_edges.py 2>&1 | tee gpurun_out/e2e_edges_$tag.txt | tail -20
echo "== band traces"; DCB_PIPE_TRACE=1 timeout 120 python tools/e2e_edges.py equal8 both both_fine > gpurun_out/e2e_trace_$tag.txt 2>&1; tail -45 gpurun_out/e2e_trace_$tag.txt
