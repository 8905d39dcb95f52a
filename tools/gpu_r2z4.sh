#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-r2z8}
bash tools/gpu_full.sh $tag
echo "== sizes"; timeout 300 python tools/e2e_sizes.py 2>&1 | tee gpurun_out/e2e_sizes_$tag.txt
echo "== schedules"; DCB_PIPE_PROBE=1 timeout 300 python tools/e2e_edges.py 2>&1 | tee gpurun_out/e2e_edges_$tag.txt | tail -30
