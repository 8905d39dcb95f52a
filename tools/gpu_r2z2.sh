#!/bin/bash
# duplex PCIe ceiling in chunks + band pipeline with split copy streams
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-r2z2}
DCB_PIPE_PROBE=1 timeout 600 python tools/e2e_edges.py 0 2>&1 | tee gpurun_out/e2e_split_$tag.txt | tail -60
echo "== band traces (linear copies)"; DCB_PIPE_TRACE=1 timeout 120 python tools/e2e_edges.py equal8 head6 > gpurun_out/e2e_trace_$tag.txt 2>&1; tail -34 gpurun_out/e2e_trace_$tag.txt
