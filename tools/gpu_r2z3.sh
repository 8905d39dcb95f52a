#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-r2z4}
export DCB_PIPE_DIRECT=1
echo "== direct stores: parity"; timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q -m gpu 2>&1 | tail -3
for c in 0 74; do
echo "== DIRECT, DCB_PIPE_CTAS $c"; DCB_PIPE_CTAS=$c DCB_PIPE_TRACE=1 timeout 120 python tools/e2e_edges.py equal8 head6 2>&1 | tail -25
DCB_PIPE_CTAS=$c DCB_PIPE_PROBE=1 timeout 120 python tools/e2e_edges.py 0 2>&1 | tail -5
done
