#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-r2z12}
echo "== pytest gpu"; timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -12
echo "== persp e2e"; timeout 300 python tools/e2e_persp.py 2>&1 | tee gpurun_out/e2e_persp_$tag.txt
