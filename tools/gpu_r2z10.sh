#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-r2z15}
echo "== pytest gpu"; timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
echo "== pageable dtypes"; timeout 300 python tools/e2e_pageable.py 2>&1 | tee gpurun_out/e2e_pageable_$tag.txt
