#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-r2l}
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
echo "== stack bench"; timeout 900 python tools/bench_stack.py --reps 5 2>&1 | tee gpurun_out/bench_stack_$tag.jsonl | cut -c1-230
echo "== special values"; timeout 600 python tests/special_values_check.py 2>&1 | tail -4
echo "== ncu stack"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:remap_stack -s 2 -c 1 -f -o gpurun_out/ncu_stack_$tag python tools/bench_stack.py --cases cfg2x16 --blends exact --reps 1 > gpurun_out/ncu_stack_$tag.log 2>&1
ls -la gpurun_out/ncu_stack_$tag.ncu-rep
