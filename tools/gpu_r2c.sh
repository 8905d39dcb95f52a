#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python tools/stats_probe.py 2 1 3 4 5 2>&1 | tee gpurun_out/stats_r2c.txt
for b in exact lerp32; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:remap_image -s 20 -c 1 -f -o gpurun_out/ncu_img_fast_$b python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 0 --blend $b > gpurun_out/ncu_img_fast_$b.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
