#!/bin/bash
# ncu --set full captures of the shipped kernels, summarised on the box (the .ncu-rep files are
# too large to travel together), + launch list of the default bench command
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-r2ac}
sumup() {  # rep npx out
  { python tools/ncu_summary.py $1 $2; echo; echo "== hot instructions (stall samples)"; python tools/ncu_hot.py $1 40 $2; } > $3 2>&1
  rm -f $1
}
for blend in exact lerp32; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:remap_image -s 20 -c 1 -f -o gpurun_out/ncu_image_${tag}_$blend python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 0 --blend $blend > gpurun_out/ncu_image_${tag}_$blend.log 2>&1
sumup gpurun_out/ncu_image_${tag}_$blend.ncu-rep 16777216 gpurun_out/ncu_image_${tag}_$blend.txt
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:remap_stack -s 2 -c 1 -f -o gpurun_out/ncu_stack_${tag}_exact python tools/bench_stack.py --cases cfg2x16 --blends exact --reps 1 > gpurun_out/ncu_stack_${tag}_exact.log 2>&1
sumup gpurun_out/ncu_stack_${tag}_exact.ncu-rep $((16*4096*4096)) gpurun_out/ncu_stack_${tag}_exact_16x4096.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:remap_stack -s 2 -c 1 -f -o gpurun_out/ncu_stack_${tag}_cfg4 python tools/bench_stack.py --cases cfg4shard --blends exact --reps 1 > gpurun_out/ncu_stack_${tag}_cfg4.log 2>&1
sumup gpurun_out/ncu_stack_${tag}_cfg4.ncu-rep $((64*2560*2560)) gpurun_out/ncu_stack_${tag}_cfg4_64x2560.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/launches_$tag.log 2>&1
ls -la gpurun_out/*$tag*
