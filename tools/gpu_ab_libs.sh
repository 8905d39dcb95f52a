#!/bin/bash
# A/B of library builds on the single-image bench: tools/gpu_ab_libs.sh tag "libA libB ..." "blends"
set -u
cd "$(dirname "$0")/.."
tag=${1:-ab}; libs=${2:-default}; blends=${3:-"exact lerp32"}
mkdir -p gpurun_out
for round in 1 2; do
for lib in $libs; do
  if [ "$lib" = default ]; then unset DCB_LIB; else export DCB_LIB=$PWD/discorpy_b200/lib/$lib; fi
  for blend in $blends; do
    timeout 300 python bench.py --steps 20 --warmup 3 --blend $blend --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
x = d.get('extras', {}).get('single_image_kernel_us', {})
print('%-22s %-7s kernel %.2f us  frac %.3f  order0 %.2f' % ('$lib', d['config']['blend'], d['roofline']['kernel_us'], d['roofline']['frac'], x.get('order0', 0)))
" | tee -a gpurun_out/ab_libs_$tag.txt
  done
done; done
