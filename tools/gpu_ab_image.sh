#!/bin/bash
# A/B of library builds on the single-image bench: tools/gpu_ab_image.sh "libA libB ..." "blends"
set -u
cd "$(dirname "$0")/.."
libs=${1:-default}; blends=${2:-"exact lerp32"}
for round in 1 2; do
for lib in $libs; do
  if [ "$lib" = default ]; then unset DCB_LIB; else export DCB_LIB=$PWD/$lib; fi
  for blend in $blends; do
    timeout 300 python bench.py --steps 20 --warmup 3 --blend $blend --no-cpu-baseline --no-extras --e2e-steps 0 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('%-14s %-7s kernel %.2f us  frac %.3f  clocks %s' % ('$lib'.split('/')[-1], d['config']['blend'], d['roofline']['kernel_us'], d['roofline']['frac'], d['clocks'].get('sm_mhz')))
"
  done
done; done
