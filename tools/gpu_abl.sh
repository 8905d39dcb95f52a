#!/bin/bash
# ablation libs on the single-image bench (exact blend, device-timed): tools/gpu_abl.sh tag "libs"
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-abl}; libs=${2:-default}
for r in 1 2; do for lib in $libs; do
  if [ "$lib" = default ]; then unset DCB_LIB; else export DCB_LIB=$PWD/discorpy_b200/lib/ab/libdcb_$lib.so; fi
  timeout 300 python bench.py --steps 20 --warmup 3 --blend exact --no-cpu-baseline --no-extras --e2e-steps 0 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('%-12s kernel %.2f us  frac %.3f' % ('$lib', d['roofline']['kernel_us'], d['roofline']['frac']))
" | tee -a gpurun_out/abl_$tag.txt
done; done
