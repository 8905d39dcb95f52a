#!/bin/bash
# A/B of library builds: tools/gpu_ab2.sh tag "lib1 lib2" [pytest]   (lib = default | name under lib/ab)
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-ab2}; libs=${2:-default}; dotest=${3:-}
one() {
  timeout 300 python bench.py --steps 20 --warmup 3 --blend $1 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
x = d.get('extras', {}).get('single_image_kernel_us', {})
print('%-12s %-7s kernel %.2f us  frac %.3f  clocks %s extras %s' % ('$2', d['config']['blend'], d['roofline']['kernel_us'], d['roofline']['frac'], d['clocks'].get('sm_mhz'), {k: round(v, 2) for k, v in x.items()}))
" | tee -a gpurun_out/ab2_$tag.txt
}
for lib in $libs; do
  if [ "$lib" = default ]; then unset DCB_LIB; else export DCB_LIB=$PWD/discorpy_b200/lib/ab/libdcb_$lib.so; fi
  if [ -n "$dotest" ]; then echo "== pytest gpu ($lib)"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q -m gpu 2>&1 | tail -4; fi
done
for r in 1 2; do for lib in $libs; do
  if [ "$lib" = default ]; then unset DCB_LIB; else export DCB_LIB=$PWD/discorpy_b200/lib/ab/libdcb_$lib.so; fi
  one exact $lib
done; done
