#!/bin/bash
# Round 2, step A: parity of the patch path + A/B of DCB_IMG_FAST on the single-image bench.
set -u
cd "$(dirname "$0")/.."
tag=${1:-r2a}; what=${2:-"smoke tests stats ab"}
has() { case " $what " in *" $1 "*) return 0;; *) return 1;; esac; }
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv | tail -1
if has smoke; then echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3; fi
if has tests; then echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu_$tag.txt; fi
if has stats; then echo "== stats"; timeout 300 python tools/stats_probe.py 2 1 3 4 5 2>&1 | tee gpurun_out/stats_$tag.txt; fi
one() {
  timeout 300 python bench.py --steps 20 --warmup 3 --blend $1 --no-cpu-baseline $2 --e2e-steps 0 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
x = d.get('extras', {}).get('single_image_kernel_us')
print('FAST=%s %-7s kernel %.2f us  frac %.3f  clocks %s  extras %s' % ('$DCB_IMG_FAST', d['config']['blend'], d['roofline']['kernel_us'], d['roofline']['frac'], d['clocks'].get('sm_mhz'), x))
" | tee -a gpurun_out/ab_fast_$tag.txt
}
if has ab; then
for round in 1 2; do
for f in 1 0; do
  export DCB_IMG_FAST=$f
  one exact ""
  one lerp64 --no-extras
  one lerp32 --no-extras
done; done
unset DCB_IMG_FAST
fi
echo "== done"
