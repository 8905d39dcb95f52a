#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-r2e}
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 300 python tools/stats_probe.py 2 1 3 4 5 2>&1 | tee gpurun_out/stats_$tag.txt
one() {
  timeout 300 python bench.py --steps 20 --warmup 3 --blend $1 --no-cpu-baseline $2 --e2e-steps 0 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
x = d.get('extras', {}).get('single_image_kernel_us')
print('FAST=%s %-7s kernel %.2f us  frac %.3f  clocks %s  extras %s' % ('$DCB_IMG_FAST', d['config']['blend'], d['roofline']['kernel_us'], d['roofline']['frac'], d['clocks'].get('sm_mhz'), x))
" | tee -a gpurun_out/ab_fast_$tag.txt
}
for f in 1 0; do
  export DCB_IMG_FAST=$f
  one exact ""
  one lerp32 --no-extras
done
unset DCB_IMG_FAST
for b in exact lerp32; do
timeout 900 ncu --set full --clock-control none --import-source on -k regex:remap_image -s 20 -c 1 -f -o gpurun_out/ncu_img_${tag}_$b python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 0 --blend $b > gpurun_out/ncu_img_${tag}_$b.log 2>&1
done
ls -la gpurun_out/*$tag*.ncu-rep
