#!/bin/bash
# A/B: exact blend on the raw float32 box (default build) vs on pre-widened float64 tiles
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-r2n}
echo "== pytest gpu (default build)"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
one() {
  timeout 300 python bench.py --steps 20 --warmup 3 --blend $1 --no-cpu-baseline --no-extras --e2e-steps 0 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('%-22s %-7s kernel %.2f us  frac %.3f  clocks %s' % ('${DCB_LIB:-default(raw)}'.split('/')[-1], d['config']['blend'], d['roofline']['kernel_us'], d['roofline']['frac'], d['clocks'].get('sm_mhz')))
" | tee -a gpurun_out/ab_exact_raw_$tag.txt
}
for round in 1 2; do
  unset DCB_LIB; one exact
  export DCB_LIB=$PWD/discorpy_b200/lib/libdcb_exact_wide.so; one exact
done
unset DCB_LIB
timeout 900 ncu --set full --clock-control none --import-source on -k regex:remap_image -s 20 -c 1 -f -o gpurun_out/ncu_img_${tag}_exact python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 0 --blend exact > gpurun_out/ncu_img_${tag}_exact.log 2>&1
ls -la gpurun_out/ncu_img_${tag}_exact.ncu-rep
