#!/bin/bash
# fractional tile ranges + programmatic dependent launch: parity, bench A/B (PDL on/off), timeline
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-r2u}
echo "== pytest gpu (parity subset)"; timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q -m gpu 2>&1 | tail -4
one() {
  timeout 300 python bench.py --steps 20 --warmup 3 --blend $1 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
x = d.get('extras', {}).get('single_image_kernel_us', {})
print('PDL=%s %-7s kernel %.2f us  frac %.3f  clocks %s extras %s' % ('${DCB_IMG_PDL:-1}', d['config']['blend'], d['roofline']['kernel_us'], d['roofline']['frac'], d['clocks'].get('sm_mhz'), {k: round(v, 2) for k, v in x.items()}))
" | tee -a gpurun_out/ab_pdl_$tag.txt
}
for r in 1 2; do
  DCB_IMG_PDL=1 one exact
  DCB_IMG_PDL=0 one exact
done
export DCB_LIB=$PWD/discorpy_b200/lib/ab/libdcb_timeline.so
for v in "1 exact" "1 lerp32"; do
  timeout 200 python tools/timeline_probe.py $v 2>&1 | grep -v "^  warp\|event log\|slowest\|fastest" | tee -a gpurun_out/timeline_$tag.txt
done
