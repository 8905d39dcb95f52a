#!/bin/bash
# pooled tile scheduling: full GPU tests, bench, timeline
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-r2v}
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for r in 1 2; do
timeout 300 python bench.py --steps 20 --warmup 3 --blend exact --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
x = d.get('extras', {}).get('single_image_kernel_us', {})
print('exact kernel %.2f us  frac %.3f  clocks %s extras %s' % (d['roofline']['kernel_us'], d['roofline']['frac'], d['clocks'].get('sm_mhz'), {k: round(v, 2) for k, v in x.items()}))
print({k: (v.get('roofline_frac_vs_8B_per_px') or v.get('roofline_frac_per_gpu')) for k, v in d['extras'].items() if k.startswith('cfg')}, d['extras'].get('single_image_cold_call_us'))
" | tee -a gpurun_out/bench_$tag.txt
done
export DCB_LIB=$PWD/discorpy_b200/lib/ab/libdcb_timeline.so
for v in "1 exact" "1 lerp32" "0 exact"; do
  timeout 200 python tools/timeline_probe.py $v 2>&1 | grep -v "^  warp\|event log\|slowest\|fastest" | tee -a gpurun_out/timeline_$tag.txt
done
