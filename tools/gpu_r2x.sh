#!/bin/bash
# pool parameters A/B (DCB_IMG_POOL="pool,halves,depth"), exact / lerp32 / order0 kernels
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-r2x}
echo "== pytest gpu"; timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fuzz.py -x -q -m gpu 2>&1 | tail -2
for cfg in "25,10,2" "25,10,3" "30,15,3" "20,10,3" "25,10,4" "15,10,2"; do
DCB_IMG_POOL=$cfg timeout 300 python bench.py --steps 20 --warmup 3 --blend exact --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
x = d.get('extras', {}).get('single_image_kernel_us', {})
print('pool $cfg exact kernel %.2f us  frac %.3f  extras %s' % (d['roofline']['kernel_us'], d['roofline']['frac'], {k: round(v, 2) for k, v in x.items()}))
" | tee -a gpurun_out/pool_$tag.txt
done
export DCB_LIB=$PWD/discorpy_b200/lib/ab/libdcb_timeline.so
for v in "1 exact" "0 exact"; do
  timeout 200 python tools/timeline_probe.py $v 2>&1 | grep -v "^  warp\|event log\|slowest\|fastest\|^CTA start\|first tile ready" | tee -a gpurun_out/timeline_$tag.txt
done
