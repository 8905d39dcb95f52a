#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-stk3}
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 900 python tools/bench_stack.py --reps 5 --cases cfg2x16,cfg2x64,cfg4shard,cfg4deep,cfg4chunk,cfg5shard --blends exact,lerp32 2>&1 | tee -a gpurun_out/bench_stack_$tag.jsonl | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l)
        print('%-10s %-7s cr %d  %.3f ms  frac %.3f  us/4096sq %.2f  box %dx%d' % (d['case'], d['blend'], d['coord_round'], d['ms'], d['frac'], d['us_per_4096sq'], d['plan']['box_w'], d['plan']['box_h']))
    except Exception:
        print(l[:200])
"
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --e2e-steps 0 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
for k in ('cfg3', 'cfg4', 'cfg5'):
    print(k, {a: b for a, b in d['extras'][k].items() if a in ('ms', 'kernel_us_per_image', 'roofline_frac_per_gpu', 'roofline_frac_vs_8B_per_px', 'parity_ok')})
print(d['extras']['roofline_frac'])
"
