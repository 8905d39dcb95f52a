#!/bin/bash
# Z-stack kernel: tile shapes (DCB_STK_SHAPE) on the BASELINE geometries
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-stk2}; cases=${2:-cfg2x16,cfg4shard,cfg4chunk,cfg5shard}; blends=${3:-exact,lerp32}
if [ -z "${NOTEST:-}" ]; then echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3; fi
for shape in auto 0 1; do
  if [ "$shape" = auto ]; then unset DCB_STK_SHAPE; else export DCB_STK_SHAPE=$shape; fi
  timeout 900 python tools/bench_stack.py --reps 5 --cases $cases --blends $blends 2>&1 | tee -a gpurun_out/bench_stack_$tag.jsonl | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l)
        print('shape %-4s %-10s %-7s cr %d  %.3f ms  frac %.3f  us/4096sq %.2f  box %dx%d' % ('$shape', d['case'], d['blend'], d['coord_round'], d['ms'], d['frac'], d['us_per_4096sq'], d['plan']['box_w'], d['plan']['box_h']))
    except Exception:
        print(l[:200])
"
done
