#!/bin/bash
# Z-stack kernel A/B: tools/gpu_stk.sh tag "libs" [cases] [blends]
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-stk}; libs=${2:-default}; cases=${3:-cfg2x16,cfg4shard,cfg4chunk,cfg5shard}; blends=${4:-exact,lerp64}
if [ -z "${NOTEST:-}" ]; then echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3; fi
for lib in $libs; do
  if [ "$lib" = default ]; then unset DCB_LIB; else export DCB_LIB=$PWD/discorpy_b200/lib/ab/libdcb_$lib.so; fi
  timeout 900 python tools/bench_stack.py --reps 5 --cases $cases --blends $blends 2>&1 | tee -a gpurun_out/bench_stack_$tag.jsonl | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l)
        print('%-8s %-10s %-7s cr %d  %.3f ms  frac %.3f  us/4096sq %.2f' % ('$lib', d['case'], d['blend'], d['coord_round'], d['ms'], d['frac'], d['us_per_4096sq']))
    except Exception:
        print(l[:200])
"
done
