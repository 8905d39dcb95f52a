#!/bin/bash
# Z-stack kernel, config 4: forced staged-box sizes (DCB_STK_BOX="w,h"): does the box traffic bound it?
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-stk8}
for box in "" "96,64" "96,56" "96,48" "88,64" "80,64" "80,48" "72,40" "112,64" "128,64"; do
  DCB_STK_BOX=$box timeout 300 python tools/bench_stack.py --cases cfg4shard --blends exact,lerp32 --reps 10 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('box %-8s %-10s %-6s ms %.3f frac %.3f plan %s' % ('$box', d['case'], d['blend'], d['ms'], d['frac'], d['plan']))
" | tee -a gpurun_out/ab_stack_$tag.txt
done
