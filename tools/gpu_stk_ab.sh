#!/bin/bash
# Z-stack kernel A/B: default library against lib/ab builds ($2) on the BASELINE-shaped cases
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-stkab}; libs=${2:-""}
for rep in 1 2; do
for lib in default $libs; do
  if [ "$lib" = default ]; then unset DCB_LIB; else export DCB_LIB=$PWD/discorpy_b200/lib/ab/libdcb_$lib.so; fi
  timeout 300 python tools/bench_stack.py --cases cfg2x64,cfg4shard,cfg4deep,cfg5shard --blends exact,lerp32 --reps 10 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('%-8s %-10s %-6s ms %.3f frac %.3f' % ('$lib', d['case'], d['blend'], d['ms'], d['frac']))
" | tee -a gpurun_out/ab_stack_$tag.txt
done; done
