#!/bin/bash
# A/B of the Z-stack kernel: default library against lib/ab/libdcb_<name>.so on the float64-coordinate cases
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-stk4}; ab=${2:-not64w}
echo "== parity (slice path)"; timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_stream.py -x -q -m gpu -k "slice or stack or chunk or stream" 2>&1 | tail -2
for rep in 1 2; do
for lib in default $ab; do
  if [ "$lib" = default ]; then unset DCB_LIB; else export DCB_LIB=$PWD/discorpy_b200/lib/ab/libdcb_$lib.so; fi
  timeout 300 python tools/bench_stack.py --cases cfg4shard,cfg4deep --blends exact --reps 5 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()[:200]); continue
    print('%-10s %-10s %s ms %.3f frac %.3f' % ('$lib', d['case'], d['blend'], d['ms'], d['frac']))
" | tee -a gpurun_out/ab_stack_$tag.txt
done; done
