#!/bin/bash
# A/B of library builds on the Z-stack bench: tools/gpu_ab.sh "libA libB ..." "cases" "blends" [reps]
# (library paths relative to the repo root; "default" = the in-tree build)
set -u
cd "$(dirname "$0")/.."
libs=${1:-default}; cases=${2:-cfg2x64}; blends=${3:-exact,lerp32}; reps=${4:-20}
mkdir -p gpurun_out
for round in 1 2; do
for lib in $libs; do
  if [ "$lib" = default ]; then unset DCB_LIB; else export DCB_LIB=$PWD/$lib; fi
  timeout 600 python tools/bench_stack.py --reps $reps --cases $cases --blends $blends 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l)
        print('%-12s %-10s %-7s %8.2f us/4096^2  frac %.3f  clocks %s' % (d['lib'], d['case'], d['blend'], d['us_per_4096sq'], d['frac'], d['clocks'].get('sm_mhz')))
    except Exception:
        print('??', l[:200])
"
done; done
