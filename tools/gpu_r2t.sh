#!/bin/bash
# per-CTA timeline of the single-image kernel (three variants), -DDCB_IMG_TIMELINE build
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-r2t}
export DCB_LIB=$PWD/discorpy_b200/lib/ab/libdcb_${2:-timeline}.so
for v in "1 exact" "1 lerp32" "0 exact"; do
  timeout 200 python tools/timeline_probe.py $v 2>&1 | tail -40 | tee -a gpurun_out/timeline_$tag.txt
done
