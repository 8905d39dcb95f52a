#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export DCB_LIB=$PWD/discorpy_b200/lib/ab/libdcb_timeline.so
timeout 200 python tools/timeline_probe.py 1 exact 2>&1 | grep -v "slowest\|fastest" | tee gpurun_out/timeline_${1:-r2y}.txt
