#!/bin/bash
# compute-sanitizer (memcheck, racecheck, initcheck, synccheck) on tools/sanitizer_case.py
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-san}
for tool in memcheck racecheck initcheck synccheck; do
  echo "-- $tool"; timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitizer_case.py > gpurun_out/sanitizer_${tool}_$tag.txt 2>&1; tail -3 gpurun_out/sanitizer_${tool}_$tag.txt
done
