#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-r2j}
echo "== bench default"; ( time timeout 900 python bench.py --steps 20 --warmup 5 ) > gpurun_out/bench_default_$tag.json 2> gpurun_out/bench_default_$tag.err
tail -c 6000 gpurun_out/bench_default_$tag.json; tail -5 gpurun_out/bench_default_$tag.err
echo "== reference arm"; timeout 600 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | cut -c1-700
