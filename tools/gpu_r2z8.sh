#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-r2z13}
echo "== stream tests"; timeout 900 python -m pytest tests/test_gpu_stream.py -x -q -m gpu 2>&1 | tail -3
echo "== stream bench (NT pool)"; timeout 300 python tools/bench_stream.py 2>&1 | tee gpurun_out/bench_stream_$tag.jsonl
echo "== stream bench (DCB_COPY_NT=0)"; DCB_COPY_NT=0 timeout 300 python tools/bench_stream.py 2>&1 | grep pageable | tee gpurun_out/bench_stream_nt0_$tag.jsonl
