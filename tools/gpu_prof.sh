#!/bin/bash
# ncu --set full capture of the dominant kernel + launch list; args: tag [bench args...]
set -u
cd "$(dirname "$0")/.."
tag=${1:-prof}; shift || true
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:remap_ -s 20 -c 1 -f -o gpurun_out/$tag python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 1 "$@" > gpurun_out/${tag}_run.log 2>&1
tail -2 gpurun_out/${tag}_run.log
ls -la gpurun_out/$tag.ncu-rep
