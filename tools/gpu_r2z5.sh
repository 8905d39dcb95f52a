#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-r2z11}
echo "== pytest gpu"; timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for nt in 1 0; do
for n in 4 8 16; do
DCB_COPY_NT=$nt DCB_COPY_THREADS=$n timeout 100 python tools/e2e_copy_threads.py $n 2>&1 | sed "s/^/NT=$nt /" | tee -a gpurun_out/e2e_copy_nt_$tag.txt
done; done
