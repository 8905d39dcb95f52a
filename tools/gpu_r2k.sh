#!/bin/bash
# N-GPU check: the driver's launch line, NCCL through the C ABI, sharded extras, exchange step
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
n=${1:-2}; tag=${2:-r2k}
nvidia-smi -L | head -8
echo "== bench N=$n (torchrun launch line)"
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $n --steps 20 --warmup 5 ) > gpurun_out/bench_${n}gpu_$tag.json 2> gpurun_out/bench_${n}gpu_$tag.err
tail -c 5000 gpurun_out/bench_${n}gpu_$tag.json; tail -8 gpurun_out/bench_${n}gpu_$tag.err
echo "== sinogram exchange check"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29611 tests/two_rank_sinogram_check.py --time 2>&1 | tail -3 | tee gpurun_out/sinogram_exchange_${n}gpu_$tag.txt
echo "== pytest gpu (incl. the two-rank test)"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
