#!/bin/bash
# N-GPU run as the driver does it + the host-side PCIe ceiling probe
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
n=${1:-8}; tag=${2:-r2s}
nvidia-smi -L | wc -l
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $1 "${@:2}"; }
echo "== pcie probe N=$n"; timeout 300 bash -c "$(declare -f run); n=$n; run 29544 tools/pcie_probe_multi.py" 2>/dev/null | tail -1 | tee gpurun_out/pcie_probe_${n}gpu_$tag.json
echo "== bench N=$n"; ( time timeout 900 bash -c "$(declare -f run); n=$n; run 29533 bench.py --gpus $n --steps 20 --warmup 5" ) > gpurun_out/bench_${n}gpu_$tag.json 2> gpurun_out/bench_${n}gpu_$tag.err
tail -1 gpurun_out/bench_${n}gpu_$tag.json | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('value', d['value'], 'e2e', d['e2e']['value'], 'pageable', d['e2e_pageable']['value'], d['e2e_pageable']['reused_buffers_value'])
for k in ('cfg4', 'cfg5'):
    print(k, {a: d['extras'][k][a] for a in ('ms', 'Mpixels_per_s', 'roofline_frac_per_gpu', 'parity_ok')})
print('exchange', d.get('exchange'))
"
tail -4 gpurun_out/bench_${n}gpu_$tag.err
