#!/bin/bash
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
echo "== tma probe"; timeout 300 python tools/gpu_probe_tma.py 2>&1 | tee gpurun_out/tma_probe.txt | tail -60
echo "== sanitizer on a failing case"
timeout 600 compute-sanitizer --tool memcheck python - <<'PY' 2>&1 | tail -40 | tee gpurun_out/sanitizer.txt
import numpy as np, discorpy_b200 as dcb, discorpy_b200.post.postprocessing as post
m = np.random.default_rng(0).random((37,53), dtype=np.float32)
try:
    post.unwarp_image_backward(m, 53/3.0, 37/1.7, [0.8], order=0)
    print("ran fine")
except Exception as e:
    print("EXC", e)
PY
