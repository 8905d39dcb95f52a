#!/bin/bash
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-r2m}
echo "== pytest gpu"; timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
echo "== bench default"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default_$tag.json 2> gpurun_out/bench_default_$tag.err
python - <<PY
import json
d=json.loads(open("gpurun_out/bench_default_$tag.json").read().strip().splitlines()[-1])
print("kernel_us", d["roofline"]["kernel_us"], "frac", d["roofline"]["frac"])
print("e2e", d["e2e"]["value"], d["e2e_pageable"], d.get("pcie"))
print({k:(v.get("roofline_frac_per_gpu") or v.get("roofline_frac_vs_8B_per_px"), v.get("parity_ok")) for k,v in d["extras"].items() if k.startswith("cfg")})
PY
tail -3 gpurun_out/bench_default_$tag.err
echo "== sanitizers (small cases)"
cat > /tmp/san_case.py <<PY
import sys, numpy as np
sys.path.insert(0, ".")
import discorpy_b200 as dcb
import discorpy_b200.post.postprocessing as post
dcb.set_device(0)
rng = np.random.default_rng(0)
mat = rng.random((300, 400), dtype=np.float32)
fact = [1.0, -2e-5, 6e-8, -1e-10, 5e-14]
for order in (1, 0):
    post.unwarp_image_backward(dcb.DeviceArray.from_host(mat), 201.3, 149.2, fact, order=order).to_host()
post.config["blend"] = dcb.BLEND_LERP32
post.unwarp_image_backward(dcb.DeviceArray.from_host(mat), 201.3, 149.2, fact).to_host()
post.config["blend"] = dcb.BLEND_EXACT
stack = rng.random((5, 96, 400), dtype=np.float32)
post.unwarp_chunk_slices_backward(stack, 201.3, 47.2, fact, 0, 95)
post.unwarp_slice_backward(stack, 201.3, 47.2, fact, 40)
post.correct_perspective_image(mat, [1.02, 0.01, -5.0, 0.005, 1.01, -3.0, 8e-6, -5e-6])
print("sanitizer case done")
PY
for tool in memcheck racecheck initcheck; do
  echo "-- $tool"; timeout 900 compute-sanitizer --tool $tool --print-limit 20 python /tmp/san_case.py > gpurun_out/sanitizer_${tool}_$tag.txt 2>&1; tail -4 gpurun_out/sanitizer_${tool}_$tag.txt
done
