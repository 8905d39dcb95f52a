#!/bin/bash
# full check: GPU tests, the default bench line, smoke
set -u
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
tag=${1:-full}
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
echo "== pytest gpu"; timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -4 | tee gpurun_out/pytest_gpu_$tag.txt
echo "== bench"; timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default_$tag.json 2> gpurun_out/bench_default_$tag.err; tail -c 600 gpurun_out/bench_default_$tag.err
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_default_$tag.json").read().strip().splitlines()[-1])
print("value", d["value"], "kernel_us", d["roofline"]["kernel_us"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "pageable", d.get("e2e_pageable", {}).get("value"))
print(d["extras"]["single_image_kernel_us"], d["extras"]["roofline_frac"])
for k in ("cfg3", "cfg4", "cfg5"):
    print(k, {a: b for a, b in d["extras"][k].items() if a in ("ms", "kernel_us_per_image", "roofline_frac_per_gpu", "roofline_frac_vs_8B_per_px", "parity_ok")})
print("cold call us", d["extras"].get("single_image_cold_call_us"), "cpu", d["cpu_baseline"]["value"])
PY
