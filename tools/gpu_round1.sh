#!/bin/bash
# First GPU pass: smoke, parity tests, pipe-rate probes, bench variants, ncu.
# Everything lands in gpurun_out/.  Each step is bounded by `timeout`.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/.."
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/gpu.csv 2>&1
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -5
echo "== microbench"
timeout 120 python - <<'PY' 2>&1 | tee gpurun_out/microbench.txt
import ctypes, discorpy_b200 as dcb
from discorpy_b200 import _cabi
dcb.set_device(0)
names = {0: "DFMA", 1: "F2F f32<->f64 (pairs)", 2: "MUFU.RSQ64H", 3: "radial coords px", 4: "FFMA", 5: "DFMA+2xF2F mixed"}
for w in (4, 0, 1, 2, 5, 3):
    g = ctypes.c_double()
    _cabi.call("dcb_microbench", w, ctypes.byref(g))
    print("microbench %d %-24s %10.1f Gops/s" % (w, names[w], g.value))
PY
echo "== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.txt
echo "== bench variants"
for blend in exact lerp64 lerp32; do for path in auto direct; do
  timeout 300 python bench.py --steps 20 --warmup 3 --blend $blend --path $path --no-cpu-baseline --e2e-steps 2 2>&1 | tail -1 >> gpurun_out/bench_variants.jsonl
done; done
python - <<'PY'
import json
for l in open("gpurun_out/bench_variants.jsonl"):
    try:
        d = json.loads(l)
        print(d["config"]["blend"], d["config"]["path"], "value %.0f Mpix/s" % d["value"], "kernel %.1f us" % d["roofline"]["kernel_us"], "frac %.3f" % d["roofline"]["frac"], "e2e %.0f" % d["e2e"]["value"], d["clocks"])
    except Exception as e:
        print("bad line", l[:200])
PY
echo "== bench default"; timeout 600 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_default.json
echo "== bench reference"; timeout 600 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_reference.json
echo "== ncu launches"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launch_run.log 2>&1
tail -3 gpurun_out/launches.csv
echo "== ncu full"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:remap_tile -s 20 -c 2 -f -o gpurun_out/prof_r1 python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full_run.log 2>&1
ls -la gpurun_out
