#!/bin/bash
# quick iteration: smoke + gpu tests + bench variants + default bench + ncu full; arg: tag
set -u
cd "$(dirname "$0")/.."
tag=${1:-iter}
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
echo "== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
rm -f gpurun_out/bench_variants_$tag.jsonl
for blend in exact lerp64 lerp32; do
  timeout 300 python bench.py --steps 20 --warmup 3 --blend $blend --no-cpu-baseline --e2e-steps 2 2>&1 | tail -1 >> gpurun_out/bench_variants_$tag.jsonl
done
python - <<PY
import json
for l in open("gpurun_out/bench_variants_$tag.jsonl"):
    try:
        d = json.loads(l)
        print(d["config"]["blend"], d["config"]["path"], "value %.0f Mpix/s" % d["value"], "kernel %.1f us" % d["roofline"]["kernel_us"], "frac %.3f" % d["roofline"]["frac"], "e2e %.0f" % d["e2e"]["value"], d["clocks"], d["config"]["plan"])
    except Exception as e:
        print("bad line", l[:300])
PY
bash tools/gpu_prof.sh prof_$tag
