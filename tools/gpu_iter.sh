#!/bin/bash
# quick iteration: smoke + gpu tests + bench variants + ncu full; args: tag "flags list" "blend list"
set -u
cd "$(dirname "$0")/.."
tag=${1:-iter}; flags=${2:-0}; blends=${3:-"exact lerp32"}
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
echo "== pytest gpu"; timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
rm -f gpurun_out/bench_variants_$tag.jsonl
for f in $flags; do for blend in $blends; do
  DCB_FLAGS=$f timeout 300 python bench.py --steps 20 --warmup 3 --blend $blend --no-cpu-baseline --no-extras --e2e-steps 0 2>&1 | tail -1 | sed "s/^{/{\"flags\": $f, /" >> gpurun_out/bench_variants_$tag.jsonl
done; done
python - <<PY
import json
for l in open("gpurun_out/bench_variants_$tag.jsonl"):
    try:
        d = json.loads(l)
        print("flags", d["flags"], d["config"]["blend"], "value %.0f Mpix/s" % d["value"], "kernel %.1f us" % d["roofline"]["kernel_us"], "frac %.3f" % d["roofline"]["frac"], d["clocks"], d["config"]["plan"])
    except Exception as e:
        print("bad line", l[:300])
PY
