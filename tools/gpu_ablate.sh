#!/bin/bash
# ablation runs (A/B library only): args: tag "flags list" "blend list"
set -u
cd "$(dirname "$0")/.."
tag=${1:-abl}; flags=${2:-0}; blends=${3:-"exact lerp32"}
mkdir -p gpurun_out
timeout 120 python - <<'PY' 2>&1 | tee gpurun_out/microbench_$tag.txt
import ctypes, discorpy_b200 as dcb
from discorpy_b200 import _cabi
dcb.set_device(0)
names = {6: "DFMA latency (cycles)", 7: "F2F f->d->f latency (cycles, 2 cvt)", 8: "MUFU.RSQ64H latency (cycles)", 9: "LDS.64 latency (cycles, +LOP)"}
for w in (6, 7, 8, 9):
    g = ctypes.c_double()
    _cabi.call("dcb_microbench", w, ctypes.byref(g))
    print("microbench %d %-36s %10.2f" % (w, names[w], g.value))
PY
rm -f gpurun_out/bench_variants_$tag.jsonl
for f in $flags; do for blend in $blends; do
  DCB_FLAGS=$f timeout 300 python bench.py --steps 20 --warmup 3 --blend $blend --no-cpu-baseline --no-extras --e2e-steps 0 2>&1 | tail -1 | sed "s/^{/{\"flags\": $f, /" >> gpurun_out/bench_variants_$tag.jsonl
done; done
python - <<PY
import json
for l in open("gpurun_out/bench_variants_$tag.jsonl"):
    try:
        d = json.loads(l)
        print("flags", d["flags"], d["config"]["blend"], "kernel %.1f us" % d["roofline"]["kernel_us"], "frac %.3f" % d["roofline"]["frac"], d["config"]["plan"])
    except Exception as e:
        print("bad line", l[:300])
PY
