#!/bin/bash
# One GPU round: smoke, the GPU parity suite, the bench lines (default + blend
# variants + Z-stack cases), an ncu launch list of the bench command and ncu
# --set full captures of the single-image and the Z-stack kernel.
# Usage: tools/gpu_round.sh TAG [skip-list]   (skip-list: words out of
#        "tests bench stack ncu"); everything lands in gpurun_out/.
set -u
cd "$(dirname "$0")/.."
tag=${1:-round}; skip=${2:-}
mkdir -p gpurun_out
has() { case " $skip " in *" $1 "*) return 0;; *) return 1;; esac; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,clocks_throttle_reasons.active --format=csv | tail -2
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -3
if ! has tests; then
  echo "== pytest gpu"; timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 | tee gpurun_out/pytest_gpu_$tag.txt
fi
if ! has bench; then
  echo "== bench default"
  timeout 900 python bench.py 2> gpurun_out/bench_default_$tag.err | tail -1 > gpurun_out/bench_default_$tag.json
  cut -c1-1500 gpurun_out/bench_default_$tag.json; tail -3 gpurun_out/bench_default_$tag.err
  rm -f gpurun_out/bench_variants_$tag.jsonl
  for blend in lerp64 lerp32; do
    timeout 300 python bench.py --steps 20 --warmup 3 --blend $blend --no-cpu-baseline --no-extras --e2e-steps 0 2>&1 | tail -1 >> gpurun_out/bench_variants_$tag.jsonl
  done
  echo "== bench reference arm"
  timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_reference_$tag.json | cut -c1-600
fi
if ! has stack; then
  echo "== stack bench"
  timeout 900 python tools/bench_stack.py --reps 5 2>&1 | tee gpurun_out/bench_stack_$tag.jsonl | cut -c1-260
fi
if ! has ncu; then
  echo "== ncu launch list"
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
      python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --e2e-steps 1 > gpurun_out/launches_$tag.log 2>&1
  tail -1 gpurun_out/launches_$tag.log | cut -c1-200
  echo "== ncu full: single-image kernel"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:remap_image -s 20 -c 1 -f -o gpurun_out/ncu_image_$tag \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 0 > gpurun_out/ncu_image_$tag.log 2>&1
  echo "== ncu full: Z-stack kernel (exact, 16 x 4096^2)"
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:remap_stack -s 2 -c 1 -f -o gpurun_out/ncu_stack_$tag \
      python tools/bench_stack.py --cases cfg2x16 --blends exact --reps 1 > gpurun_out/ncu_stack_$tag.log 2>&1
  ls -la gpurun_out/*.ncu-rep
fi
echo "== done"
