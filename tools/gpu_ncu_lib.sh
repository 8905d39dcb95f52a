#!/bin/bash
# ncu --set full of the single-image kernel of one library build: tools/gpu_ncu_lib.sh tag lib blend
set -u
cd "$(dirname "$0")/.."
tag=${1:-ncu}; lib=${2:-default}; blend=${3:-exact}
mkdir -p gpurun_out
if [ "$lib" != default ]; then export DCB_LIB=$PWD/discorpy_b200/lib/ab/libdcb_$lib.so; fi
timeout 900 ncu --set full --clock-control none --import-source on -k regex:remap_image -s 20 -c 1 -f -o gpurun_out/ncu_img_$tag python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --e2e-steps 0 --blend $blend > gpurun_out/ncu_img_$tag.log 2>&1
ls -la gpurun_out/ncu_img_$tag.ncu-rep
