#!/bin/bash
# A/B build of the library: tools/build_ab.sh name "-DFLAG=1 ..."  ->  discorpy_b200/lib/ab/libdcb_<name>.so
# (load it with DCB_LIB=...; lib/ab travels to the GPU box, it is git-ignored)
set -eu
cd "$(dirname "$0")/../discorpy_b200/csrc"
name=$1; shift
mkdir -p ../lib/ab
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I../../include \
  -DDCB_NT_ONLY=5 "$@" -shared -o ../lib/ab/libdcb_$name.so api.cu diag.cu mg.cu hostpipe.cu -ldl
ls -la ../lib/ab/libdcb_$name.so
