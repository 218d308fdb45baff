#!/bin/bash
# Build a variant of libcmgpu.so with extra -D flags into build/variants/<name>.so (experiments: CMGPU_LIB=... python ...)
# usage: scripts/build_variant.sh <name> [-DCMGPU_X=1 ...]
set -e
name=$1; shift
cd "$(dirname "$0")/../threecore_b200/csrc"
mkdir -p ../../build/variants
nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
  -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math,-Wall,-fvisibility=hidden -Xptxas -v -cudart static "$@" \
  -shared -o ../../build/variants/$name.so cm_gpu.cu cm_host_records.cpp cm_load.cpp cm_patchgen.cpp cm_api.cpp sv_world_batch.cpp > ../../build/variants/$name.log 2>&1
grep -A2 "tracePoolKernelILb0ELi3ELb1" ../../build/variants/$name.log | grep -i "registers\|spill" || true
