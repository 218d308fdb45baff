#!/bin/bash
# step time of the bench workload for a list of library builds: gpurun -- 'bash scripts/gpu_variants.sh <tag> name1 name2 ...' (build/variants/<name>.so; "-" = the product library)
tag=$1; shift
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = "-" ]; then unset CMGPU_LIB; else export CMGPU_LIB=build/variants/$v.so; fi
  echo "== $v"; python scripts/exp_opts.py 2>&1 | tail -n 1
done | tee gpurun_out/${tag}_variants.log
