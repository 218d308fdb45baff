#!/bin/bash
# option sweeps of the bench workload on one B200: gpurun -- 'bash scripts/gpu_exp.sh <tag> <lib or -> "opt=v,opt=v" ...'
tag=$1; lib=$2; shift 2
mkdir -p gpurun_out
if [ "$lib" != "-" ]; then export CMGPU_LIB=$lib; fi
python scripts/exp_opts.py "$@" 2>&1 | tee gpurun_out/${tag}_exp.log
