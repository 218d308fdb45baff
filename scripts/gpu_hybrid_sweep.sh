#!/bin/bash
# Runs ON the GPU box: static sweeps of the shared pipeline
tag=${1:-r02x}; shift
mkdir -p gpurun_out
nproc > gpurun_out/${tag}_host.txt
( timeout 500 python scripts/exp_hybrid.py "$@" 2>&1 ) | tee gpurun_out/${tag}_sweep.log
