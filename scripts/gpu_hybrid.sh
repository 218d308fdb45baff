#!/bin/bash
# Runs ON the GPU box: sweeps of the host / device split of a host batch's records.  usage: gpurun --timeout 900 -- 'bash scripts/gpu_hybrid.sh <tag> <specs...>'
tag=${1:-r02x}; shift
mkdir -p gpurun_out
nproc > gpurun_out/${tag}_host.txt
timeout 600 python scripts/exp_hybrid.py "$@" > gpurun_out/${tag}_hybrid.log 2>&1; echo "rc=$?"
cat gpurun_out/${tag}_hybrid.log | tail -n 40
