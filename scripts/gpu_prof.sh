#!/bin/bash
# ncu --set full of the walk kernel on 4 Mi rays of the bench workload.  usage: gpurun -- 'bash scripts/gpu_prof.sh <tag> [lib] [opts]'
tag=${1:-r02p}; lib=${2:-threecore_b200/libcmgpu.so}; opts=${3:-}
mkdir -p gpurun_out
CMGPU_LIB=$lib python scripts/prof_trace.py 4194304 3 $opts > gpurun_out/${tag}_plain.log 2>&1 && \
CMGPU_LIB=$lib ncu --set full --clock-control none --import-source on -k regex:tracePool -s 1 -c 1 -f -o gpurun_out/prof_${tag} \
    python scripts/prof_trace.py 4194304 3 $opts > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
tail -n 3 gpurun_out/${tag}_ncu_full.log
