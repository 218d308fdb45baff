#!/bin/bash
tag=${1:-r02t}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:traceKernel -s 2 -c 1 -f -o gpurun_out/prof_${tag}_cfg3 python scripts/prof_cfg3.py > gpurun_out/${tag}_ncu_cfg3.log 2>&1; echo "ncu rc=$?"; tail -n 2 gpurun_out/${tag}_ncu_cfg3.log
