#!/bin/bash
# Runs ON the GPU box: the shared pipeline's GPU test, a quick timing, and the bench under ncu's launch list (short timeouts)
tag=${1:-r02x}
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_host_records.py -m gpu -x -q 2>&1 | tail -n 3
timeout 200 python scripts/exp_hybrid.py host_records=-1 2>&1 | tee gpurun_out/${tag}_hybrid.log
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-arms > gpurun_out/${tag}_ncu_list.log 2>&1; echo "ncu list rc=$?"
tail -n 2 gpurun_out/${tag}_ncu_list.log | cut -c1-300; wc -l gpurun_out/${tag}_launches.csv
