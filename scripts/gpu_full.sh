#!/bin/bash
# (ncu commands carry short timeouts: a capture that hangs once cost this round 23 GPU-minutes)
# Runs ON the GPU box (via gpurun): smoke, GPU tests, bench (both arms), ncu launch list, one ncu --set full capture of the
# walk kernel, DRAM traffic of the bench's own launch.  usage: gpurun --timeout 1800 -- 'bash scripts/gpu_full.sh <tag>'
tag=${1:-rXX}
mkdir -p gpurun_out
nproc > gpurun_out/${tag}_host.txt; nvidia-smi -L >> gpurun_out/${tag}_host.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 1 gpurun_out/${tag}_smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -n 3 gpurun_out/${tag}_pytest_gpu.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/${tag}_bench.json
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-arms > gpurun_out/${tag}_ncu_list.log 2>&1; echo "ncu list rc=$?"
python scripts/prof_trace.py 4194304 3 > gpurun_out/${tag}_plain.log 2>&1 && \
timeout 240 ncu --set full --clock-control none --import-source on -k regex:tracePool -s 1 -c 1 -f -o gpurun_out/prof_${tag} \
    python scripts/prof_trace.py 4194304 3 > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full rc=$?"
# DRAM traffic of the dominant kernels at the bench's own launch size (16 Mi rays)
timeout 240 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'tracePool|expandRecords' -s 6 -c 2 --csv \
    --log-file gpurun_out/${tag}_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra-arms > gpurun_out/${tag}_ncu_traffic.log 2>&1; echo "ncu traffic rc=$?"
