#!/bin/bash
# Runs ON the GPU box: host-side record expansion microbench, smoke, GPU tests, bench.  usage: gpurun --timeout 1500 -- 'bash scripts/gpu_check.sh <tag>'
tag=${1:-r02x}
mkdir -p gpurun_out
nproc > gpurun_out/${tag}_host.txt; nvidia-smi -L >> gpurun_out/${tag}_host.txt; lscpu | head -20 >> gpurun_out/${tag}_host.txt
for T in 1 2 4 8 12 16; do for m in 0 1 3; do scripts/_bin/hostexp_bench 8388608 $T $m | tail -n 1; done; done > gpurun_out/${tag}_hostexp.log 2>&1
cat gpurun_out/${tag}_hostexp.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/${tag}_smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -n 14 gpurun_out/${tag}_pytest_gpu.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
cut -c1-400 gpurun_out/${tag}_bench.json; tail -n 5 gpurun_out/${tag}_bench.err
