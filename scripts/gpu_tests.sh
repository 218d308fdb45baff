#!/bin/bash
# Runs ON the GPU box: smoke, GPU tests, short bench (both arms).  usage: gpurun --timeout 1500 -- 'bash scripts/gpu_tests.sh <tag>'
tag=${1:-r02x}
mkdir -p gpurun_out
nproc > gpurun_out/${tag}_host.txt; nvidia-smi -L >> gpurun_out/${tag}_host.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"; tail -n 2 gpurun_out/${tag}_smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/${tag}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -n 25 gpurun_out/${tag}_pytest_gpu.log
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
cut -c1-600 gpurun_out/${tag}_bench.json; tail -n 5 gpurun_out/${tag}_bench.err
