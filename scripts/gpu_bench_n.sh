#!/bin/bash
# usage: gpurun --gpus N -- 'bash scripts/gpu_bench_n.sh <tag> N'
tag=${1:-r02f}; N=${2:-1}
mkdir -p gpurun_out
nproc > gpurun_out/${tag}_host.txt; nvidia-smi -L >> gpurun_out/${tag}_host.txt; nvidia-smi topo -m >> gpurun_out/${tag}_host.txt 2>&1
if [ "$N" -gt 1 ]; then
  timeout 600 python -m pytest tests/test_peer_results.py -m gpu -q > gpurun_out/${tag}_peer.log 2>&1; echo "peer rc=$?"; tail -n 3 gpurun_out/${tag}_peer.log
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_bench_n$N.err; echo "bench rc=$?"
else
  timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_bench_n$N.err; echo "bench rc=$?"
fi
tail -n 5 gpurun_out/${tag}_bench_n$N.err
python - <<P
import json
d=json.load(open("gpurun_out/${tag}_bench_n$N.json"))
for k in ("value","ms_per_step","e2e","e2e_pageable","gathered","strong","rollout","parity"):
    print(k, json.dumps(d.get(k))[:700])
P
