#!/bin/bash
tag=${1:-r02j}; N=${2:-8}
mkdir -p gpurun_out
nproc > gpurun_out/${tag}_host.txt; nvidia-smi -L >> gpurun_out/${tag}_host.txt; nvidia-smi topo -m >> gpurun_out/${tag}_host.txt 2>&1
timeout 600 python -m pytest tests/test_peer_results.py -m gpu -q > gpurun_out/${tag}_peer_test.log 2>&1; echo "peer test rc=$?"; tail -n 3 gpurun_out/${tag}_peer_test.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 scripts/peer_bench.py 16777216 \
  "remote_pipeline=1,remote_pieces=4,remote_expand_ctas=64" "remote_expand_ctas=148" "remote_pieces=2,remote_expand_ctas=64" "remote_pieces=8,remote_expand_ctas=64" 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tee gpurun_out/${tag}_peer_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/${tag}_bench_n$N.json 2> gpurun_out/${tag}_bench_n$N.err; echo "bench rc=$?"
tail -n 3 gpurun_out/${tag}_bench_n$N.err
python - <<P
import json
d=json.load(open("gpurun_out/${tag}_bench_n$N.json"))
for k in ("value","ms_per_step","e2e","e2e_pageable","gathered","strong","rollout","parity"):
    print(k, json.dumps(d.get(k))[:600])
P
