#!/bin/bash
tag=${1:-r02h}; N=${2:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_peer_results.py -m gpu -q 2>&1 | tail -n 3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 scripts/peer_bench.py 16777216 \
  "remote_pipeline=1,remote_pieces=4,remote_expand_ctas=64" "remote_expand_ctas=16" "remote_expand_ctas=32" "remote_expand_ctas=128" "remote_expand_ctas=296" \
  "remote_pieces=2,remote_expand_ctas=64" "remote_pieces=8" "remote_pieces=16" 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tee gpurun_out/${tag}_peer_n$N.log
