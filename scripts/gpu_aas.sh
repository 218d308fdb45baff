#!/bin/bash
tag=${1:-r02y}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_aas.py -m gpu -x -q 2>&1 | tail -n 4
timeout 600 python scripts/aas_bench.py 2>&1 | tee gpurun_out/${tag}_aas.log
