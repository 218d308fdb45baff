#!/bin/bash
tag=${1:-r02y}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_aas.py -m gpu -x -q 2>&1 | tail -n 12
