#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_cm_api.py tests/test_sv_world.py -m gpu -x -q 2>&1 | tail -n 4
timeout 300 python scripts/latency_small.py 2>&1 | tee gpurun_out/r02_latency_small.log
