#!/bin/bash
tag=${1:-r02s}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_configs.py tests/test_gpu_at_size.py tests/test_cm_api.py tests/test_sv_world.py -m gpu -x -q 2>&1 | tail -n 6
timeout 600 python scripts/config_bench.py 2>&1 | tee gpurun_out/${tag}_configs.log
