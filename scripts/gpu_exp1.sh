#!/bin/bash
tag=${1:-r02b}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_rollout.py tests/test_gpu_at_size.py -m gpu -x -q > gpurun_out/${tag}_pytest.log 2>&1; echo "pytest rc=$?"; tail -n 4 gpurun_out/${tag}_pytest.log
timeout 600 python scripts/exp_opts.py "entry_grid=0" "entry_grid=1" "entry_cell=64" "entry_cell=256" "entry_cell=128" "gang_brush=4" "gang_brush=6" "pool_reps=2" > gpurun_out/${tag}_exp_default.log 2>&1; echo "exp rc=$?"; cat gpurun_out/${tag}_exp_default.log
CMGPU_LIB=build/variants/slab1.so timeout 600 python scripts/exp_opts.py "entry_grid=0" "entry_grid=1" > gpurun_out/${tag}_exp_slab1.log 2>&1; echo "exp rc=$?"; cat gpurun_out/${tag}_exp_slab1.log
CMGPU_LIB=build/variants/slab1.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/${tag}_pytest_slab1.log 2>&1; echo "pytest slab1 rc=$?"; tail -n 4 gpurun_out/${tag}_pytest_slab1.log
