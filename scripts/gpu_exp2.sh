#!/bin/bash
tag=${1:-r02d}
mkdir -p gpurun_out
for v in hoist hoist_nl2 hoist_nl4 hoist_le3; do
  echo "== $v"; CMGPU_LIB=build/variants/$v.so timeout 300 python scripts/exp_opts.py "gang=3" "gang=5" 2>&1 | tee -a gpurun_out/${tag}_exp.log
done
CMGPU_LIB=build/variants/hoist.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -n 3
