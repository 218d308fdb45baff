#!/bin/bash
tag=${1:-r02e}
mkdir -p gpurun_out
echo "== default"; timeout 300 python scripts/exp_opts.py "gang_brush=4" 2>&1 | tee -a gpurun_out/${tag}_exp.log
for v in pairs1 pairs2; do
  echo "== $v"; CMGPU_LIB=build/variants/$v.so timeout 300 python scripts/exp_opts.py "gang=4" 2>&1 | tee -a gpurun_out/${tag}_exp.log
done
CMGPU_LIB=build/variants/pairs2.so timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -n 3
