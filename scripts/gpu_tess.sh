#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_loader.py -m gpu -x -q 2>&1 | tail -n 8
python - <<'P'
import time, sys
sys.path.insert(0, '.')
from threecore_b200 import bspgen
from threecore_b200.engine import FlatMap
m = bspgen.patch_map(n_brushes=2000, n_patches=512, seed=0xC0FFEE03)
FlatMap.from_bsp(m.data, tess_device=0)
for name, dev in (("host", None), ("gpu", 0)):
    t0 = time.perf_counter(); FlatMap.from_bsp(m.data, tess_device=dev); print(f"load 512-patch map, patches on {name}: {1e3*(time.perf_counter()-t0):.1f} ms", flush=True)
P
