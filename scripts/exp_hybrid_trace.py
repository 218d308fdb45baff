"""One traced call of the hybrid host pipeline (CMGPU_HYBRID_TRACE=1): python scripts/exp_hybrid_trace.py "opt=v,..." """
import os, sys, time
os.environ["CMGPU_HYBRID_TRACE"] = "1"
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
import bench
from threecore_b200 import bspgen
from threecore_b200.engine import Engine
n = 1 << 24
m, starts, ends = bench.make_workload(n)
eng = Engine(0); eng.load_bsp(m.data)
h_s = torch.from_numpy(starts).pin_memory(); h_e = torch.from_numpy(ends).pin_memory(); h_out = torch.empty((n, 56), dtype=torch.uint8).pin_memory()
for spec in sys.argv[1:]:
    opts = {kv.split("=")[0]: float(kv.split("=")[1]) for kv in spec.split(",")}
    for k, v in opts.items(): eng.set_option(k, v)
    for i in range(16):
        sys.stderr.write(f"---- {spec} call {i}\n"); sys.stderr.flush()
        t0 = time.perf_counter()
        eng.box_trace_hostptr(n, h_s.data_ptr(), h_e.data_ptr(), bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, h_out.data_ptr())
        sys.stderr.write(f"---- call took {(time.perf_counter()-t0)*1e3:.2f} ms\n"); sys.stderr.flush()
