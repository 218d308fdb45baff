"""End-to-end call (pinned host buffers, cmgpu_box_trace_batch) under option settings: python scripts/exp_e2e.py "pipe_pool_min_items=1" ..."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
import bench
from threecore_b200 import bspgen
from threecore_b200.engine import Engine
n = 1 << 24
m, starts, ends = bench.make_workload(n)
eng = Engine(0); eng.load_bsp(m.data)
h_s = torch.from_numpy(starts).pin_memory(); h_e = torch.from_numpy(ends).pin_memory()
h_out = torch.empty((n, 56), dtype=torch.uint8).pin_memory()
ref = None
for spec in ["base"] + sys.argv[1:] + ["base"]:
    opts = {} if spec == "base" else {kv.split("=")[0]: float(kv.split("=")[1]) for kv in spec.split(",")}
    for k, v in opts.items(): eng.set_option(k, v)
    f = lambda: eng.box_trace_hostptr(n, h_s.data_ptr(), h_e.data_ptr(), bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, h_out.data_ptr())
    for _ in range(2): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(5): f()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
    if ref is None: ref = h_out[: 1 << 20].clone()
    print(f"{spec:44s} {dt*1e3:7.2f} ms  {n/dt/1e6:7.1f} Mtraces/s same={bool(torch.equal(h_out[: 1 << 20], ref))}", flush=True)
    for k in opts:
        if ("D_" + k) in os.environ: eng.set_option(k, float(os.environ["D_" + k]))
