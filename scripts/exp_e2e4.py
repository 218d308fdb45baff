import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
import bench
from threecore_b200 import bspgen
from threecore_b200.engine import Engine, traces_from_tensor
n = 1 << 24
m, starts, ends = bench.make_workload(n)
h_s = torch.from_numpy(starts).pin_memory(); h_e = torch.from_numpy(ends).pin_memory()
h_out = torch.empty((n, 56), dtype=torch.uint8).pin_memory()
eng = Engine(0); eng.load_bsp(m.data)
ref = None
for ramp in (0, 1):
    eng.set_option("pipe_ramp", ramp)
    for chunks in (8, 16, 32):
        eng.set_option("pipe_chunks", chunks)
        for _ in range(2):
            eng.box_trace_hostptr(n, h_s.data_ptr(), h_e.data_ptr(), bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, h_out.data_ptr())
        t = time.perf_counter()
        for _ in range(4):
            eng.box_trace_hostptr(n, h_s.data_ptr(), h_e.data_ptr(), bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, h_out.data_ptr())
        dt = (time.perf_counter() - t) / 4
        if ref is None: ref = h_out.clone()
        print(f"ramp {ramp} pipe_chunks {chunks:3d}: {dt*1e3:7.2f} ms  {n/dt/1e6:8.1f} Mtraces/s same={bool(torch.equal(ref, h_out))}", flush=True)
