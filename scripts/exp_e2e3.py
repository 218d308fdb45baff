import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
import bench
from threecore_b200 import bspgen
from threecore_b200.engine import Engine
n = 1 << 24
m, starts, ends = bench.make_workload(n)
h_s = torch.from_numpy(starts).pin_memory(); h_e = torch.from_numpy(ends).pin_memory()
h_out = torch.empty((n, 56), dtype=torch.uint8).pin_memory()
eng = Engine(0); eng.load_bsp(m.data)
for pm in (1 << 20, 1 << 19, 1 << 18):
    eng.set_option("pool_min_items", pm)
    for chunks in (12, 16, 24, 32, 48):
        eng.set_option("pipe_chunks", chunks)
        for _ in range(2):
            eng.box_trace_hostptr(n, h_s.data_ptr(), h_e.data_ptr(), bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, h_out.data_ptr())
        t = time.perf_counter()
        for _ in range(4):
            eng.box_trace_hostptr(n, h_s.data_ptr(), h_e.data_ptr(), bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, h_out.data_ptr())
        dt = (time.perf_counter() - t) / 4
        print(f"pool_min_items {pm:8d} pipe_chunks {chunks:3d}: {dt*1e3:7.2f} ms  {n/dt/1e6:8.1f} Mtraces/s", flush=True)
dev = torch.device("cuda", 0)
ds = torch.from_numpy(starts).to(dev); de = torch.from_numpy(ends).to(dev)
d_out = torch.empty((n, 56), dtype=torch.uint8, device=dev)
print("--- resident, by size and kernel")
for k in (262144, 524288, 1 << 20, 1 << 21, 1 << 22):
    row = []
    for label, opts in (("simple", dict(simple_max_items=1 << 30, pool=0)), ("persistent", dict(simple_max_items=0, pool=0)), ("pool", dict(simple_max_items=0, pool=1, pool_min_items=1))):
        for kk, v in opts.items(): eng.set_option(kk, v)
        o = d_out[:k]
        for _ in range(3): eng.box_trace_device(ds[:k], de[:k], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, o)
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): eng.box_trace_device(ds[:k], de[:k], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, o)
        b.record(); torch.cuda.synchronize()
        row.append(f"{label} {a.elapsed_time(b)*100:.0f} us")
        for kk, v in dict(pool=1, pool_min_items=1 << 20, simple_max_items=393216).items(): eng.set_option(kk, v)
    print(f"n={k:8d}: " + "   ".join(row), flush=True)
