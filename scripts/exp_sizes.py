"""Resident launch time by size and kernel (which kernel should serve which launch size)."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
import bench
from threecore_b200 import bspgen
from threecore_b200.engine import Engine
n = 1 << 22
m, starts, ends = bench.make_workload(n)
dev = torch.device("cuda", 0)
ds = torch.from_numpy(starts).to(dev); de = torch.from_numpy(ends).to(dev)
d_out = torch.empty((n, 56), dtype=torch.uint8, device=dev)
eng = Engine(0); eng.load_bsp(m.data)
for k in (16384, 65536, 131072, 196608, 262144, 393216, 524288, 786432, 1 << 20, 1 << 21):
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
        for kk, v in dict(pool=1, pool_min_items=196609, simple_max_items=786432).items(): eng.set_option(kk, v)
    print(f"n={k:8d}: " + "   ".join(row), flush=True)
