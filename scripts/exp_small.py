"""Experiment: latency of small resident launches vs the smallest cursor chunk.  usage: exp_small.py [min_chunk ...]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
import bench
from threecore_b200 import bspgen
from threecore_b200.engine import Engine
n = 1 << 21
m, starts, ends = bench.make_workload(n)
dev = torch.device("cuda", 0)
ds = torch.from_numpy(starts).to(dev); de = torch.from_numpy(ends).to(dev)
d_out = torch.empty((n, 56), dtype=torch.uint8, device=dev)
eng = Engine(0); eng.load_bsp(m.data)
for mc in [int(a) for a in sys.argv[1:]] or [32, 16, 8, 4]:
    eng.set_option("min_chunk", mc)
    row = []
    for k in (1024, 4096, 32768, 65536, 196608, 655360):
        o = d_out[:k]
        for _ in range(3): eng.box_trace_device(ds[:k], de[:k], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, o)
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): eng.box_trace_device(ds[:k], de[:k], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, o)
        b.record(); torch.cuda.synchronize()
        row.append(f"n={k}: {a.elapsed_time(b)*100:.0f} us")
    print(f"min_chunk {mc:2d}  " + "  ".join(row), flush=True)
