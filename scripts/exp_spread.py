"""Launch time of the thread-per-item kernel by size and lanes per item (simple_spread; -1 = automatic)."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
import bench
from threecore_b200 import bspgen
from threecore_b200.engine import Engine
n = 1 << 20
m, starts, ends = bench.make_workload(n)
dev = torch.device("cuda", 0)
ds = torch.from_numpy(starts).to(dev); de = torch.from_numpy(ends).to(dev)
d_out = torch.empty((n, 56), dtype=torch.uint8, device=dev)
eng = Engine(0); eng.load_bsp(m.data)
eng.set_option("simple_max_items", 1 << 30); eng.set_option("pool", 0)
sizes = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [64, 1024, 4096, 16384, 65536, 196608, 524288]
ref = {}
for k in sizes:
    row = []
    for sp in (0, 1, 2, 3, 4, 5, -1):
        eng.set_option("simple_spread", sp)
        o = d_out[:k]
        for _ in range(3): eng.box_trace_device(ds[:k], de[:k], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, o)
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): eng.box_trace_device(ds[:k], de[:k], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, o)
        b.record(); torch.cuda.synchronize(); eng.check_status()
        if k not in ref: ref[k] = o.clone()
        ok = bool(torch.equal(o, ref[k]))
        row.append(f"s{sp}:{a.elapsed_time(b)*100:.0f}us{'' if ok else '!DIFF'}")
    print(f"n={k:7d}: " + "  ".join(row), flush=True)
