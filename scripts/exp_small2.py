"""Small resident launches: whole call vs the trace kernel alone (binning and launch gaps are the difference)."""
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
for opts in [dict(), dict(bin=0), dict(bin_min_items=1 << 17), dict(min_chunk=4), dict(gang=2), dict(gang=1, gang_fetch=1), dict(light_reps=6)]:
    for k, v in opts.items(): eng.set_option(k, v)
    row = []
    for k in (4096, 65536, 196608):
        o = d_out[:k]
        for _ in range(3): eng.box_trace_device(ds[:k], de[:k], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, o)
        torch.cuda.synchronize()
        eng.kernel_timing(True)
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): eng.box_trace_device(ds[:k], de[:k], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, o)
        b.record(); torch.cuda.synchronize()
        kms, nl = eng.kernel_time_ms(); eng.kernel_timing(False)
        row.append(f"n={k}: call {a.elapsed_time(b)*100:.0f} us, kernel {kms/nl*1e3:.0f} us")
    print(f"{str(opts):28s} " + "   ".join(row), flush=True)
    for k, v in dict(bin=1, bin_min_items=1 << 15, min_chunk=8, gang=4, gang_fetch=4, light_reps=3).items(): eng.set_option(k, v)
print("--- thread-per-item kernel vs persistent kernel by size")
for k in (1024, 4096, 16384, 65536, 131072, 262144, 524288, 1 << 20):
    row = []
    for label, opts in (("simple", dict(simple_max_items=1 << 30, pool=0)), ("simple,bin0", dict(simple_max_items=1 << 30, pool=0, bin=0)), ("persistent", dict(simple_max_items=0, pool=0)), ("pool", dict(simple_max_items=0, pool=1, pool_min_items=1))):
        for kk, v in opts.items(): eng.set_option(kk, v)
        o = d_out[:k]
        for _ in range(3): eng.box_trace_device(ds[:k], de[:k], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, o)
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10): eng.box_trace_device(ds[:k], de[:k], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, o)
        b.record(); torch.cuda.synchronize()
        row.append(f"{label} {a.elapsed_time(b)*100:.0f} us")
        for kk, v in dict(bin=1, pool=1, pool_min_items=1 << 20, simple_max_items=1 << 17).items(): eng.set_option(kk, v)
    print(f"n={k:8d}: " + "   ".join(row), flush=True)
