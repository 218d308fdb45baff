"""Experiment: kernel time with rays in input order vs host-sorted by Morton code of the start point."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
import bench
from threecore_b200 import bspgen
from threecore_b200.engine import Engine

def morton(starts, lo, cell):
    q = np.clip(((starts - lo) / cell).astype(np.int64), 0, 1023)
    def spread(v):
        v = v & 0x3ff
        v = (v | (v << 16)) & 0x30000ff
        v = (v | (v << 8)) & 0x300f00f
        v = (v | (v << 4)) & 0x30c30c3
        v = (v | (v << 2)) & 0x9249249
        return v
    return spread(q[:, 0]) | (spread(q[:, 1]) << 1) | (spread(q[:, 2]) << 2)

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
m, starts, ends = bench.make_workload(n)
eng = Engine(0); eng.load_bsp(m.data)
dev = torch.device("cuda", 0)
d_mins = bspgen.PLAYER_MINS; d_maxs = bspgen.PLAYER_MAXS
d_mask = bspgen.MASK_PLAYERSOLID
d_out = torch.empty((n, 56), dtype=torch.uint8, device=dev)
def run(s, e, label):
    ds = torch.from_numpy(np.ascontiguousarray(s)).to(dev); de = torch.from_numpy(np.ascontiguousarray(e)).to(dev)
    for _ in range(2): eng.box_trace_device(ds, de, d_mins, d_maxs, d_mask, d_out)
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): eng.box_trace_device(ds, de, d_mins, d_maxs, d_mask, d_out)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    print(f"{label}: {ms:.2f} ms  {n/ms/1e3:.1f} Mtraces/s", flush=True)
run(starts, ends, "input order")
lo = m.world_mins.astype(np.float64)
for cell in (16.0, 64.0):
    k = morton(starts.astype(np.float64), lo, cell)
    o = np.argsort(k, kind="stable")
    run(starts[o], ends[o], f"morton cell {cell}")
ln = np.linalg.norm(ends - starts, axis=1)
k = morton(starts.astype(np.float64), lo, 32.0) | ((ln > 512).astype(np.int64) << 40)
o = np.argsort(k, kind="stable")
run(starts[o], ends[o], "lenclass+morton32")
mid = 0.5 * (starts + ends)
k = morton(mid.astype(np.float64), lo, 32.0) | ((ln > 512).astype(np.int64) << 40)
o = np.argsort(k, kind="stable")
run(starts[o], ends[o], "lenclass+morton32(mid)")
