"""Experiment: kernel time by ray class (short / long / zero-length / in-solid), unsorted and Morton-sorted."""
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

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
m = bspgen.arena_map(n_brushes=bench.N_BRUSHES, seed=bench.MAP_SEED)
eng = Engine(0); eng.load_bsp(m.data)
dev = torch.device("cuda", 0)
d_out = torch.empty((n, 56), dtype=torch.uint8, device=dev)
def run(s, e, label):
    ds = torch.from_numpy(np.ascontiguousarray(s)).to(dev); de = torch.from_numpy(np.ascontiguousarray(e)).to(dev)
    for _ in range(2): eng.box_trace_device(ds, de, bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, d_out)
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): eng.box_trace_device(ds, de, bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, d_out)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 3
    print(f"{label:34s}: {ms:7.2f} ms  {n/ms/1e3:8.1f} Mtraces/s", flush=True)
lo = m.world_mins.astype(np.float64)
for label, kw in (("short 1-512", dict(len_short=(1, 512), len_long=(1, 512), long_frac=0.0, in_solid_frac=0.0, zero_len_frac=0.0)),
                  ("long 512-4096", dict(len_short=(512, 4096), len_long=(512, 4096), long_frac=0.0, in_solid_frac=0.0, zero_len_frac=0.0)),
                  ("tiny 0.25-32 (ground/step)", dict(len_short=(0.25, 32), len_long=(0.25, 32), long_frac=0.0, in_solid_frac=0.0, zero_len_frac=0.0)),
                  ("zero length", dict(len_short=(1, 512), len_long=(1, 512), long_frac=0.0, in_solid_frac=0.0, zero_len_frac=1.0)),
                  ("bench mix", dict())):
    s, e = bspgen.make_sweeps(m, n, 5, z_range=(24.0, 2048.0), **kw)
    run(s, e, label)
    k = morton(s.astype(np.float64), lo, 32.0)
    o = np.argsort(k, kind="stable")
    run(s[o], e[o], label + " sorted")
