"""How much do brushes with non-axial sides cost the hot kernel?  Same workload on arenas with 0 / 10 / 30 % rotated brushes."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
import bench
from threecore_b200 import bspgen
from threecore_b200.engine import Engine
n = 1 << 24
dev = torch.device("cuda", 0)
for rf in (0.0, 0.10, 0.30):
    m = bspgen.arena_map(n_brushes=bench.N_BRUSHES, seed=bench.MAP_SEED, rotated_frac=rf)
    s, e = bspgen.make_sweeps(m, n, bench.RAY_SEED, z_range=(24.0, 2048.0))
    ds, de = torch.from_numpy(s).to(dev), torch.from_numpy(e).to(dev)
    out = torch.empty((n, 56), dtype=torch.uint8, device=dev)
    eng = Engine(0); eng.load_bsp(m.data)
    for _ in range(2): eng.box_trace_device(ds, de, bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, out)
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): eng.box_trace_device(ds, de, bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, out)
    b.record(); torch.cuda.synchronize(); eng.check_status()
    ms = a.elapsed_time(b) / 3
    print(f"rotated_frac {rf:.2f}: {ms:6.2f} ms  {n/ms/1e3:7.1f} Mtraces/s  hit {float((out[:, 8:12].view(torch.float32) < 1).float().mean()):.2f}", flush=True)
    eng.close()
