"""Short single-GPU run of the bench workload for ncu (fewer rays, a few launches)."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
import bench
from threecore_b200 import bspgen
from threecore_b200.engine import Engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 21
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
m, starts, ends = bench.make_workload(n)
eng = Engine(0)
eng.load_bsp(m.data)
for kv in (sys.argv[3].split(',') if len(sys.argv) > 3 else []):
    eng.set_option(kv.split('=')[0], float(kv.split('=')[1]))
dev = torch.device("cuda", 0)
d_starts = torch.from_numpy(starts).to(dev); d_ends = torch.from_numpy(ends).to(dev)
d_mins = bspgen.PLAYER_MINS; d_maxs = bspgen.PLAYER_MAXS
d_mask = bspgen.MASK_PLAYERSOLID
d_out = torch.empty((n, 56), dtype=torch.uint8, device=dev)
for _ in range(steps):
    eng.box_trace_device(d_starts, d_ends, d_mins, d_maxs, d_mask, d_out)
torch.cuda.synchronize()
eng.check_status()
print("ok", n, steps, float((d_out[:, 8:12].view(torch.float32) < 1).float().mean()))
