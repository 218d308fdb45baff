"""Small launches of the kernel paths added in round 2 (entry grid, patches as warp work, AAS tracer and predictor, calls of
a few items, GPU tessellation) for `compute-sanitizer --tool memcheck|racecheck python scripts/sanitize_r02.py`."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
from helpers import get_map
from threecore_b200 import aasgen, bspgen
from threecore_b200.engine import Engine
eng = Engine(0)
dev = torch.device("cuda", 0)
m = get_map("arena_small")
eng.load_bsp(m.data)
n = 40001
s, e = bspgen.make_sweeps(m, n, 9)
ds, de = torch.from_numpy(s).to(dev), torch.from_numpy(e).to(dev)
for k, v in dict(pool_min_items=1, simple_max_items=0, bin_min_items=1).items():
    eng.set_option(k, v)
o = torch.zeros((n, 56), dtype=torch.uint8, device=dev)
eng.box_trace_device(ds, de, bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, o)     # pool kernel + entry grid
torch.cuda.synchronize(); eng.check_status()
eng.set_option("entry_grid", 0)
o2 = torch.zeros_like(o)
eng.box_trace_device(ds, de, bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, o2)
torch.cuda.synchronize(); assert torch.equal(o, o2)
eng.box_trace(s[:5], e[:5], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, mask=bspgen.MASK_PLAYERSOLID)     # mapped-memory path
pm = get_map("patches_small")
eng.load_bsp(pm.data, gpu_patches=True)                                                              # GPU tessellation
ps, pe, mins, maxs = bspgen.make_patch_sweeps(pm, 20000, 5)
eng.box_trace(ps, pe, mins, maxs, mask=bspgen.MASK_ALL)                                               # patches as warp work
w = aasgen.make_world(n_boxes=60, seed=3, extent=2048.0, height=512.0)
eng.aas_upload(w)
ls, le = aasgen.make_lines(w, 20000, 1)
eng.aas_trace_client_bbox(ls, le, 2)
eng.aas_predict_client_movement(aasgen.make_moves(w, 5000, 2))
print("sanitize_r02 ok")
