"""config 3 run for ncu: patches map, mixed point/hull traces"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
from threecore_b200 import bspgen
from threecore_b200.engine import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
dev = torch.device("cuda", 0)
eng = Engine(0)
m = bspgen.patch_map(n_brushes=2000, n_patches=512, seed=0xC0FFEE03); eng.load_bsp(m.data)
s, e, mins, maxs = bspgen.make_patch_sweeps(m, n, 0xBADC0DE3)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
ds, de, dmn, dmx = T(s), T(e), T(mins), T(maxs); out = torch.empty((n, 56), dtype=torch.uint8, device=dev)
for _ in range(3): eng.box_trace_device(ds, de, dmn, dmx, bspgen.MASK_ALL, out)
torch.cuda.synchronize(); eng.check_status(); print("ok")
