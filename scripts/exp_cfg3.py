"""config 3 step time under option settings: python scripts/exp_cfg3.py "facet_steps=1" "facet_steps=4" ..."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
from threecore_b200 import bspgen
from threecore_b200.engine import Engine
dev = torch.device("cuda", 0)
eng = Engine(0)
m = bspgen.patch_map(n_brushes=2000, n_patches=512, seed=0xC0FFEE03); eng.load_bsp(m.data)
n = 1 << 22
s, e, mins, maxs = bspgen.make_patch_sweeps(m, n, 0xBADC0DE3)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
ds, de, dmn, dmx = T(s), T(e), T(mins), T(maxs); out = torch.empty((n, 56), dtype=torch.uint8, device=dev)
ref = None
for spec in ["base"] + sys.argv[1:]:
    if spec != "base":
        for kv in spec.split(","):
            eng.set_option(kv.split("=")[0], float(kv.split("=")[1]))
    fn = lambda: eng.box_trace_device(ds, de, dmn, dmx, bspgen.MASK_ALL, out)
    for _ in range(2): fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(4): fn()
    b.record(); torch.cuda.synchronize(); eng.check_status()
    ms = a.elapsed_time(b) / 4
    same = "" if ref is None else f" same={bool(torch.equal(out, ref))}"
    if ref is None: ref = out.clone()
    print(f"{spec:30s} {ms:8.2f} ms  {n / ms / 1e3:7.1f} Mtraces/s{same}", flush=True)
