"""config 3 (patches) resident throughput for scheduling knobs of the general kernel"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
from threecore_b200 import bspgen
from threecore_b200.engine import Engine
n = 1 << 21
dev = torch.device("cuda", 0)
eng = Engine(0)
m = bspgen.patch_map(n_brushes=2000, n_patches=512, seed=0xC0FFEE03); eng.load_bsp(m.data)
s, e, mins, maxs = bspgen.make_patch_sweeps(m, n, 0xBADC0DE3)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
ds, de, dmn, dmx = T(s), T(e), T(mins), T(maxs); out = torch.empty((n, 56), dtype=torch.uint8, device=dev)
pt = (mins == 0).all(axis=1)
dsp, dep = T(s[pt]), T(e[pt]); dsh, deh = T(s[~pt]), T(e[~pt])
ref = None
for opts in [dict(), dict(facet_steps=2), dict(facet_steps=4), dict(facet_steps=6), dict(facet_steps=24), dict(facet_steps=4, light_reps=6), dict(facet_steps=6, light_reps=8), dict(facet_steps=24, light_reps=12)]:
    for k, v in opts.items(): eng.set_option(k, v)
    if "bin_cell" in opts: eng.load_bsp(m.data)
    def run(a, b, mn, mx, o):
        for _ in range(2): eng.box_trace_device(a, b, mn, mx, bspgen.MASK_ALL, o)
        x = torch.cuda.Event(enable_timing=True); y = torch.cuda.Event(enable_timing=True)
        x.record()
        for _ in range(3): eng.box_trace_device(a, b, mn, mx, bspgen.MASK_ALL, o)
        y.record(); torch.cuda.synchronize(); eng.check_status()
        return x.elapsed_time(y) / 3
    ms = run(ds, de, dmn, dmx, out)
    if ref is None: ref = out.clone()
    same = bool(torch.equal(ref, out))
    msp = run(dsp, dep, None, None, out[: dsp.shape[0]])
    msh = run(dsh, deh, bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, out[: dsh.shape[0]])
    print(f"{str(opts):32s} mixed {n/ms/1e3:7.1f} M/s same={same}   points alone {dsp.shape[0]/msp/1e3:7.1f}   hulls alone {dsh.shape[0]/msh/1e3:7.1f}", flush=True)
    for k, v in dict(gang=4, gang_fetch=4, light_reps=3, bin=1, bin_cell=256, facet_steps=12).items(): eng.set_option(k, v)
    if "bin_cell" in opts: eng.load_bsp(m.data)
