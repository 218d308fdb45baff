import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
from threecore_b200 import bspgen
from threecore_b200.engine import Engine
m = bspgen.arena_map(n_brushes=20000, seed=0xC0FFEE02)
eng = Engine(0); eng.load_bsp(m.data)
s, e = bspgen.make_sweeps(m, 65536, 5, z_range=(24.0, 2048.0))
mn = np.repeat(np.asarray(bspgen.PLAYER_MINS, np.float32)[None], 65536, 0); mx = np.repeat(np.asarray(bspgen.PLAYER_MAXS, np.float32)[None], 65536, 0)
ref = {}
for mc in (8, 4, 2, 1):
    eng.set_option("min_chunk", mc)
    for n in (16, 128, 1024, 8192, 65536):
        f = lambda: eng.box_trace(s[:n], e[:n], mn[:n], mx[:n], mask=bspgen.MASK_PLAYERSOLID)
        for _ in range(10): r = f()
        if n not in ref: ref[n] = r.copy()
        same = bool((np.ascontiguousarray(r).view(np.uint8) == np.ascontiguousarray(ref[n]).view(np.uint8)).all())
        t0 = time.perf_counter()
        for _ in range(100): f()
        us = (time.perf_counter() - t0) / 100 * 1e6
        print(f"min_chunk={mc} n={n:6d} general kernel: {us:7.1f} us per call same={same}", flush=True)
