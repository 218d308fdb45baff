"""Latency of small host-pointer calls (the drop-in CM_BoxTrace is a batch of one): mapped-memory path on / off."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np
from threecore_b200 import bspgen
from threecore_b200.engine import Engine
m = bspgen.arena_map(n_brushes=20000, seed=0xC0FFEE02)
eng = Engine(0); eng.load_bsp(m.data)
s, e = bspgen.make_sweeps(m, 4096, 5, z_range=(24.0, 2048.0))
mn = np.repeat(np.asarray(bspgen.PLAYER_MINS, np.float32)[None], 4096, 0); mx = np.repeat(np.asarray(bspgen.PLAYER_MAXS, np.float32)[None], 4096, 0)
for tiny in (0, 1):
    eng.set_option("tiny_calls", tiny)
    for n in (1, 16, 128, 129, 1024):
        for per_item in (False, True):
            f = (lambda: eng.box_trace(s[:n], e[:n], mn[:n], mx[:n], mask=bspgen.MASK_PLAYERSOLID)) if per_item else \
                (lambda: eng.box_trace(s[:n], e[:n], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, mask=bspgen.MASK_PLAYERSOLID))
            for _ in range(20): f()
            t0 = time.perf_counter()
            for _ in range(300): f()
            us = (time.perf_counter() - t0) / 300 * 1e6
            print(f"tiny_calls={tiny} n={n:5d} {'per-item hulls (general kernel)' if per_item else 'shared hull (hot path)        '}: {us:7.1f} us per call", flush=True)
