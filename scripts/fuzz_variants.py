"""One-off exactness fuzz: the kernels with approximations in front of the exact arithmetic (pool / persistent /
thread-per-item: slab pre-test, reject-first brush clip, one-division winner) against the general kernel, which has
none, on millions of TROUBLEMAKER sweeps (axis-aligned, lattice-snapped, starting on brush corners, tiny, zero-length)
plus the bench mix, several hulls.  Any differing byte is a bug."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
import cases
from helpers import get_map
from threecore_b200 import bspgen
from threecore_b200.engine import Engine

dev = torch.device("cuda", 0)
eng = Engine(0)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
bad_total = 0
for kind in ("arena_small", "room", "arena_models"):
    m = get_map(kind)
    eng.load_bsp(m.data)
    for seed in range(3):
        rng = np.random.default_rng(100 + seed)
        es, ee = cases._edge_sweeps(m, rng, 1 << 20)
        ms, me = bspgen.make_sweeps(m, 1 << 20, 200 + seed, len_short=(0.01, 64.0), len_long=(64.0, 3000.0))
        s = np.concatenate([es, ms]); e = np.concatenate([ee, me])
        ds, de = T(s), T(e)
        n = len(s)
        for hull in ((None, None), (bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS), (np.array([-7.3, -0.1, -33.7], np.float32), np.array([12.9, 0.3, 1.1], np.float32)),
                     (np.array([-0.125, -0.125, -0.125], np.float32), np.array([0.125, 0.125, 0.125], np.float32))):
            outs = {}
            for label, opts in (("general", dict(fast=0)), ("pool", dict(fast=1, pool=1, pool_min_items=1, simple_max_items=0)),
                                ("persistent", dict(fast=1, pool=0, simple_max_items=0)), ("simple", dict(fast=1, pool=0, simple_max_items=1 << 30)),
                                ("pool-unbinned", dict(fast=1, pool=1, pool_min_items=1, simple_max_items=0, bin=0)),
                                ("pool-plain", dict(fast=1, pool=1, pool_min_items=1, simple_max_items=0, node_thresholds=0, two_phase=0))):
                for k, v in opts.items(): eng.set_option(k, v)
                o = torch.empty((n, 56), dtype=torch.uint8, device=dev)
                eng.box_trace_device(ds, de, hull[0], hull[1], bspgen.MASK_ALL, o)
                torch.cuda.synchronize(); eng.check_status()
                outs[label] = o
                for k, v in dict(fast=1, pool=1, pool_min_items=196609, simple_max_items=786432, bin=1, node_thresholds=1, two_phase=1).items(): eng.set_option(k, v)
            for label, o in outs.items():
                if label == "general": continue
                bad = int((o != outs["general"]).any(dim=1).sum())
                bad_total += bad
                if bad: print("MISMATCH", kind, seed, None if hull[0] is None else hull[0].tolist(), label, bad, flush=True)
            hits = float((outs["general"][:, 8:12].view(torch.float32) < 1).float().mean())
        print(f"{kind} seed {seed}: {n} sweeps x 4 hulls x 5 variants ok so far (bad={bad_total}), hit fraction {hits:.2f}", flush=True)
print("TOTAL MISMATCHES", bad_total)
