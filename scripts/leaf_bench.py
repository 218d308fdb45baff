"""Leaf / cluster / visibility queries on the device against the reference's cm_test.c and the
snapshot PVS lines on one host core: the 20k-brush arena with clusters, 5 areas and a visibility lump.

    python scripts/leaf_bench.py [points] [boxes] [views] [entities]
"""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
import leaf_cases as lc
from threecore_b200 import bspgen
from threecore_b200.engine import Engine, ENT_CLUSTERS_DTYPE

nP = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
nB = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 20
nV = int(sys.argv[3]) if len(sys.argv) > 3 else 1024
nE = int(sys.argv[4]) if len(sys.argv) > 4 else 4096
m = bspgen.with_visibility(bspgen.arena_map(n_brushes=20000, seed=0xC0FFEE02), seed=0x715, n_areas=5)
print(f"map: {m.stats['leafs']} leafs, {m.stats['clusters']} clusters, {m.stats['areas']} areas, "
      f"visibility {m.stats['clusters'] * (((m.stats['clusters'] + 63) & ~63) >> 3) / 1e6:.1f} MB", flush=True)
pts = lc.viewpoints(m, nP, 51)
# entity-like boxes: SV_LinkEntity's clientele (players, items, movers up to a few hundred units)
rng = np.random.default_rng(52)
c = m.world_mins + rng.random((nB, 3)).astype(np.float32) * (m.world_maxs - m.world_mins)
half = np.where(rng.random((nB, 1)) < 0.8, rng.uniform(8, 40, (nB, 3)), rng.uniform(100, 600, (nB, 3))).astype(np.float32)
mins, maxs = c - half, c + half
views = lc.viewpoints(m, nV, 53)

from oracle import cmref
ref = cmref.RefCM().load(m.data) if cmref.available() else None
if ref is not None:
    kp, kb = min(nP, 2_000_000), min(nB, 200_000)
    t = time.perf_counter(); want_leafs = ref.point_leafnum(pts[:kp]); tp = time.perf_counter() - t
    t = time.perf_counter(); want_lists = ref.box_leafnums(mins[:kb], maxs[:kb], 128); tb = time.perf_counter() - t
    print(f"reference, one host core: CM_PointLeafnum {kp / tp / 1e6:.2f} M/s, CM_BoxLeafnums(128) {kb / tb / 1e6:.3f} M/s", flush=True)

dev = torch.device("cuda", 0)
eng = Engine(0)
eng.upload(eng.load_bsp(m.data))
eng.adjust_area_portal_state(0, 1, True)
eng.adjust_area_portal_state(1, 3, True)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

def timed(fn, reps=10):
    for _ in range(3): fn()
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize(); eng.check_status()
    return a.elapsed_time(b) / reps

dp, dl = T(pts), torch.empty(nP, dtype=torch.int32, device=dev)
ms = timed(lambda: eng.point_leafnum_device(dp, dl))
print(f"device CM_PointLeafnum: {nP} points {ms:.3f} ms  {nP / ms / 1e3:.0f} M/s  ({nP * 16 / ms / 1e6:.0f} GB/s of 12 B in + 4 B out)", flush=True)
dmn, dmx = T(mins), T(maxs)
lists = torch.empty((nB, 128), dtype=torch.int32, device=dev)
cnt, last = torch.empty(nB, dtype=torch.int32, device=dev), torch.empty(nB, dtype=torch.int32, device=dev)
ms = timed(lambda: eng.box_leafnums_device(dmn, dmx, 128, lists, cnt, last))
print(f"device CM_BoxLeafnums(128): {nB} boxes {ms:.3f} ms  {nB / ms / 1e3:.1f} M/s, {cnt.float().mean().item():.1f} leafs per box", flush=True)
recs = torch.empty((nB, 21), dtype=torch.int32, device=dev)
ms = timed(lambda: eng.entity_clusters_device(dmn, dmx, recs))
print(f"device SV_LinkEntity clusters/areas: {nB} boxes {ms:.3f} ms  {nB / ms / 1e3:.1f} M/s  ({nB * 108 / ms / 1e6:.0f} GB/s of 24 B in + 84 B out)", flush=True)
dv = T(views)
words = (nE + 31) // 32
vis = torch.empty((nV, words), dtype=torch.int32, device=dev)
ents = recs[:nE].contiguous()
ms = timed(lambda: eng.visible_from_points_device(dv, nE, ents, vis))
frac = np.unpackbits(vis.cpu().numpy().view(np.uint8), bitorder="little").mean()
print(f"device snapshot PVS stage: {nV} viewpoints x {nE} entities {ms:.3f} ms  {nV * nE / ms / 1e3:.0f} Mpairs/s, {frac:.2f} visible", flush=True)
if ref is not None:
    assert np.array_equal(dl.cpu().numpy()[:kp], want_leafs)
    g = (lists.cpu().numpy()[:kb], cnt.cpu().numpy()[:kb], last.cpu().numpy()[:kb])
    ok = np.array_equal(g[1], want_lists[1]) and np.array_equal(g[2], want_lists[2])
    idx = np.arange(128)[None, :] < g[1][:, None]
    ok = ok and np.array_equal(g[0][idx], want_lists[0][idx])
    print(f"first {kp} leaf numbers and {kb} leaf lists identical to the reference's: {ok}", flush=True)
