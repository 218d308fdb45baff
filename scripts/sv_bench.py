"""Whole SV_Trace on the device: the 20k-brush arena with 256 inline models, 1024 linked entities
(movers, boxes, missiles), 4 Mi player-hull moves resident in HBM, against the reference's SV_Trace on the host.

    python scripts/sv_bench.py [moves] [entities]
"""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
import sv_scene
from threecore_b200 import bspgen
from threecore_b200.engine import Engine, SV_ENTITY_DTYPE, traces_from_tensor

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 22
E = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
m = bspgen.arena_map(n_brushes=20000, seed=0xC0FFEE02, n_models=256)
ents = sv_scene.make_entities(m, E, seed=81)
# movement-like moves, a third of them aimed at entities
s, e = bspgen.make_sweeps(m, n, 0xBADC0DE6)
rng = np.random.default_rng(82)
aim = rng.random(n) < 0.33
tgt = rng.integers(0, E, size=n)
d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
reach = rng.uniform(40, 300, size=(n, 1))
s[aim] = (ents["origin"][tgt] - d * reach)[aim].astype(np.float32)
e[aim] = (ents["origin"][tgt] + d * reach * rng.uniform(0.0, 1.0, size=(n, 1)))[aim].astype(np.float32)
passent = np.where(rng.random(n) < 0.5, ents["number"][rng.integers(0, E, size=n)], sv_scene.ENTITYNUM_NONE).astype(np.int32)
mask = bspgen.MASK_PLAYERSOLID

from oracle import cmref
have_ref = cmref.available()
if have_ref:
    ref = cmref.RefCM().load(m.data)
    sv_scene.link_all_ref(ref, ents)
    order = ref.sv_area_entities([-1e6] * 3, [1e6] * 3)
    bounds = {int(k): ref.sv_entity_bounds(int(k)) for k in ents["number"]}
else:
    raise SystemExit("needs oracle/_ref/libcmref.so (linking is the engine's job; this script uses the reference's)")
# the reference's SV_Trace on every host core: forked workers (the CM and SV layers are global state), each timing
# its own slice of the first `kcpu` moves; throughput = moves / slowest worker.  Done BEFORE CUDA is initialised.
import multiprocessing as mp
import bench as _b
cores = _b.host_cores()
kcpu = min(n, 40000 * cores)
def _work(w):
    lo, hi = w * kcpu // cores, (w + 1) * kcpu // cores
    try:
        os.sched_setaffinity(0, {sorted(os.sched_getaffinity(0))[w % len(os.sched_getaffinity(0))]})
    except OSError:
        pass
    return ref.sv_time_traces(s[lo:hi], e[lo:hi], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, passent[lo:hi], mask)
with mp.get_context("fork").Pool(cores) as pool:
    secs_all = max(pool.map(_work, range(cores)))
secs_one = ref.sv_time_traces(s[:40000], e[:40000], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, passent[:40000], mask)
print(f"reference SV_Trace on the host: {40000 / secs_one / 1e6:.3f} Mmoves/s on one core, {kcpu / secs_all / 1e6:.3f} Mmoves/s on {cores} cores "
      f"(forked workers, first {kcpu} moves)", flush=True)
row = {int(k): i for i, k in enumerate(ents["number"])}
snap = np.zeros(len(order), SV_ENTITY_DTYPE)
for j, num in enumerate(order):
    i = row[int(num)]
    snap[j] = (num, ents["owner"][i], ents["contents"][i], ents["bmodel"][i], ents["modelindex"][i], ents["mins"][i], ents["maxs"][i],
               bounds[int(num)][0], bounds[int(num)][1], ents["origin"][i], ents["angles"][i])
owners = np.full(4096, sv_scene.ENTITYNUM_NONE, np.int32)
owners[ents["number"]] = ents["owner"]
owners[2000] = ents["number"][0]
dev = torch.device("cuda", 0)
eng = Engine(0)
eng.load_bsp(m.data)
eng.sv_set_entities(snap, owners)

T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
ds, de, dp = T(s), T(e), T(passent)
out = torch.empty((n, 56), dtype=torch.uint8, device=dev)
fn = lambda: eng.sv_trace_device(ds, de, bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, dp, mask, out)
for _ in range(2): fn()
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
reps = 5
a.record()
for _ in range(reps): fn()
b.record(); torch.cuda.synchronize(); eng.check_status()
ms = a.elapsed_time(b) / reps
cand = eng.sv_last_candidates
got = traces_from_tensor(out)
hit_ent = (got["entityNum"] < sv_scene.ENTITYNUM_WORLD).mean()
print(f"device SV_Trace: {n} moves, {E} entities, {cand} entity clips ({cand / n:.2f} per move): {ms:.2f} ms  {n / ms / 1e3:.1f} Mmoves/s, "
      f"{hit_ent:.2f} stop at an entity", flush=True)
# world pass alone, for the split
fn2 = lambda: eng.box_trace_device(ds, de, bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, mask, out)
for _ in range(2): fn2()
a.record()
for _ in range(reps): fn2()
b.record(); torch.cuda.synchronize()
print(f"  world pass alone: {a.elapsed_time(b) / reps:.2f} ms", flush=True)
# parity on a sample of the same moves
k = min(n, 200000)
fn()
torch.cuda.synchronize()
got = traces_from_tensor(out)[:k]
want = ref.sv_trace(s[:k], e[:k], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, passent[:k], mask)
same = bool((np.ascontiguousarray(got).view(np.uint8).reshape(k, 56) == np.ascontiguousarray(want).view(np.uint8).reshape(k, 56)).all())
print(f"first {k} records identical to the reference's: {same};  device / {cores}-core reference = {n / ms / 1e3 / (kcpu / secs_all / 1e6):.0f}x", flush=True)
