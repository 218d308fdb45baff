"""Resident throughput of the other BASELINE configs (1, 3, 4) -- parity cases, not bench lines; numbers for DESIGN.md."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
from threecore_b200 import bspgen
from threecore_b200.engine import Engine, rotation_matrices

dev = torch.device("cuda", 0)
eng = Engine(0)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
def timed(label, n, fn, reps=5):
    for _ in range(2): fn()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize(); eng.check_status()
    ms = a.elapsed_time(b) / reps
    print(f"{label:58s}: {ms:8.2f} ms  {n/ms/1e3:8.1f} M/s", flush=True)

# config 1
m = bspgen.room_map(n_brushes=500, seed=0xC0FFEE01); eng.load_bsp(m.data)
n = 1 << 20
s, e = bspgen.make_point_sweeps(m, n, 0xBADC0DE1)
ds, de = T(s), T(e); out = torch.empty((n, 56), dtype=torch.uint8, device=dev)
timed("config 1: 1Mi point traces, 500-brush room", n, lambda: eng.box_trace_device(ds, de, None, None, bspgen.MASK_SOLID, out))
# config 3
m = bspgen.patch_map(n_brushes=2000, n_patches=512, seed=0xC0FFEE03); eng.load_bsp(m.data)
n = 1 << 22
s, e, mins, maxs = bspgen.make_patch_sweeps(m, n, 0xBADC0DE3)
ds, de, dmn, dmx = T(s), T(e), T(mins), T(maxs); out = torch.empty((n, 56), dtype=torch.uint8, device=dev)
timed("config 3: 4Mi point+hull traces, 2k brushes + 512 patches", n, lambda: eng.box_trace_device(ds, de, dmn, dmx, bspgen.MASK_ALL, out))
# config 4
m = bspgen.arena_map(n_brushes=20000, seed=0xC0FFEE02, n_models=256); eng.load_bsp(m.data)
n = 1 << 22
t = bspgen.make_entity_traces(m, n, 0xBADC0DE4)
rot, ridx = rotation_matrices(t["angles"])
d = {k: T(v) for k, v in t.items() if k != "angles"}
drot, dridx = T(rot), T(ridx)
timed("config 4: 4Mi transformed traces, 256 inline models + temp boxes", n,
      lambda: eng.box_trace_device(d["starts"], d["ends"], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_ALL, out,
                                   models=d["models"], origins=d["origins"], rotations=drot, rot_index=dridx,
                                   boxmins=d["boxmins"], boxmaxs=d["boxmaxs"]))
npts = 1 << 24
pts = T(bspgen.make_points(m, npts, 0xBADC0DE5)); pc = torch.empty(npts, dtype=torch.int32, device=dev)
timed("config 4: 16Mi CM_PointContents (world)", npts, lambda: eng.point_contents_device(pts, pc))
pc2 = torch.empty(n, dtype=torch.int32, device=dev)
timed("config 4: 4Mi CM_TransformedPointContents", n,
      lambda: eng.point_contents_device(d["starts"], pc2, models=d["models"], origins=d["origins"], rotations=drot, rot_index=dridx,
                                        boxmins=d["boxmins"], boxmaxs=d["boxmaxs"]))

# ---- the reference's CPU code on the same inputs (oracle/_ref, one forked worker per host core; shared hulls only,
# so config 3 is timed as its point half and its hull half) -- a reported baseline like bench.py's cpu_baseline
try:
    from oracle import cmref
    import bench as _b
    cores = _b.host_cores()
    if cmref.available():
        def cpu(label, data, s, e, mins, maxs, mask, k=200000):
            r = cmref.RefCM().load(data)
            secs, _, _ = r.bench_box_trace(s[:k], e[:k], mins, maxs, mask, nworkers=cores, repeats=1)
            print(f"reference, {cores} cores, {label:44s}: {min(k, len(s))/secs/1e6:8.2f} M/s", flush=True)
        m1 = bspgen.room_map(n_brushes=500, seed=0xC0FFEE01)
        s, e = bspgen.make_point_sweeps(m1, 1 << 20, 0xBADC0DE1)
        cpu("config 1 point traces", m1.data, s, e, None, None, bspgen.MASK_SOLID, k=1 << 20)
        m3 = bspgen.patch_map(n_brushes=2000, n_patches=512, seed=0xC0FFEE03)
        s, e, mins, maxs = bspgen.make_patch_sweeps(m3, 1 << 20, 0xBADC0DE3)
        pt = (mins == 0).all(axis=1)
        cpu("config 3 point half", m3.data, s[pt], e[pt], None, None, bspgen.MASK_ALL, k=400000)
        cpu("config 3 player-hull half", m3.data, s[~pt], e[~pt], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_ALL, k=400000)
except Exception as ex:      # the reference library is optional on the GPU box
    print("reference timing skipped:", ex)
