"""Throughput of the batched AAS tracer (AAS_TraceClientBBox) on a generated AAS world, lines resident in HBM; the
unmodified reference on the host cores beside it."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
from threecore_b200 import aasgen
from threecore_b200.engine import Engine
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
w = aasgen.make_world(n_boxes=2000, seed=0xAA5004, extent=8192.0, height=1024.0)
print(w.stats, flush=True)
eng = Engine(0); eng.aas_upload(w)
s, e = aasgen.make_lines(w, n, 11)
dev = torch.device("cuda", 0)
ds, de = torch.from_numpy(s).to(dev), torch.from_numpy(e).to(dev)
out = torch.empty((n, 36), dtype=torch.uint8, device=dev)
for _ in range(2): eng.aas_trace_client_bbox_device(ds, de, aasgen.PRESENCE_NORMAL, out)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): eng.aas_trace_client_bbox_device(ds, de, aasgen.PRESENCE_NORMAL, out)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 5
print(f"AAS_TraceClientBBox x {n}: {ms:.3f} ms  {n / ms / 1e3:.0f} Mtraces/s  ({n * 60 / ms / 1e6:.0f} GB/s of 24 B in + 36 B out)", flush=True)
try:
    from oracle import cmref
    if cmref.available():
        r = cmref.RefCM().aas_load(w)
        k = 1 << 21
        cores = len(os.sched_getaffinity(0))
        t0 = time.perf_counter()
        want = cmref.fork_map(k, cmref.AAS_TRACE_DTYPE, lambda lo, hi: r.aas_trace_client_bbox(s[lo:hi], e[lo:hi], aasgen.PRESENCE_NORMAL))
        secs = time.perf_counter() - t0
        got = out[:k].cpu().numpy().view(cmref.AAS_TRACE_DTYPE).reshape(-1)
        print(f"reference, {cores} cores: {k / secs / 1e6:.1f} Mtraces/s; records equal: {got.tobytes() == want.tobytes()}", flush=True)
except Exception as ex:
    print("reference timing skipped:", ex)

# ---- AAS_PredictClientMovement: predictions resident in HBM, the reference on the host cores beside it
from threecore_b200 import bspgen
m = bspgen.arena_map(n_brushes=20000, seed=0xC0FFEE02, size_xy=8192.0, size_z=1024.0)
eng.load_bsp(m.data)
nm = 1 << 21
p = aasgen.make_moves(w, nm, 13)
dp = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in p.items()}
mo = torch.empty((nm, 84), dtype=torch.uint8, device=dev)
for _ in range(2): eng.aas_predict_client_movement_device(dp, mo)
a.record()
for _ in range(3): eng.aas_predict_client_movement_device(dp, mo)
b.record(); torch.cuda.synchronize()
ms = a.elapsed_time(b) / 3
frames = float(mo.cpu().numpy().view(np.int32).reshape(-1, 21)[:, 20].mean())
print(f"AAS_PredictClientMovement x {nm}: {ms:.2f} ms  {nm / ms / 1e3:.1f} M predictions/s ({frames:.1f} frames each on average)", flush=True)
try:
    if cmref.available():
        r = cmref.RefCM().load(m.data); r.aas_load(w)
        k = 1 << 18
        t0 = time.perf_counter()
        want = cmref.fork_map(k, cmref.AAS_MOVE_DTYPE, lambda lo, hi: r.aas_predict_client_movement({kk: v[lo:hi] for kk, v in p.items()})[0])
        secs = time.perf_counter() - t0
        got = mo[:k].cpu().numpy().view(cmref.AAS_MOVE_DTYPE).reshape(-1)
        print(f"reference, {cores} cores: {k / secs / 1e6:.2f} M predictions/s; records equal: {got.tobytes() == want.tobytes()}", flush=True)
except Exception as ex:
    print("reference timing skipped:", ex)
