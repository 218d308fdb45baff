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
