"""Experiment: end-to-end (host pointers, pinned) rate of the bench workload vs pipeline depth; raw PCIe copy rates."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
import bench
from threecore_b200 import bspgen
from threecore_b200.engine import Engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
m, starts, ends = bench.make_workload(n)
dev = torch.device("cuda", 0)
h_s = torch.from_numpy(starts).pin_memory(); h_e = torch.from_numpy(ends).pin_memory()
h_out = torch.empty((n, 56), dtype=torch.uint8).pin_memory()
d_out = torch.empty((n, 56), dtype=torch.uint8, device=dev)
d_s = torch.empty((n, 3), dtype=torch.float32, device=dev)
for name, fn, nbytes in (("D2H 896MiB", lambda: h_out.copy_(d_out, non_blocking=True), n * 56),
                         ("H2D 192MiB", lambda: d_s.copy_(h_s, non_blocking=True), n * 12)):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(3): fn()
    torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 3
    print(f"{name}: {dt*1e3:.2f} ms  {nbytes/dt/1e9:.1f} GB/s", flush=True)
# both directions at once
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
torch.cuda.synchronize(); t = time.perf_counter()
for _ in range(3):
    with torch.cuda.stream(s1): h_out.copy_(d_out, non_blocking=True)
    with torch.cuda.stream(s2): d_s.copy_(h_s, non_blocking=True); d_s.copy_(h_e, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t) / 3
print(f"D2H 896MiB + H2D 384MiB together: {dt*1e3:.2f} ms", flush=True)
eng = Engine(0); eng.load_bsp(m.data)
for chunks in (4, 8, 16, 32, 64):
    eng.set_option("pipe_chunks", chunks)
    for _ in range(2):
        eng.box_trace_hostptr(n, h_s.data_ptr(), h_e.data_ptr(), bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, h_out.data_ptr())
    t = time.perf_counter()
    for _ in range(3):
        eng.box_trace_hostptr(n, h_s.data_ptr(), h_e.data_ptr(), bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, h_out.data_ptr())
    dt = (time.perf_counter() - t) / 3
    print(f"pipe_chunks {chunks:3d}: {dt*1e3:7.2f} ms  {n/dt/1e6:8.1f} Mtraces/s", flush=True)
# small-batch latency, rays resident on the device
ds = torch.from_numpy(starts).to(dev); de = torch.from_numpy(ends).to(dev)
for k in (4096, 32768, 65536, 655360, 1 << 21):
    o = d_out[:k]
    for _ in range(3): eng.box_trace_device(ds[:k], de[:k], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, o)
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): eng.box_trace_device(ds[:k], de[:k], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, o)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"resident n={k:8d}: {ms*1e3:8.1f} us  {k/ms/1e3:8.1f} Mtraces/s", flush=True)
