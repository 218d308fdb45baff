"""End-to-end (host pointers, pinned) call time of the bench workload under option settings, plus raw PCIe copy rates."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
import bench
from threecore_b200 import bspgen
from threecore_b200.engine import Engine
n = 1 << 24
m, starts, ends = bench.make_workload(n)
h_s = torch.from_numpy(starts).pin_memory(); h_e = torch.from_numpy(ends).pin_memory()
h_out = torch.empty((n, 56), dtype=torch.uint8).pin_memory()
d = torch.empty((n, 56), dtype=torch.uint8, device="cuda")
for name, fn in (("D2H 896 MiB", lambda: h_out.copy_(d, non_blocking=True)), ("H2D 896 MiB", lambda: d.copy_(h_out, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(3): fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 3
    print(f"{name}: {dt*1e3:.2f} ms  {n*56/dt/1e9:.1f} GB/s", flush=True)
eng = Engine(0); eng.load_bsp(m.data)
ref = None
for spec in ["base"] + sys.argv[1:] + ["base"]:
    if spec != "base":
        for kv in spec.split(","):
            eng.set_option(kv.split("=")[0], float(kv.split("=")[1]))
    call = lambda: eng.box_trace_hostptr(n, h_s.data_ptr(), h_e.data_ptr(), bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, h_out.data_ptr())
    for _ in range(2): call()
    t = time.perf_counter()
    for _ in range(5): call()
    dt = (time.perf_counter() - t) / 5
    if ref is None: ref = h_out.clone()
    print(f"{spec:40s} {dt*1e3:7.2f} ms  {n/dt/1e6:8.1f} Mtraces/s same={bool(torch.equal(ref, h_out))}", flush=True)
