"""The host-pointer call under settings of the result split: python scripts/exp_hybrid.py "host_records=-1,host_threads=8" ...
Every setting's records are compared, all 16 Mi x 56 bytes, with those of host_records=0 (every record written by the device)."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
import bench
from threecore_b200 import bspgen
from threecore_b200.engine import Engine
n = int(os.environ.get("N", 1 << 24))
m, starts, ends = bench.make_workload(n)
eng = Engine(0); eng.load_bsp(m.data)
pageable = os.environ.get("PAGEABLE", "0") == "1"
h_s = torch.from_numpy(starts); h_e = torch.from_numpy(ends); h_out = torch.empty((n, 56), dtype=torch.uint8)
if not pageable:
    h_s = h_s.pin_memory(); h_e = h_e.pin_memory(); h_out = h_out.pin_memory()
ref = None
for spec in ["host_records=0"] + sys.argv[1:]:
    opts = {kv.split("=")[0]: float(kv.split("=")[1]) for kv in spec.split(",")}
    for k, v in opts.items(): eng.set_option(k, v)
    f = lambda: eng.box_trace_hostptr(n, h_s.data_ptr(), h_e.data_ptr(), bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, h_out.data_ptr())
    h_out.zero_()
    for _ in range(10): f()
    times = []
    for _ in range(6):
        torch.cuda.synchronize(); t0 = time.perf_counter(); f(); times.append(time.perf_counter() - t0)
    dt = float(np.median(times))
    if ref is None: ref = h_out.clone()
    print(f"{spec:60s} {dt*1e3:7.2f} ms (min {min(times)*1e3:6.2f})  {n/dt/1e6:7.1f} Mtraces/s  split host/device={eng.last_batch_split()}  same={bool(torch.equal(h_out, ref))}", flush=True)
