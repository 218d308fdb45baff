"""A/B of the two-phase result path (16-byte results by slot + expandRecordsKernel) against direct 56-byte record
stores, on the bench workload: step time, trace-kernel time, and byte equality of all records."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
import bench
from threecore_b200 import bspgen
from threecore_b200.engine import Engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
m, starts, ends = bench.make_workload(n)
eng = Engine(0)
eng.load_bsp(m.data)
dev = torch.device("cuda", 0)
ds, de = torch.from_numpy(starts).to(dev), torch.from_numpy(ends).to(dev)
outs = {}
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for name, opts in (("direct", {"two_phase": 0}), ("two-phase", {"two_phase": 1}), ("direct", {"two_phase": 0}), ("two-phase", {"two_phase": 1})):
    for k, v in opts.items():
        eng.set_option(k, v)
    out = torch.zeros((n, 56), dtype=torch.uint8, device=dev)
    fn = lambda: eng.box_trace_device(ds, de, bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, out)
    for _ in range(3): fn()
    eng.kernel_timing(True)
    a.record()
    for _ in range(10): fn()
    b.record(); torch.cuda.synchronize(); eng.check_status()
    kms = eng.kernel_time_ms()
    eng.kernel_timing(False)
    ms = a.elapsed_time(b) / 10
    print(f"{name:10s} step {ms:.3f} ms  {n / ms / 1e3:.0f} Mtraces/s   trace kernel {kms}", flush=True)
    outs[name] = out
print("records identical:", bool(torch.equal(outs["direct"], outs["two-phase"])))
