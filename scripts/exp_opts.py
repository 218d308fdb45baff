"""Step time of the bench workload under a list of option settings: python scripts/exp_opts.py "gang_brush=6" "gang_brush=8,gang=3" ..."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
import bench
from threecore_b200 import bspgen
from threecore_b200.engine import Engine

n = 1 << 24
m, starts, ends = bench.make_workload(n)
eng = Engine(0)
eng.load_bsp(m.data)
dev = torch.device("cuda", 0)
ds, de = torch.from_numpy(starts).to(dev), torch.from_numpy(ends).to(dev)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
out = torch.zeros((n, 56), dtype=torch.uint8, device=dev)
ref = None
defaults = {}
for spec in ["base"] + sys.argv[1:] + ["base"]:
    opts = {} if spec == "base" else {kv.split("=")[0]: float(kv.split("=")[1]) for kv in spec.split(",")}
    for k, v in opts.items():
        eng.set_option(k, v)
    if "entry_cell" in opts or "l2_window" in opts:      # options that take effect at upload
        eng.load_bsp(m.data)
    fn = lambda: eng.box_trace_device(ds, de, bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, out)
    for _ in range(2): fn()
    eng.kernel_timing(True)
    a.record()
    for _ in range(5): fn()
    b.record(); torch.cuda.synchronize(); eng.check_status()
    kms, kn = eng.kernel_time_ms()
    eng.kernel_timing(False)
    ms = a.elapsed_time(b) / 5
    same = "" if ref is None else f" same={bool(torch.equal(out, ref))}"
    if ref is None: ref = out.clone()
    chk = int(out.view(torch.int64).sum().item()) & 0xffffffffffff
    print(f"{spec:40s} step {ms:.3f} ms  {n / ms / 1e3:.0f} Mtraces/s   kernels {kms / kn:.3f} ms{same} checksum {chk:012x}", flush=True)
    for k in opts:                              # back to the defaults given on the command line as D:k=v
        eng.set_option(k, float(os.environ.get("D_" + k, "nan"))) if ("D_" + k) in os.environ else None
