"""Experiment: resident kernel time of the bench workload for binning knobs.  usage: exp_bin.py [rays] [name=value ...]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
import bench
from threecore_b200 import bspgen
from threecore_b200.engine import Engine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
m, starts, ends = bench.make_workload(n)
dev = torch.device("cuda", 0)
ds = torch.from_numpy(starts).to(dev); de = torch.from_numpy(ends).to(dev)
d_out = torch.empty((n, 56), dtype=torch.uint8, device=dev)
ref = None
def run(label, **opts):
    global ref
    eng = Engine(0)
    for k, v in opts.items(): eng.set_option(k, v)
    eng.load_bsp(m.data)
    for _ in range(2): eng.box_trace_device(ds, de, bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, d_out)
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): eng.box_trace_device(ds, de, bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, d_out)
    b.record(); torch.cuda.synchronize(); eng.check_status()
    ms = a.elapsed_time(b) / 3
    if ref is None: ref = d_out.clone()
    same = bool(torch.equal(ref, d_out))
    print(f"{label:40s}: {ms:7.2f} ms  {n/ms/1e3:8.1f} Mtraces/s  same={same}", flush=True)
    eng.close()
sets = [a for a in sys.argv[2:]]
if sets:
    for spec in sets:
        opts = {kv.split('=')[0]: float(kv.split('=')[1]) for kv in spec.split(',')}
        run(spec, **opts)
else:
    run("bin off", bin=0)
    for cell in (512, 256, 128, 64, 32):
        run(f"bin cell {cell}", bin_cell=cell)
    for ll in (128, 256, 1024, 1e9):
        run(f"bin cell 128 long {ll}", bin_cell=128, bin_long_len=ll)
    for g in (2, 8, 16):
        run(f"bin cell 128 gang {g}", bin_cell=128, gang=g)
    for lr in (1, 2, 5, 8):
        run(f"bin cell 128 light_reps {lr}", bin_cell=128, light_reps=lr)
