"""BASELINE config 5 numbers: 65,536 agents x 10 traces/tick on a 50k-brush arena (one GPU, or one rank's share)."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
from threecore_b200 import bspgen
from threecore_b200.engine import Engine
from threecore_b200.rollout import Rollout, TRACES_PER_TICK

agents = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 100
t0 = time.time()
m = bspgen.arena_map(n_brushes=50000, seed=0xC0FFEE05)
print(f"map: 50k brushes generated in {time.time()-t0:.1f} s", flush=True)
eng = Engine(0); eng.load_bsp(m.data)
r = Rollout(eng, agents, m.world_mins, m.world_maxs, seed=5)
r.run(5); torch.cuda.synchronize()
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
tr0, l0 = r.traces, r.launches
a.record(); r.run(ticks); b.record(); torch.cuda.synchronize()
eng.check_status()
ms = a.elapsed_time(b)
n = r.traces - tr0
print(f"agents {agents} ticks {ticks}: {ms/ticks:.3f} ms/tick, {n/ms/1e3:.1f} Mtraces/s, {(r.launches-l0)/ticks:.0f} launches/tick "
      f"({TRACES_PER_TICK} traces per agent and tick)", flush=True)
