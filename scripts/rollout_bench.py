"""BASELINE config 5 numbers: 65,536 agents x 10 traces/tick on a 50k-brush arena.

    python scripts/rollout_bench.py [agents] [ticks]                                   one GPU
    torchrun --nproc-per-node N ... scripts/rollout_bench.py [agents] [ticks] [weak] [graph]   agents sharded by id over N GPUs;
                                                                                       graph: the tick replayed as a CUDA graph
                                                                                       (weak: `agents` per GPU)
The BSP is replicated, every rank steps its own agents (Rollout(first_agent=...)), there is no exchange inside a tick;
the time of a run is the slowest rank's (CUDA events, max over ranks)."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch
from threecore_b200 import bspgen
from threecore_b200.engine import Engine
from threecore_b200.rollout import Rollout, TRACES_PER_TICK
from threecore_b200.sharding import shard_range

agents = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
ticks = int(sys.argv[2]) if len(sys.argv) > 2 else 100
weak = "weak" in sys.argv[3:]
graph = "graph" in sys.argv[3:]
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
if world > 1:
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
total = agents * world if weak else agents
lo, hi = shard_range(total, rank, world)
t0 = time.time()
m = bspgen.arena_map(n_brushes=50000, seed=0xC0FFEE05)
if rank == 0:
    print(f"map: 50k brushes generated in {time.time()-t0:.1f} s", flush=True)
eng = Engine(local); eng.load_bsp(m.data)
r = Rollout(eng, hi - lo, m.world_mins, m.world_maxs, seed=5, first_agent=lo)
r.run(5); torch.cuda.synchronize()
if graph:
    r.capture()
if world > 1:
    dist.barrier()
a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
tr0, l0 = r.traces, r.launches
a.record(); r.run(ticks); b.record(); torch.cuda.synchronize()
eng.check_status()
ms = torch.tensor([a.elapsed_time(b)], dtype=torch.float64, device=f"cuda:{local}")
n = torch.tensor([float(r.traces - tr0)], dtype=torch.float64, device=f"cuda:{local}")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
ms, n = float(ms.item()), float(n.item())
if rank == 0:
    print(f"agents {total} on {world} GPU(s) ({'weak' if weak else 'strong'} scaling{', CUDA graph' if graph else ''}), ticks {ticks}: {ms/ticks:.3f} ms/tick, "
          f"{n/ms/1e3:.1f} Mtraces/s, {(r.launches-l0)/ticks:.0f} launches/tick ({TRACES_PER_TICK} traces per agent and tick)", flush=True)
if world > 1:
    dist.destroy_process_group()
