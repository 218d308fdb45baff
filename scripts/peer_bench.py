"""N ranks, each tracing its 16 Mi-ray shard: records kept local vs stored into rank 0's buffer over NVLink (PeerResults).
run: python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 scripts/peer_bench.py"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R)
import numpy as np, torch, torch.distributed as dist
import bench
from threecore_b200 import bspgen
from threecore_b200.engine import Engine
from threecore_b200.sharding import PeerResults

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
per = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 24
m, s, e = bench.make_workload(per, rank)
eng = Engine(local); eng.load_bsp(m.data)
ds, de = torch.from_numpy(s).to(dev), torch.from_numpy(e).to(dev)
out = torch.empty((per, 56), dtype=torch.uint8, device=dev)
pr = PeerResults(eng, per * world)
lo, hi = pr.local_range()
def timed(target, label):
    for _ in range(2):
        eng.box_trace_device(ds, de, bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, target)
    torch.cuda.synchronize(); dist.barrier()
    a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5):
        eng.box_trace_device(ds, de, bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, target)
    b.record(); torch.cuda.synchronize(); dist.barrier()
    t = torch.tensor([a.elapsed_time(b) / 5], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(f"{label:48s}: {t.item():7.2f} ms/step  {world * per / t.item() / 1e3:8.1f} Mtraces/s (all ranks)", flush=True)
timed(out, "records stay on each rank")
for spec in ["remote_pipeline=0"] + sys.argv[2:]:
    for kv in spec.split(","):
        eng.set_option(kv.split("=")[0], float(kv.split("=")[1]))
    timed(pr.pointer(lo), f"records into rank 0 over NVLink [{spec}]")
pr.finish()
if rank == 0:
    rec = pr.records()
    print("rank-0 view equals its own shard:", bool(torch.equal(rec[:per], out)), flush=True)
pr.close(); eng.close(); dist.destroy_process_group()
