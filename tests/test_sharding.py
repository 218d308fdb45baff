"""Multi-rank plumbing on CPU: world_size-2 gloo, the oracle standing in for the per-rank GPU engine."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT, bspgen, cmoracle, flat_for, get_map
from threecore_b200 import sharding


def test_shard_ranges_tile_the_batch():
    for n in (0, 1, 2, 7, 64, 65537, 1 << 24):
        for world in (1, 2, 3, 4, 8):
            r = sharding.shard_ranges(n, world)
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
            sizes = [h - l for l, h in r]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        m = get_map("room")
        o = cmoracle.OracleCM(flat_for(m.data))
        s, e = bspgen.make_sweeps(m, n, 21)

        def trace_fn(ts, te):
            out = o.box_trace(ts.numpy(), te.numpy(), bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, mask=bspgen.MASK_PLAYERSOLID)
            return torch.from_numpy(np.ascontiguousarray(out).view(np.uint8).reshape(len(out), 56).copy())

        tr = sharding.ShardedTracer(trace_fn)
        full = tr.trace(torch.from_numpy(s), torch.from_numpy(e), gather=True)
        local = tr.trace(torch.from_numpy(s), torch.from_numpy(e))
        lo, hi = tr.local_range(n)
        ok_local = torch.equal(local, full[lo:hi])
        # max-over-ranks reduction used by bench.py
        t = torch.tensor([float(rank + 1)], dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            want = trace_fn(torch.from_numpy(s), torch.from_numpy(e))
            q.put((bool(torch.equal(full, want)), ok_local, float(t.item())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [1001, 4096])
def test_two_rank_gather_equals_single_process(n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n % 7
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    same, ok_local, tmax = q.get(timeout=10)
    assert same and ok_local and tmax == 2.0
