"""Multi-GPU result gather fused into the trace kernels' stores (sharding.PeerResults, cmgpu_ipc_*): 2, 4 and 8 ranks,
each case skipped where the box has fewer GPUs (`gpurun --gpus N -- python -m pytest tests/test_peer_results.py -m gpu`)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import ROOT, bspgen

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from threecore_b200.engine import Engine
        from threecore_b200.sharding import PeerResults
        m = bspgen.arena_map(n_brushes=2500, seed=0xC0FFEE12, size_xy=4096.0, size_z=2048.0)
        s, e = bspgen.make_sweeps(m, n, 77)
        eng = Engine(rank)
        eng.load_bsp(m.data)
        eng.set_option("pipe_pool_min_items", 100000)      # the rank's range is cut into pieces whose expansions stream to rank 0 (tracePieces)
        dev = torch.device("cuda", rank)
        pr = PeerResults(eng, n)
        lo, hi = pr.local_range()
        ds, de = torch.from_numpy(s[lo:hi]).to(dev), torch.from_numpy(e[lo:hi]).to(dev)
        eng.box_trace_device(ds, de, bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, pr.pointer(lo))
        pr.finish()
        eng.check_status()
        if rank == 0:
            got = pr.records().cpu().numpy()
            # the same batch traced by this rank alone
            one = torch.empty((n, 56), dtype=torch.uint8, device=dev)
            eng.box_trace_device(torch.from_numpy(s).to(dev), torch.from_numpy(e).to(dev), bspgen.PLAYER_MINS,
                                 bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, one)
            torch.cuda.synchronize()
            q.put(bool(np.array_equal(got, one.cpu().numpy())))
        pr.close()
        eng.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_peer_results(world):
    """world ranks (one GPU each) trace their ranges; rank 0's buffer must hold exactly the records of a single-rank run"""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, this box has {torch.cuda.device_count()}")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n = 1_500_001
    procs = [ctx.Process(target=_worker, args=(r, world, 29641 + world, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True
