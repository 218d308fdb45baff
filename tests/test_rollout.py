"""BASELINE config 5: the agent rollout (threecore_b200/rollout.py).  Its kinematics are ours; parity is pinned at the
CM_BoxTrace boundary by replaying every batch the rollout issued through the oracle."""
import numpy as np
import pytest

from helpers import assert_traces_equal, bspgen, cmoracle, cmref, get_map

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from threecore_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def test_rollout_batches_replay_through_the_oracle(eng):
    from threecore_b200.rollout import Rollout, TRACES_PER_TICK
    m = get_map("arena_small")
    fm = eng.load_bsp(m.data)
    o = cmoracle.OracleCM(fm.to_dict())
    batches = []
    n, ticks = 3000, 6
    r = Rollout(eng, n, m.world_mins, m.world_maxs, seed=9, record=lambda s, e, rec: batches.append((s, e, rec)))
    r.run(ticks)
    assert r.traces == n * ticks * TRACES_PER_TICK
    hits = moved = 0
    for i, (s, e, rec) in enumerate(batches):
        want = o.box_trace(s, e, bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, mask=bspgen.MASK_PLAYERSOLID)
        got = np.ascontiguousarray(rec).view(cmref.TRACE_DTYPE).reshape(-1)
        assert_traces_equal(got, want, f"rollout batch {i}")
        hits += int((got["fraction"] < 1).sum())
        moved += int((s != e).any(axis=1).sum())
    assert hits > 0.05 * r.traces and moved > 0.5 * r.traces, "the rollout is not exercising collisions"


def test_rollout_is_independent_of_the_sharding(eng):
    """agents are independent: tracing [0,n) on one rank or as two ranges gives the same agents"""
    import torch
    from threecore_b200.rollout import Rollout
    from threecore_b200.sharding import shard_range
    m = get_map("arena_small")
    eng.load_bsp(m.data)
    n, ticks = 4096, 4
    whole = Rollout(eng, n, m.world_mins, m.world_maxs, seed=3).run(ticks)
    parts = []
    for rank in range(2):
        lo, hi = shard_range(n, rank, 2)
        parts.append(Rollout(eng, hi - lo, m.world_mins, m.world_maxs, seed=3, first_agent=lo).run(ticks))
    torch.cuda.synchronize()
    assert torch.equal(whole.origin, torch.cat([p.origin for p in parts]))
    assert torch.equal(whole.velocity, torch.cat([p.velocity for p in parts]))


@pytest.mark.parametrize("n", [4096, 40000])
def test_captured_tick_replays_like_the_eager_one(eng, n):
    """a tick recorded as a CUDA graph (the device-pointer calls only enqueue work, so they can be captured) moves the
    agents exactly like ticks issued launch by launch; 40000 agents make the first phase a binned launch (120000 traces)"""
    import torch
    from threecore_b200.rollout import Rollout
    m = get_map("arena_small")
    eng.load_bsp(m.data)
    eager = Rollout(eng, n, m.world_mins, m.world_maxs, seed=4).run(7)
    graph = Rollout(eng, n, m.world_mins, m.world_maxs, seed=4).capture(warmup=3).run(4)
    torch.cuda.synchronize()
    eng.check_status()
    assert graph.traces == eager.traces and graph.launches == eager.launches
    assert torch.equal(graph.origin, eager.origin)
    assert torch.equal(graph.velocity, eager.velocity)
    # eager calls still work on the same context afterwards (the capture left no event or workspace state behind)
    again = Rollout(eng, n, m.world_mins, m.world_maxs, seed=4).run(7)
    torch.cuda.synchronize()
    assert torch.equal(again.origin, eager.origin)
