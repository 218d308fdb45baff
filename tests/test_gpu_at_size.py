"""BASELINE.json configs 2-5 at their FULL sizes, the CUDA records against the UNMODIFIED reference (libcmref.so) on
the same map bytes and the same rays -- not against our restatement and not through the product's loader: the
reference loads the .bsp image itself (CM_LoadMap, cm_load.c:606-745) and traces every item (cm_trace.c:545-805,
cm_test.c:217-294) in forked workers, one per host core.  Every record of every batch is compared, all 56 bytes.

  config 2  16 Mi player-hull sweeps, 20k-brush arena                      (the bench workload)
  config 3  4 Mi point + hull sweeps, 2k brushes + 512 Bezier patches
  config 4  4 Mi CM_TransformedBoxTrace + 16 Mi CM_PointContents + 4 Mi CM_TransformedPointContents, 256 inline models
  config 5  65 536 agents x 10 traces x 10 ticks on the 50k-brush arena, every batch replayed
"""
import os

import numpy as np
import pytest

from helpers import assert_traces_equal, bspgen, cmref

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not cmref.available(), reason="oracle/_ref/libcmref.so not shipped")]

CORES = len(os.sched_getaffinity(0))


@pytest.fixture(scope="module")
def eng():
    from threecore_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def test_config2_all_16Mi_records_equal_the_reference(eng):
    import torch
    from threecore_b200.engine import traces_from_tensor
    m = bspgen.arena_map(n_brushes=20000, seed=0xC0FFEE02)
    n = 1 << 24
    s, e = bspgen.make_sweeps(m, n, 0xBADC0DE2, z_range=(24.0, 2048.0))
    eng.load_bsp(m.data)
    dev = torch.device("cuda", 0)
    out = torch.empty((n, 56), dtype=torch.uint8, device=dev)
    launches = eng.launch_count
    eng.box_trace_device(torch.from_numpy(s).to(dev), torch.from_numpy(e).to(dev), bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS,
                         bspgen.MASK_PLAYERSOLID, out)
    torch.cuda.synchronize()
    eng.check_status()
    assert eng.launch_count > launches
    got = traces_from_tensor(out)
    del out
    r = cmref.RefCM().load(m.data)
    _, _, want = r.bench_box_trace(s, e, bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, nworkers=CORES,
                                   want_results=True)
    assert_traces_equal(got, want, "config 2, all 16 Mi records")
    f = got["fraction"]
    assert 0.2 < (f < 1).mean() < 0.9 and (got["startsolid"] != 0).mean() > 0.01


def test_config3_all_4Mi_patch_records_equal_the_reference(eng):
    m = bspgen.patch_map(n_brushes=2000, n_patches=512, seed=0xC0FFEE03)
    n = 1 << 22
    s, e, mins, maxs = bspgen.make_patch_sweeps(m, n, 0xBADC0DE3)
    eng.load_bsp(m.data)
    got, hit = eng.box_trace(s, e, mins, maxs, mask=bspgen.MASK_ALL, want_hit=True)
    r = cmref.RefCM().load(m.data)
    want = cmref.fork_map(n, cmref.TRACE_DTYPE,
                          lambda lo, hi: r.box_trace(s[lo:hi], e[lo:hi], mins[lo:hi], maxs[lo:hi], mask=bspgen.MASK_ALL))
    assert_traces_equal(got, want, "config 3, all 4 Mi records")
    assert (hit <= -2).mean() > 0.15, "patches are not being hit"


def test_config4_all_movers_and_point_contents_equal_the_reference(eng):
    m = bspgen.arena_map(n_brushes=20000, seed=0xC0FFEE02, n_models=256)
    eng.load_bsp(m.data)
    r = cmref.RefCM().load(m.data)
    n = 1 << 22
    t = bspgen.make_entity_traces(m, n, 0xBADC0DE4)
    got = eng.transformed_box_trace(t["starts"], t["ends"], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, t["models"],
                                    bspgen.MASK_ALL, t["origins"], t["angles"], t["boxmins"], t["boxmaxs"])
    want = cmref.fork_map(n, cmref.TRACE_DTYPE, lambda lo, hi: r.transformed_box_trace(
        t["starts"][lo:hi], t["ends"][lo:hi], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, t["models"][lo:hi], bspgen.MASK_ALL,
        t["origins"][lo:hi], t["angles"][lo:hi], t["boxmins"][lo:hi], t["boxmaxs"][lo:hi]))
    assert_traces_equal(got, want, "config 4, all 4 Mi transformed traces")
    assert 0.05 < (got["fraction"] < 1).mean() < 0.95
    pts = bspgen.make_points(m, 1 << 24, 0xBADC0DE5)
    c = eng.point_contents(pts)
    wantc = cmref.fork_map(len(pts), np.int32, lambda lo, hi: r.point_contents(pts[lo:hi]))
    assert np.array_equal(c, wantc), f"config 4: {(c != wantc).sum()} of 16 Mi contents words differ"
    assert 0.05 < (c != 0).mean() < 0.9
    pc = eng.point_contents(t["starts"], t["models"], t["origins"], t["angles"], t["boxmins"], t["boxmaxs"])
    wantp = cmref.fork_map(n, np.int32, lambda lo, hi: r.transformed_point_contents(
        t["starts"][lo:hi], t["models"][lo:hi], t["origins"][lo:hi], t["angles"][lo:hi], t["boxmins"][lo:hi],
        t["boxmaxs"][lo:hi]))
    assert np.array_equal(pc, wantp), f"config 4: {(pc != wantp).sum()} of 4 Mi transformed contents words differ"


def test_config5_rollout_65536_agents_replays_through_the_reference(eng):
    """every batch the rollout issues on the 50k-brush arena (65 536 agents, 10 ticks: 6.5 M traces) is traced again
    by the reference; the agents' kinematics are ours, parity is pinned at CM_BoxTrace"""
    from threecore_b200.rollout import Rollout, TRACES_PER_TICK
    m = bspgen.arena_map(n_brushes=50000, seed=0xC0FFEE05)
    eng.load_bsp(m.data)
    r = cmref.RefCM().load(m.data)
    n, ticks = 65536, 10
    stats = dict(batches=0, traces=0, hits=0)

    def replay(s, e, rec):
        got = np.ascontiguousarray(rec).view(cmref.TRACE_DTYPE).reshape(-1)
        _, _, want = r.bench_box_trace(s, e, bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID,
                                       nworkers=CORES, want_results=True)
        assert_traces_equal(got, want, f"rollout batch {stats['batches']}")
        stats["batches"] += 1
        stats["traces"] += len(s)
        stats["hits"] += int((got["fraction"] < 1).sum())

    ro = Rollout(eng, n, m.world_mins, m.world_maxs, seed=9, record=replay)
    ro.run(ticks)
    assert stats["traces"] == ro.traces == n * ticks * TRACES_PER_TICK
    assert stats["hits"] > 0.05 * stats["traces"], "the rollout is not exercising collisions"
