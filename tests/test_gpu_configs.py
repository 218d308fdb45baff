"""BASELINE.json configs 1, 3 and 4 at (or near) their full sizes: the engine's records against the oracle on a
sample, whole-batch invariants, and agreement between the host-pointer and device-pointer entry points.
(config 2 = tests/test_gpu_parity.py `big`, config 5 = tests/test_rollout.py.)"""
import numpy as np
import pytest

from helpers import assert_traces_equal, bspgen, cmoracle

pytestmark = pytest.mark.gpu

SAMPLE = 60000


@pytest.fixture(scope="module")
def eng():
    from threecore_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def _sample(n, seed):
    return np.sort(np.random.default_rng(seed).choice(n, size=min(SAMPLE, n), replace=False))


def test_config1_point_traces_room(eng):
    """1 Mi point traces, ~500-brush axial room, MASK_SOLID (the reference's own CPU-runnable case)"""
    m = bspgen.room_map(n_brushes=500, seed=0xC0FFEE01)
    fm = eng.load_bsp(m.data)
    n = 1 << 20
    s, e = bspgen.make_point_sweeps(m, n, 0xBADC0DE1)
    got = eng.box_trace(s, e, None, None, mask=bspgen.MASK_SOLID)
    idx = _sample(n, 1)
    want = cmoracle.OracleCM(fm.to_dict()).box_trace(s[idx], e[idx], None, None, mask=bspgen.MASK_SOLID)
    assert_traces_equal(got[idx], want, "config 1")
    f = got["fraction"]
    assert ((f >= 0) & (f <= 1)).all() and 0.2 < (f < 1).mean() < 0.99
    assert np.array_equal(got["endpos"][f == 1], e[f == 1])


def test_config3_patches_mixed_hulls(eng):
    """2k brushes + 512 Bezier patches, 1 Mi traces, half points / half player hulls, aimed so that patches get hit"""
    m = bspgen.patch_map(n_brushes=2000, n_patches=512, seed=0xC0FFEE03)
    fm = eng.load_bsp(m.data)
    n = 1 << 20
    s, e, mins, maxs = bspgen.make_patch_sweeps(m, n, 0xBADC0DE3)
    got, hit = eng.box_trace(s, e, mins, maxs, mask=bspgen.MASK_ALL, want_hit=True)
    idx = _sample(n, 3)
    want, info = cmoracle.OracleCM(fm.to_dict()).box_trace(s[idx], e[idx], mins[idx], maxs[idx], mask=bspgen.MASK_ALL,
                                                          want_info=True)
    assert_traces_equal(got[idx], want, "config 3")
    want_hit = np.where(info["hitPatch"] >= 0, -2 - info["hitPatch"], info["hitBrush"])
    assert np.array_equal(hit[idx], want_hit)
    assert (hit <= -2).mean() > 0.15, "patches are not being hit"


def test_config4_movers_and_point_contents(eng):
    """20k-brush arena + 256 inline models: transformed traces (rotations, temp boxes) and batched point contents"""
    m = bspgen.arena_map(n_brushes=20000, seed=0xC0FFEE02, n_models=256)
    fm = eng.load_bsp(m.data)
    o = cmoracle.OracleCM(fm.to_dict())
    n = 1 << 20
    t = bspgen.make_entity_traces(m, n, 0xBADC0DE4)
    got = eng.transformed_box_trace(t["starts"], t["ends"], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, t["models"],
                                    bspgen.MASK_ALL, t["origins"], t["angles"], t["boxmins"], t["boxmaxs"])
    idx = _sample(n, 4)
    want = o.transformed_box_trace(t["starts"][idx], t["ends"][idx], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS,
                                   t["models"][idx], bspgen.MASK_ALL, t["origins"][idx], t["angles"][idx],
                                   t["boxmins"][idx], t["boxmaxs"][idx])
    assert_traces_equal(got[idx], want, "config 4 transformed")
    assert 0.05 < (got["fraction"] < 1).mean() < 0.95
    pts = bspgen.make_points(m, 1 << 22, 0xBADC0DE5)
    c = eng.point_contents(pts)
    idx = _sample(len(pts), 5)
    assert np.array_equal(c[idx], o.point_contents(pts[idx])), "config 4 point contents"
    assert 0.05 < (c != 0).mean() < 0.9
    pc = eng.point_contents(t["starts"], t["models"], t["origins"], t["angles"], t["boxmins"], t["boxmaxs"])
    idx = _sample(n, 6)
    wantc = o.point_contents(t["starts"][idx], t["models"][idx], t["origins"][idx], t["angles"][idx],
                             t["boxmins"][idx], t["boxmaxs"][idx])
    assert np.array_equal(pc[idx], wantc), "config 4 transformed point contents"
