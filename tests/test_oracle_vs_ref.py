"""The C restatement (oracle/cm_oracle.c) against the UNMODIFIED reference (oracle/_ref/libcmref.so).

This is what pins the oracle: same generated maps, same rays, byte-identical 56-byte trace_t records
and identical contents words.  Skipped only where the reference library has not been built
(it is built wherever /root/reference exists and travels to the GPU box as a binary)."""
import numpy as np
import pytest

import cases
from helpers import assert_traces_equal, bspgen, cmoracle, cmref, get_map, have_ref

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref/libcmref.so not built (no reference tree here)")

N = 12000


def _pair(kind, **cvars):
    m = get_map(kind)
    r = cmref.RefCM().load(m.data)
    if cvars:
        r.set_cvars(**cvars)
    o = cmoracle.OracleCM(cmoracle.flat_from_ref_dump(r.dump()), **cvars)
    return r, o


def _check(r, o, case):
    a = cases.run_case(r, case)
    b = cases.run_case(o, case)
    if case["kind"] == "points":
        assert np.array_equal(a, b), f"{case['name']}: {(a != b).sum()} contents words differ"
    else:
        assert_traces_equal(b, a, case["name"])
    return a


@pytest.mark.parametrize("kind", ["tiny", "room", "arena_small", "arena_models", "patches_small"])
def test_world_traces(kind):
    r, o = _pair(kind)
    hits = 0
    for case in cases.world_cases(kind, n=N):
        a = _check(r, o, case)
        hits += int((a["fraction"] < 1).sum())
    assert hits > 0


@pytest.mark.parametrize("kind", ["arena_models", "patches_small"])
def test_entity_traces_and_point_contents(kind):
    r, o = _pair(kind)
    for case in cases.entity_cases(kind, n=N):
        a = _check(r, o, case)
        if case["kind"] != "points":
            assert (a["fraction"] < 1).mean() > 0.05, case["name"]


@pytest.mark.parametrize("cvars", [dict(noCurves=1, playerCurveClip=1), dict(noCurves=0, playerCurveClip=0)])
def test_patch_cvars(cvars):
    r, o = _pair("patches_small", **cvars)
    for case in cases.world_cases("patches_small", n=N // 2)[:2]:
        _check(r, o, case)
    r.set_cvars(0, 1)


def test_patch_hits_happen():
    """the patch cases must actually exercise facet code: a share of hits carries a patch plane"""
    kind = "patches_small"
    m = get_map(kind)
    r, o = _pair(kind)
    s, e, mins, maxs = cases.bspgen.make_patch_sweeps(m, N, 77)
    a = r.box_trace(s, e, mins, maxs, mask=cases.bspgen.MASK_ALL)
    b, info = o.box_trace(s, e, mins, maxs, mask=cases.bspgen.MASK_ALL, want_info=True)
    assert_traces_equal(b, a, "patch sweeps")
    assert (info["hitPatch"] >= 0).mean() > 0.1


def test_angle_vectors_match_reference():
    r = cmref.RefCM()
    o = cmoracle.OracleCM(cmoracle.flat_from_ref_dump(r.load(get_map("tiny").data).dump()))
    rng = np.random.default_rng(5)
    for ang in np.concatenate([rng.uniform(-720, 720, size=(500, 3)), np.array([[0, 90, 0], [90, 0, 0], [0, 0, 180], [45, 45, 45]])]):
        fa, ra, ua = r.angle_vectors(ang)
        fb, rb, ub = o.angle_vectors(ang)
        assert fa.tobytes() == fb.tobytes() and ra.tobytes() == rb.tobytes() and ua.tobytes() == ub.tobytes()


def test_repeat_visits_do_not_change_results():
    """the oracle (and the GPU) drop the reference's checkcount stamps; a long sweep through many leafs
    that list the same big brushes must still agree"""
    r, o = _pair("arena_small")
    m = get_map("arena_small")
    rng = np.random.default_rng(9)
    lo, hi = m.world_mins + 100, m.world_maxs - 100
    s = rng.uniform(lo, hi, size=(N, 3)).astype(np.float32)
    e = rng.uniform(lo, hi, size=(N, 3)).astype(np.float32)      # wall-to-wall sweeps
    a = r.box_trace(s, e, cases.bspgen.PLAYER_MINS, cases.bspgen.PLAYER_MAXS, mask=-1)
    b, info = o.box_trace(s, e, cases.bspgen.PLAYER_MINS, cases.bspgen.PLAYER_MAXS, mask=-1, want_info=True)
    assert_traces_equal(b, a, "long sweeps")
    assert info["leafs"].max() > 8


def test_deep_comb_tree_restatement_matches_reference():
    """a BSP 155 levels deep: the reference recurses, the restatement keeps explicit stacks"""
    from helpers import comb_map
    m, half = comb_map(150)
    r = cmref.RefCM().load(m.data)
    o = cmoracle.OracleCM(cmoracle.flat_from_ref_dump(r.dump()))
    rng = np.random.default_rng(5)
    n = 3000
    s = np.stack([rng.uniform(-half + 70, half - 70, n), rng.uniform(-20, 20, n), rng.uniform(-20, 20, n)], 1).astype(np.float32)
    e = s.copy()
    e[:, 0] = np.where(rng.random(n) < 0.5, half - 70, -half + 70).astype(np.float32)
    e[: n // 8] = s[: n // 8]
    for mask in (bspgen.MASK_PLAYERSOLID, 0x40000000):
        for mins, maxs in ((bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS), (None, None)):
            assert_traces_equal(o.box_trace(s, e, mins, maxs, mask=mask), r.box_trace(s, e, mins, maxs, mask=mask), f"comb {mask:#x}")
    lo, hi = np.array([[-half, -40, -40]], np.float32), np.array([[half, 40, 40]], np.float32)
    a, b = o.box_leafnums(lo, hi, 1024), r.box_leafnums(lo, hi, 1024)
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and a[1][0] > 150
