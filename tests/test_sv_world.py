"""SV_Trace / SV_PointContents as batches (include/sv_world_batch.h, cmgpu_sv_* in include/cm_gpu.h) against the
reference's own server/sv_world.c, which oracle/_ref/libcmref.so holds compiled verbatim (oracle/sv_shim.c)."""
import ctypes as C

import numpy as np
import pytest

import sv_scene
from helpers import assert_traces_equal, bspgen, cmref, get_map, have_ref, load_golden_sv
from sv_scene import ENTITYNUM_NONE, MAX_GENTITIES, f3

needs_ref = pytest.mark.skipif(not have_ref(), reason="oracle/_ref/libcmref.so not built")


def _lib():
    from threecore_b200.engine import load_library
    return sv_scene.declare(load_library())


@needs_ref
def test_host_linking_matches_the_reference():
    """SV_ClearWorld / SV_LinkEntity / SV_UnlinkEntity / SV_AreaEntities of the product's host side: the boxes that
    linking derives, and the content AND ORDER of area queries, as the reference's sv_world.c gives them"""
    m = get_map("arena_models")
    ref = cmref.RefCM().load(m.data)
    ents = sv_scene.make_entities(m, 400, seed=5)
    sv_scene.link_all_ref(ref, ents)
    lib = _lib()
    arr = sv_scene.new_game_array()
    lib.SV_LocateGameData(C.byref(arr), MAX_GENTITIES, C.sizeof(sv_scene.SvEntity))
    mb = ref.dump()["model_bounds"][0]
    lo, hi = f3(mb[:3]), f3(mb[3:])
    lib.SV_ClearWorldForBounds(lo.ctypes.data, hi.ctypes.data)
    sv_scene.link_all_product(lib, arr, ents)
    for i in range(len(ents["number"])):
        num = int(ents["number"][i])
        a, b, linked = ref.sv_entity_bounds(num)
        assert linked == 1 and arr[num].linked == 1 and arr[num].linkcount == 1
        assert f3(list(arr[num].absmin)).tobytes() == a.tobytes() and f3(list(arr[num].absmax)).tobytes() == b.tobytes(), num

    def compare_queries(seed, count):
        rng = np.random.default_rng(seed)
        buf = np.zeros(MAX_GENTITIES, np.int32)
        sizes = 0
        for q in range(count):
            c = rng.uniform(m.world_mins, m.world_maxs)
            h = rng.uniform(0, 600, size=3) * (rng.random() < 0.8) + (rng.random() < 0.1) * 6000
            qlo, qhi = f3(c - h), f3(c + h)
            want = ref.sv_area_entities(qlo, qhi)
            n = lib.SV_AreaEntities(qlo.ctypes.data, qhi.ctypes.data, buf.ctypes.data, MAX_GENTITIES)
            assert np.array_equal(buf[:n], want), (q, buf[:n], want)
            sizes += n
        return sizes

    assert compare_queries(1, 400) > 2000
    # the whole world in one query: the global order every other query is a subsequence of
    big = f3([-1e6] * 3), f3([1e6] * 3)
    assert len(ref.sv_area_entities(*big)) == len(ents["number"])
    # move a third of the entities, unlink some, link them again: chains are edited in place
    rng = np.random.default_rng(9)
    for i in rng.permutation(len(ents["number"]))[:130]:
        num = int(ents["number"][i])
        if rng.random() < 0.3:
            ref.sv_unlink_entity(num)
            lib.SV_UnlinkEntity(C.byref(arr[num]))
            assert arr[num].linked == 0
            continue
        ents["origin"][i] = f3(np.floor(rng.uniform(m.world_mins + 200, m.world_maxs - 200)))
        ref.sv_link_entity(num, ents["bmodel"][i], ents["modelindex"][i], ents["mins"][i], ents["maxs"][i], ents["origin"][i],
                           ents["angles"][i], ents["contents"][i], ents["owner"][i])
        sv_scene.fill_game_entity(arr[num], ents, i)
        lib.SV_LinkEntity(C.byref(arr[num]))
    assert compare_queries(2, 300) > 1000
    # a capped list stops where the reference's does
    buf = np.zeros(7, np.int32)
    n = lib.SV_AreaEntities(big[0].ctypes.data, big[1].ctypes.data, buf.ctypes.data, 7)
    assert n == 7 and np.array_equal(buf, ref.sv_area_entities(*big, maxcount=7))


@pytest.fixture(scope="module")
def world():
    lib = _lib()
    lib.CM_LoadMapFromMemory.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_int)]
    m = get_map("arena_models")
    assert lib.CM_LoadMapFromMemory(b"maps/test.bsp", m.data, len(m.data), None) == 0, lib.CM_LastError()
    arr = sv_scene.new_game_array()
    lib.SV_LocateGameData(C.byref(arr), MAX_GENTITIES, C.sizeof(sv_scene.SvEntity))
    lib.SV_ClearWorld()
    ents = sv_scene.make_entities(m, 300, seed=71)
    sv_scene.link_all_product(lib, arr, ents)
    yield lib, m, arr, ents
    lib.CM_ClearMap()


def _sv_trace(lib, s, e, mins, maxs, hull_stride, passent, pass_stride, masks, mask_stride):
    n = len(s)
    out = np.zeros(n, cmref.TRACE_DTYPE)
    rc = lib.SV_TraceBatch(out.ctypes.data, n, s.ctypes.data, mins.ctypes.data if mins is not None else None,
                          maxs.ctypes.data if maxs is not None else None, hull_stride, e.ctypes.data,
                          passent.ctypes.data, pass_stride, masks.ctypes.data, mask_stride)
    assert rc == 0, lib.CM_LastError()
    return out


@pytest.mark.gpu
@needs_ref
def test_sv_trace_batch_matches_the_reference(world):
    """20 000 moves with their own hulls, pass entities and masks through ONE device batch vs 20 000 SV_Trace calls"""
    lib, m, arr, ents = world
    ref = cmref.RefCM().load(m.data)
    sv_scene.link_all_ref(ref, ents)
    s, e, mins, maxs, passent, masks = sv_scene.make_moves(m, ents, 20000, seed=72)
    got = _sv_trace(lib, s, e, mins, maxs, 1, passent, 1, masks, 1)
    want = ref.sv_trace(s, e, mins, maxs, passent, masks)
    assert_traces_equal(got, want, "SV_TraceBatch")
    hit_ent = (got["entityNum"] != ENTITYNUM_NONE) & (got["entityNum"] != sv_scene.ENTITYNUM_WORLD)
    assert hit_ent.mean() > 0.2, "entities are not being hit"
    assert (got["startsolid"] == 1).sum() > 50 and (got["allsolid"] == 1).sum() > 20
    assert lib.SV_LastBatchCandidates() > 20000
    # shared hull / pass entity / mask, NULL hull (= point, sv_world.c:565-570)
    one_pass = np.array([int(ents["number"][3])], np.int32)
    one_mask = np.array([bspgen.MASK_PLAYERSOLID], np.int32)
    pm, px = f3(bspgen.PLAYER_MINS), f3(bspgen.PLAYER_MAXS)
    got = _sv_trace(lib, s, e, pm, px, 0, one_pass, 0, one_mask, 0)
    assert_traces_equal(got, ref.sv_trace(s, e, pm, px, one_pass[0], one_mask[0]), "SV_TraceBatch, shared arguments")
    got = _sv_trace(lib, s[:3000], e[:3000], None, None, 0, one_pass, 0, one_mask, 0)
    assert_traces_equal(got, ref.sv_trace(s[:3000], e[:3000], None, None, one_pass[0], one_mask[0]), "SV_TraceBatch, NULL hull")


@pytest.mark.gpu
@needs_ref
def test_sv_trace_after_entities_move(world):
    """the snapshot is taken per batch: relinking between batches changes the answers as it does in the reference"""
    lib, m, arr, ents = world
    ref = cmref.RefCM().load(m.data)
    sv_scene.link_all_ref(ref, ents)
    rng = np.random.default_rng(17)
    ents2 = {k: v.copy() for k, v in ents.items()}
    moved = rng.permutation(len(ents["number"]))[:100]
    for i in moved:
        num = int(ents2["number"][i])
        ents2["origin"][i] = f3(ents2["origin"][i] + np.floor(rng.uniform(-300, 300, size=3)))
        ents2["angles"][i] = f3(rng.uniform(0, 360, size=3) * (rng.random() < 0.5))
        ref.sv_link_entity(num, ents2["bmodel"][i], ents2["modelindex"][i], ents2["mins"][i], ents2["maxs"][i], ents2["origin"][i],
                           ents2["angles"][i], ents2["contents"][i], ents2["owner"][i])
        sv_scene.fill_game_entity(arr[num], ents2, i)
        lib.SV_LinkEntity(C.byref(arr[num]))
    gone = [int(ents2["number"][i]) for i in rng.permutation(len(ents["number"]))[:20]]
    for num in gone:
        ref.sv_unlink_entity(num)
        lib.SV_UnlinkEntity(C.byref(arr[num]))
    s, e, mins, maxs, passent, masks = sv_scene.make_moves(m, ents2, 8000, seed=73)
    got = _sv_trace(lib, s, e, mins, maxs, 1, passent, 1, masks, 1)
    assert_traces_equal(got, ref.sv_trace(s, e, mins, maxs, passent, masks), "SV_TraceBatch after relinking")
    assert not np.isin(got["entityNum"], gone).any()
    # restore the module's scene for the tests that follow
    lib.SV_ClearWorld()
    sv_scene.link_all_product(lib, arr, ents)


@pytest.mark.gpu
@needs_ref
def test_sv_point_contents_world_batch(world):
    lib, m, arr, ents = world
    ref = cmref.RefCM().load(m.data)
    sv_scene.link_all_ref(ref, ents)
    rng = np.random.default_rng(74)
    n = 30000
    tgt = rng.integers(0, len(ents["number"]), size=n)
    pts = f3(ents["origin"][tgt] + rng.uniform(-60, 60, size=(n, 3)))
    pts[: n // 5] = f3(rng.uniform(m.world_mins, m.world_maxs, size=(n // 5, 3)))
    passent = np.where(rng.random(n) < 0.3, ents["number"][tgt], ENTITYNUM_NONE).astype(np.int32)
    got = np.zeros(n, np.int32)
    assert lib.SV_PointContentsWorldBatch(got.ctypes.data, n, pts.ctypes.data, passent.ctypes.data, 1) == 0, lib.CM_LastError()
    want = ref.sv_point_contents(pts, passent)
    assert np.array_equal(got, want), np.flatnonzero(got != want)[:10]
    world_only = ref.point_contents(pts)
    assert (got != world_only).mean() > 0.05, "entities do not contribute"
    one = np.array([ENTITYNUM_NONE], np.int32)
    assert lib.SV_PointContentsWorldBatch(got.ctypes.data, n, pts.ctypes.data, one.ctypes.data, 0) == 0
    assert np.array_equal(got, ref.sv_point_contents(pts, ENTITYNUM_NONE))


@pytest.mark.gpu
def test_sv_trace_golden(world):
    """committed vectors made with the reference's sv_world.c (tests/golden/make_golden_sv.py): needs no reference library"""
    lib, m, arr, ents = world
    g = load_golden_sv()
    assert g["bsp_len"] == len(m.data), "golden set was made for another map: rerun tests/golden/make_golden_sv.py"
    got = _sv_trace(lib, g["starts"], g["ends"], g["mins"], g["maxs"], 1, g["passent"], 1, g["masks"], 1)
    assert_traces_equal(got, g["traces"], "SV_TraceBatch vs golden")
    c = np.zeros(len(g["points"]), np.int32)
    assert lib.SV_PointContentsWorldBatch(c.ctypes.data, len(c), g["points"].ctypes.data, g["point_pass"].ctypes.data, 1) == 0
    assert np.array_equal(c, g["contents"])


@pytest.mark.gpu
def test_sv_entry_points_reject_bad_input(world):
    lib, m, arr, ents = world
    from threecore_b200.engine import load_library
    vp = C.c_void_p
    lib.CM_Context.restype = vp
    ctx = lib.CM_Context()
    lib.cmgpu_sv_set_entities.argtypes = [vp, C.c_int, vp, vp, C.c_int]
    lib.cmgpu_last_error.argtypes = [vp]
    lib.cmgpu_last_error.restype = C.c_char_p

    class E(C.Structure):      # cmgpu_sv_entity
        _fields_ = [("number", C.c_int), ("ownerNum", C.c_int), ("contents", C.c_int), ("bmodel", C.c_int), ("modelindex", C.c_int),
                    ("mins", sv_scene.V3), ("maxs", sv_scene.V3), ("absmin", sv_scene.V3), ("absmax", sv_scene.V3),
                    ("origin", sv_scene.V3), ("angles", sv_scene.V3)]
    bad = (E * 1)()
    bad[0].bmodel = 1
    bad[0].modelindex = m.num_models + 5
    assert lib.cmgpu_sv_set_entities(ctx, 1, C.byref(bad), None, 0) == 5      # CMGPU_ERR_BAD_HANDLE, CM_InlineModel's ERR_DROP
    assert b"CM_InlineModel" in lib.cmgpu_last_error(ctx)
    # the previous snapshot is still in place: an empty batch and a real one both work
    s, e, mins, maxs, passent, masks = sv_scene.make_moves(m, ents, 64, seed=75)
    assert lib.SV_TraceBatch(None, 0, None, None, None, 0, None, None, 0, None, 0) == 0
    out = _sv_trace(lib, s, e, mins, maxs, 1, passent, 1, masks, 1)
    assert np.isfinite(out["fraction"]).all()
