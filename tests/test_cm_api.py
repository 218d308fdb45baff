"""The reference-facing layer (include/cm_public_batch.h): the CM_* names of cm_public.h and the SV_ClipMoveToEntities
batching hook, called through ctypes exactly as the engine would bind them, against the oracle.

The server-side merge (sv_world.c:497-549, 562-609) is restated below in Python -- it is a dozen lines of flag logic over
the per-entity traces, which come from the oracle's CM_TransformedBoxTrace."""
import ctypes as C

import numpy as np
import pytest

from helpers import assert_traces_equal, bspgen, cmoracle, cmref, get_map

pytestmark = pytest.mark.gpu

ENTITYNUM_NONE, ENTITYNUM_WORLD = 4095, 4094
V3 = C.c_float * 3


class Touch(C.Structure):          # cm_touch_t
    _fields_ = [("entityNum", C.c_int), ("bmodel", C.c_int), ("modelindex", C.c_int),
                ("mins", V3), ("maxs", V3), ("origin", V3), ("angles", V3)]


@pytest.fixture(scope="module")
def api():
    from threecore_b200.engine import load_library
    lib = load_library()
    vp = C.c_void_p
    lib.CM_LoadMapFromMemory.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.POINTER(C.c_int)]
    lib.CM_LastError.restype = C.c_char_p
    lib.CM_TempBoxModel.argtypes = [vp, vp]
    lib.CM_BoxTrace.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int]
    lib.CM_BoxTrace.restype = None
    lib.CM_TransformedBoxTrace.argtypes = [vp, vp, vp, vp, vp, C.c_int, C.c_int, vp, vp]
    lib.CM_TransformedBoxTrace.restype = None
    lib.CM_PointContents.argtypes = [vp, C.c_int]
    lib.CM_TransformedPointContents.argtypes = [vp, C.c_int, vp, vp]
    lib.CM_BoxTraceBatch.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.c_int, C.c_int]
    lib.CM_ModelBounds.argtypes = [C.c_int, vp, vp]
    lib.CM_ModelBounds.restype = None
    lib.SV_ClipMovesToEntitiesBatch.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.c_int, vp, C.c_int, vp, vp]
    lib.SV_PointContentsBatch.argtypes = [vp, C.c_int, vp, vp, vp]
    m = get_map("arena_models")
    assert lib.CM_LoadMapFromMemory(b"maps/test.bsp", m.data, len(m.data), None) == 0, lib.CM_LastError()
    from threecore_b200.engine import FlatMap
    o = cmoracle.OracleCM(FlatMap.from_bsp(m.data).to_dict())
    yield lib, m, o
    lib.CM_ClearMap()


def f3(a):
    return np.ascontiguousarray(a, np.float32)


def test_single_calls_keep_the_reference_signatures(api):
    """CM_BoxTrace / CM_TransformedBoxTrace / CM_PointContents one item at a time, as SV_Trace and trap_Trace call them"""
    lib, m, o = api
    n = 300
    s, e = bspgen.make_sweeps(m, n, 21, len_short=(1.0, 300.0), len_long=(300.0, 2000.0))
    mins, maxs = f3(bspgen.PLAYER_MINS), f3(bspgen.PLAYER_MAXS)
    got = np.zeros(n, cmref.TRACE_DTYPE)
    for i in range(n):
        lib.CM_BoxTrace(got[i:i + 1].ctypes.data, s[i].ctypes.data, e[i].ctypes.data, mins.ctypes.data, maxs.ctypes.data,
                        0, bspgen.MASK_PLAYERSOLID)
    assert_traces_equal(got, o.box_trace(s, e, mins, maxs, mask=bspgen.MASK_PLAYERSOLID), "CM_BoxTrace")
    # NULL mins/maxs mean a point (cm_trace.c:569-574)
    got2 = np.zeros(n, cmref.TRACE_DTYPE)
    for i in range(n):
        lib.CM_BoxTrace(got2[i:i + 1].ctypes.data, s[i].ctypes.data, e[i].ctypes.data, None, None, 0, bspgen.MASK_SOLID)
    assert_traces_equal(got2, o.box_trace(s, e, None, None, mask=bspgen.MASK_SOLID), "CM_BoxTrace NULL hull")
    t = bspgen.make_entity_traces(m, n, 22)
    got3 = np.zeros(n, cmref.TRACE_DTYPE)
    pc = np.zeros(n, np.int32)
    for i in range(n):
        h = int(t["models"][i])
        if h == 1023:
            h = lib.CM_TempBoxModel(t["boxmins"][i].ctypes.data, t["boxmaxs"][i].ctypes.data)
        lib.CM_TransformedBoxTrace(got3[i:i + 1].ctypes.data, t["starts"][i].ctypes.data, t["ends"][i].ctypes.data,
                                   mins.ctypes.data, maxs.ctypes.data, h, bspgen.MASK_ALL,
                                   t["origins"][i].ctypes.data, t["angles"][i].ctypes.data)
        pc[i] = lib.CM_TransformedPointContents(t["starts"][i].ctypes.data, h, t["origins"][i].ctypes.data,
                                                t["angles"][i].ctypes.data)
    want3 = o.transformed_box_trace(t["starts"], t["ends"], mins, maxs, t["models"], bspgen.MASK_ALL, t["origins"], t["angles"],
                                    t["boxmins"], t["boxmaxs"])
    assert_traces_equal(got3, want3, "CM_TransformedBoxTrace")
    assert np.array_equal(pc, o.point_contents(t["starts"], t["models"], t["origins"], t["angles"], t["boxmins"], t["boxmaxs"]))
    pts = bspgen.make_points(m, n, 23)
    c = np.array([lib.CM_PointContents(pts[i].ctypes.data, 0) for i in range(n)], np.int32)
    assert np.array_equal(c, o.point_contents(pts))


def _sv_trace_reference(o, s, e, mins, maxs, mask, touch_lists, ents):
    """SV_Trace (sv_world.c:562-609) + SV_ClipMoveToEntities (:476-551) over the oracle's CM calls"""
    n = len(s)
    clip = o.box_trace(s, e, mins, maxs, mask=mask).copy()
    clip["entityNum"] = np.where(clip["fraction"] != 1.0, ENTITYNUM_WORLD, ENTITYNUM_NONE)
    # every (move, candidate) pair in one oracle batch
    mv = np.concatenate([np.full(len(t), i) for i, t in enumerate(touch_lists)]).astype(np.int64)
    en = np.concatenate(touch_lists).astype(np.int64)
    tr = o.transformed_box_trace(s[mv], e[mv], mins, maxs, ents["handle"][en], mask, ents["origin"][en], ents["angles"][en],
                                 ents["mins"][en], ents["maxs"][en])
    k = 0
    for i in range(n):
        for ent in touch_lists[i]:
            t = tr[k].copy()
            k += 1
            if clip[i]["allsolid"]:
                continue
            if t["allsolid"]:
                clip[i]["allsolid"] = 1
                t["entityNum"] = ents["number"][ent]
            elif t["startsolid"]:
                clip[i]["startsolid"] = 1
                t["entityNum"] = ents["number"][ent]
            if t["fraction"] < clip[i]["fraction"]:
                old = clip[i]["startsolid"]
                t["entityNum"] = ents["number"][ent]
                clip[i] = t
                clip[i]["startsolid"] |= old
    return clip


def test_sv_clip_moves_to_entities_batch(api):
    """world trace, then all candidates of all moves in ONE batch, merged with the reference's rules"""
    lib, m, o = api
    rng = np.random.default_rng(31)
    nm = m.num_models
    E = 400
    ents = dict(number=np.arange(E, dtype=np.int32) + 64,
                bmodel=(rng.random(E) < 0.5), origin=f3(np.floor(rng.uniform(m.world_mins + 300, m.world_maxs - 300, size=(E, 3)))),
                angles=f3(np.where(rng.random((E, 1)) < 0.5, 0.0, rng.uniform(0, 360, size=(E, 3)))),
                mins=f3(-np.floor(rng.uniform(8, 48, size=(E, 3)))), maxs=f3(np.floor(rng.uniform(8, 64, size=(E, 3)))))
    ents["modelindex"] = rng.integers(1, nm, size=E).astype(np.int32)
    ents["handle"] = np.where(ents["bmodel"], ents["modelindex"], 1023).astype(np.int32)
    ents["angles"][~ents["bmodel"]] = 0.0            # box entities do not rotate (sv_world.c:35-43 uses their bounds)
    n = 3000
    # moves aimed at entities so that the merge rules are exercised
    tgt = rng.integers(0, E, size=n)
    d = rng.normal(size=(n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    reach = rng.uniform(40, 300, size=(n, 1))
    s = f3(ents["origin"][tgt] + rng.uniform(-40, 40, size=(n, 3)) - d * reach)
    e = f3(ents["origin"][tgt] + rng.uniform(-40, 40, size=(n, 3)) + d * reach * rng.uniform(-0.1, 1.0, size=(n, 1)))
    s[:50] = ents["origin"][tgt[:50]]               # starts inside entities: startsolid / allsolid merging
    mins, maxs = f3(bspgen.PLAYER_MINS), f3(bspgen.PLAYER_MAXS)
    mask = bspgen.MASK_PLAYERSOLID | bspgen.CONTENTS_BODY
    # world pass through the batched CM_BoxTrace, entityNum as SV_Trace sets it
    clip = np.zeros(n, cmref.TRACE_DTYPE)
    assert lib.CM_BoxTraceBatch(clip.ctypes.data, n, s.ctypes.data, e.ctypes.data, mins.ctypes.data, maxs.ctypes.data, 0, mask) == 0
    clip["entityNum"] = np.where(clip["fraction"] != 1.0, ENTITYNUM_WORLD, ENTITYNUM_NONE)
    # broad phase stays with the caller: the target plus a few random candidates, none for moves the world blocks at once
    touch_lists = []
    for i in range(n):
        if clip[i]["fraction"] == 0:
            touch_lists.append(np.zeros(0, np.int64))
        else:
            touch_lists.append(np.unique(np.concatenate([[tgt[i]], rng.integers(0, E, size=rng.integers(0, 5))])))
    first = np.zeros(n + 1, np.int32)
    first[1:] = np.cumsum([len(t) for t in touch_lists])
    flat = np.concatenate(touch_lists).astype(np.int64)
    arr = (Touch * len(flat))()
    for k, ent in enumerate(flat):
        arr[k].entityNum = int(ents["number"][ent]); arr[k].bmodel = int(ents["bmodel"][ent]); arr[k].modelindex = int(ents["modelindex"][ent])
        arr[k].mins = V3(*ents["mins"][ent]); arr[k].maxs = V3(*ents["maxs"][ent])
        arr[k].origin = V3(*ents["origin"][ent]); arr[k].angles = V3(*ents["angles"][ent])
    masks = np.array([mask], np.int32)
    rc = lib.SV_ClipMovesToEntitiesBatch(clip.ctypes.data, n, s.ctypes.data, e.ctypes.data, mins.ctypes.data, maxs.ctypes.data, 0,
                                        masks.ctypes.data, 0, first.ctypes.data, C.byref(arr))
    assert rc == 0, lib.CM_LastError()
    want = _sv_trace_reference(o, s, e, mins, maxs, mask, touch_lists, ents)
    assert_traces_equal(clip, want, "SV_ClipMovesToEntitiesBatch")
    ent_hits = (clip["entityNum"] >= 64) & (clip["entityNum"] < 64 + E)
    assert ent_hits.mean() > 0.2, "entities are not being hit"
    assert (clip["startsolid"] == 1).sum() > 10


def test_sv_point_contents_batch(api):
    """SV_PointContents: world contents OR-ed with CM_TransformedPointContents of every candidate entity"""
    lib, m, o = api
    rng = np.random.default_rng(41)
    nm = m.num_models
    E, n = 200, 5000
    bmodel = rng.random(E) < 0.6
    origin = f3(np.floor(rng.uniform(m.world_mins + 300, m.world_maxs - 300, size=(E, 3))))
    angles = f3(np.where(rng.random((E, 1)) < 0.5, 0.0, rng.uniform(0, 360, size=(E, 3))))
    angles[~bmodel] = 0.0
    emins, emaxs = f3(-np.floor(rng.uniform(8, 48, size=(E, 3)))), f3(np.floor(rng.uniform(8, 64, size=(E, 3))))
    modelindex = rng.integers(1, nm, size=E).astype(np.int32)
    handle = np.where(bmodel, modelindex, 1023).astype(np.int32)
    tgt = rng.integers(0, E, size=n)
    pts = f3(origin[tgt] + rng.uniform(-60, 60, size=(n, 3)))
    lists = [np.unique(np.concatenate([[tgt[i]], rng.integers(0, E, size=rng.integers(0, 4))])) for i in range(n)]
    first = np.zeros(n + 1, np.int32); first[1:] = np.cumsum([len(t) for t in lists])
    flat = np.concatenate(lists).astype(np.int64)
    arr = (Touch * len(flat))()
    for k, ent in enumerate(flat):
        arr[k].entityNum = int(ent); arr[k].bmodel = int(bmodel[ent]); arr[k].modelindex = int(modelindex[ent])
        arr[k].mins = V3(*emins[ent]); arr[k].maxs = V3(*emaxs[ent]); arr[k].origin = V3(*origin[ent]); arr[k].angles = V3(*angles[ent])
    got = np.zeros(n, np.int32)
    assert lib.SV_PointContentsBatch(got.ctypes.data, n, pts.ctypes.data, first.ctypes.data, C.byref(arr)) == 0, lib.CM_LastError()
    want = o.point_contents(pts).copy()
    mv = np.concatenate([np.full(len(t), i) for i, t in enumerate(lists)])
    c2 = o.point_contents(pts[mv], handle[flat], origin[flat], angles[flat], emins[flat], emaxs[flat])
    np.bitwise_or.at(want, mv, c2)
    assert np.array_equal(got, want)
    assert (got != o.point_contents(pts)).mean() > 0.05, "entities do not contribute"


# ---- leaf queries, visibility and area portals under their cm_public.h names (cm_public.h:54-69) --------------------
@pytest.fixture(scope="module")
def vis_api(api):
    """the same library with a map that has clusters, areas and a visibility lump; the arena map is put back afterwards"""
    import leaf_cases as lc
    lib, m0, _ = api
    vp = C.c_void_p
    lib.CM_PointLeafnum.argtypes = [vp]
    lib.CM_BoxLeafnums.argtypes = [vp, vp, vp, C.c_int, vp]
    lib.CM_LeafCluster.argtypes = [C.c_int]
    lib.CM_LeafArea.argtypes = [C.c_int]
    lib.CM_PointLeafnumBatch.argtypes = [vp, C.c_int, vp]
    lib.CM_BoxLeafnumsBatch.argtypes = [C.c_int, vp, vp, C.c_int, vp, vp, vp]
    lib.CM_ClusterPVS.argtypes = [C.c_int]
    lib.CM_ClusterPVS.restype = C.POINTER(C.c_ubyte)
    lib.CM_AdjustAreaPortalState.argtypes = [C.c_int, C.c_int, C.c_int]
    lib.CM_AdjustAreaPortalState.restype = None
    lib.CM_AreasConnected.argtypes = [C.c_int, C.c_int]
    lib.CM_WriteAreaBits.argtypes = [vp, C.c_int]
    lib.SV_EntityClustersBatch.argtypes = [C.c_int, vp, vp, vp]
    lib.SV_VisibleFromPointsBatch.argtypes = [C.c_int, vp, C.c_int, vp, vp, vp]
    m = lc.vis_map("arena_small_vis")
    assert lib.CM_LoadMapFromMemory(b"maps/vis.bsp", m.data, len(m.data), None) == 0, lib.CM_LastError()
    from threecore_b200.engine import FlatMap
    o = cmoracle.OracleCM(FlatMap.from_bsp(m.data).to_dict())
    yield lib, m, o, lc
    assert lib.CM_LoadMapFromMemory(b"maps/test.bsp", m0.data, len(m0.data), None) == 0, lib.CM_LastError()


def test_leaf_queries_keep_the_reference_signatures(vis_api):
    lib, m, o, lc = vis_api
    lca = o.flat["leaf_cluster_area"]
    pts = lc.viewpoints(m, 200, 61)
    want = o.point_leafnum(pts)
    got = np.array([lib.CM_PointLeafnum(pts[i].ctypes.data) for i in range(len(pts))], np.int32)
    assert np.array_equal(got, want)
    assert [lib.CM_LeafCluster(int(l)) for l in want[:50]] == lca[want[:50], 0].tolist()
    assert [lib.CM_LeafArea(int(l)) for l in want[:50]] == lca[want[:50], 1].tolist()
    batch = np.zeros(len(pts), np.int32)
    assert lib.CM_PointLeafnumBatch(batch.ctypes.data, len(pts), pts.ctypes.data) == 0
    assert np.array_equal(batch, want)
    mins, maxs = lc.query_boxes(m, 200, 62)
    wl, wc, wlast = o.box_leafnums(mins, maxs, 32)
    for i in range(len(mins)):
        lst = np.full(32, -1, np.int32)
        last = C.c_int(-5)
        n = lib.CM_BoxLeafnums(mins[i].ctypes.data, maxs[i].ctypes.data, lst.ctypes.data, 32, C.byref(last))
        assert n == wc[i] and last.value == wlast[i] and np.array_equal(lst, wl[i]), f"box {i}"
    # lastLeaf may be NULL (cm_public.h:58 callers that do not care)
    lst = np.zeros(4, np.int32)
    assert lib.CM_BoxLeafnums(mins[0].ctypes.data, maxs[0].ctypes.data, lst.ctypes.data, 4, None) == min(4, o.box_leafnums(mins[:1], maxs[:1], 4)[1][0])
    lists, counts, lasts = np.full((len(mins), 32), -1, np.int32), np.zeros(len(mins), np.int32), np.zeros(len(mins), np.int32)
    assert lib.CM_BoxLeafnumsBatch(len(mins), mins.ctypes.data, maxs.ctypes.data, 32, lists.ctypes.data, counts.ctypes.data, lasts.ctypes.data) == 0
    assert np.array_equal(lists, wl) and np.array_equal(counts, wc) and np.array_equal(lasts, wlast)


def test_visibility_and_area_portals_keep_the_reference_signatures(vis_api):
    lib, m, o, lc = vis_api
    nc, cb, vis = lc.vis_lump(m.data)
    nA = m.stats["areas"]
    assert lib.CM_NumClusters() == nc
    for cl in (0, 5, nc - 1):
        row = np.ctypeslib.as_array(lib.CM_ClusterPVS(cl), shape=(cb,))
        assert np.array_equal(row, vis[cl * cb:(cl + 1) * cb])
    row = np.ctypeslib.as_array(lib.CM_ClusterPVS(-1), shape=(cb,))          # out of range: row 0 (cm_test.c:307-309)
    assert np.array_equal(row, vis[:cb])
    portals = np.zeros((nA, nA), np.int32)
    mins, maxs = lc.query_boxes(m, 120, 63)
    ents21 = o.entity_clusters(mins, maxs)
    ents = np.zeros((len(mins), 21), np.int32)
    assert lib.SV_EntityClustersBatch(len(mins), mins.ctypes.data, maxs.ctypes.data, ents.ctypes.data) == 0
    assert np.array_equal(ents, ents21)
    views = lc.viewpoints(m, 40, 64)
    words = (len(mins) + 31) // 32
    for step, (p, q, open_) in enumerate(lc.portal_script(nA, 65, steps=6)):
        lib.CM_AdjustAreaPortalState(p, q, int(open_))
        d = 1 if open_ else -1
        portals[p, q] += d
        portals[q, p] += d
        flood = o.flood_areas(portals)
        for a in range(nA):
            assert [bool(lib.CM_AreasConnected(a, b)) for b in range(nA)] == (flood == flood[a]).tolist()
            buf = np.zeros(8, np.uint8)
            assert lib.CM_WriteAreaBits(buf.ctypes.data, a) == (nA + 7) // 8
            assert np.array_equal(np.unpackbits(buf, bitorder="little")[:nA].astype(bool), flood == flood[a])
        assert not lib.CM_AreasConnected(-1, 0)
        bits = np.zeros((len(views), words), np.uint32)
        leafs = np.zeros(len(views), np.int32)
        assert lib.SV_VisibleFromPointsBatch(len(views), views.ctypes.data, len(mins), ents.ctypes.data, bits.ctypes.data, leafs.ctypes.data) == 0
        got = np.unpackbits(bits.view(np.uint8).reshape(len(views), -1), axis=1, bitorder="little")[:, :len(mins)].astype(bool)
        assert np.array_equal(leafs, o.point_leafnum(views))
        assert np.array_equal(got, o.visible_from_points(views, ents21, flood, nc, cb, vis)), f"step {step}"
    # close what the script left open so that the module's other tests see a map with all portals closed
    for a in range(nA):
        for b in range(a, nA):
            for _ in range(int(portals[a, b]) // (2 if a == b else 1)):
                lib.CM_AdjustAreaPortalState(a, b, 0)
