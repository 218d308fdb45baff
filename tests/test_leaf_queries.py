"""CM_PointLeafnum / CM_BoxLeafnums (cm_test.c:31-181), the PVS part of SV_LinkEntity (server/sv_world.c:255-303),
area portals (cm_test.c:322-470) and the PVS stage of SV_AddEntitiesVisibleFromPoint (server/sv_snapshot.c:316-355).

CPU part: the restatement in oracle/cm_oracle.c against the UNMODIFIED reference (libcmref.so: cm_test.c and
sv_world.c compiled where they lie) and against the committed golden set; the product's loader against the
reference's view of the same lump.  GPU part (-m gpu): the device kernels, through the C ABI, against both."""
import numpy as np
import pytest

import leaf_cases as lc
from helpers import GOLDEN, cmoracle, cmref, flat_for, have_ref

needs_ref = pytest.mark.skipif(not have_ref(), reason="oracle/_ref/libcmref.so not built (no reference tree here)")

KINDS = ["tiny_vis", "room_vis", "arena_small_vis", "arena_small_novis"]


def _oracle(kind):
    m = lc.vis_map(kind)
    return m, cmoracle.OracleCM(flat_for(m.data))


def _ent_records(r, m, n, seed):
    """link n box entities in the reference's world and read back what SV_LinkEntity derived"""
    r.sv_clear_world()
    mins, maxs, origin = lc.entities(m, n, seed)
    zero = np.zeros(3, np.float32)
    absmin, absmax, recs = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32), np.zeros((n, 20), np.int32)
    for i in range(n):
        r.sv_link_entity(i, False, 0, mins[i], maxs[i], origin[i], zero, 1)
        absmin[i], absmax[i], _ = r.sv_entity_bounds(i)
        recs[i], _ = r.sv_entity_clusters(i)
    return absmin, absmax, recs


def _cmp_ent(got21, want20, what):
    """got: numLeafs, areanum, areanum2, numClusters, lastCluster, clusters[16]; want: the svEntity_t fields"""
    got, want = np.asarray(got21), np.asarray(want20)
    assert np.array_equal(got[:, 1:5], want[:, 0:4]), f"{what}: areanum/areanum2/numClusters/lastCluster differ"
    for i in range(len(got)):
        k = got[i, 3]
        assert np.array_equal(got[i, 5:5 + k], want[i, 4:4 + k]), f"{what}: clusternums of entity {i}"


# ---------------------------------------------------------------- CPU: oracle vs the reference
@needs_ref
@pytest.mark.parametrize("kind", KINDS)
def test_oracle_leaf_lists_match_reference(kind):
    m, o = _oracle(kind)
    r = cmref.RefCM().load(m.data)
    pts = lc.viewpoints(m, 4000, 11)
    assert np.array_equal(o.point_leafnum(pts), r.point_leafnum(pts))
    mins, maxs = lc.query_boxes(m, 3000, 12)
    overflowed = 0
    for listsize in (1, 8, 128, 1024):
        a = r.box_leafnums(mins, maxs, listsize)
        b = o.box_leafnums(mins, maxs, listsize)
        assert np.array_equal(a[1], b[1]) and np.array_equal(a[2], b[2]), f"listsize {listsize}: counts / lastLeaf"
        assert np.array_equal(a[0], b[0]), f"listsize {listsize}: lists"
        overflowed += int((a[1] == listsize).sum())
    assert overflowed > 0


@needs_ref
@pytest.mark.parametrize("kind", KINDS)
def test_oracle_entity_clusters_match_sv_link_entity(kind):
    m, o = _oracle(kind)
    r = cmref.RefCM().load(m.data)
    absmin, absmax, want = _ent_records(r, m, 400, 21)
    got = o.entity_clusters(absmin, absmax)
    _cmp_ent(got, want, kind)
    if kind.startswith("arena"):
        assert (want[:, 2] == 16).any() and (want[:, 3] != 0).any(), "no entity overflowed its cluster list"
        assert (want[:, 1] != -1).any(), "no entity touched two areas"
        assert (got[:, 0] == 128).any(), "no entity overflowed the 128-leaf list"


@needs_ref
@pytest.mark.parametrize("kind", ["room_vis", "arena_small_vis", "arena_small_novis"])
def test_oracle_areas_and_visibility_match_reference(kind):
    m, o = _oracle(kind)
    r = cmref.RefCM().load(m.data)
    nA = m.stats["areas"]
    nc, cb, vis = lc.vis_lump(m.data)
    absmin, absmax, _ = _ent_records(r, m, 150, 31)
    ents = o.entity_clusters(absmin, absmax)
    views = lc.viewpoints(m, 48, 32)
    portals = np.zeros((nA, nA), np.int32)
    a1, a2 = np.divmod(np.arange(nA * nA), nA)
    seen = 0
    for step, (p, q, open_) in enumerate([(-1, -1, True)] + lc.portal_script(nA, 33)):
        r.adjust_area_portal_state(p, q, open_)
        if p >= 0:
            d = 1 if open_ else -1
            portals[p, q] += d
            portals[q, p] += d
        flood = o.flood_areas(portals)
        want_conn = r.areas_connected(a1, a2).astype(bool)
        assert np.array_equal(flood[a1] == flood[a2], want_conn), f"step {step}: CM_AreasConnected"
        if step % 4 == 0:
            want = r.sv_visible_from_points(views, np.arange(len(ents)))
            got = o.visible_from_points(views, ents, flood, nc if vis is not None else r.num_clusters(), cb, vis)
            assert np.array_equal(got, want), f"step {step}: {(got != want).sum()} verdicts differ"
            seen += int(want.sum())
    assert seen > 0


# ---------------------------------------------------------------- CPU: golden set, loader
def _golden():
    z = np.load(f"{GOLDEN}/leaf_queries.npz", allow_pickle=False)
    return {k: z[k] for k in z.files}


def _golden_inputs(g):
    m = lc.vis_map(str(g["kind"]))
    return m, g["pts"], g["mins"], g["maxs"], g["absmin"], g["absmax"], g["views"]


def test_golden_map_is_the_generated_one():
    g = _golden()
    m = lc.vis_map(str(g["kind"]))
    import zlib
    assert zlib.crc32(m.data) == int(g["bsp_crc32"]), "bspgen no longer generates the map the golden set was made on"


def test_oracle_reproduces_golden_leaf_queries():
    g = _golden()
    m, pts, mins, maxs, absmin, absmax, views = _golden_inputs(g)
    o = cmoracle.OracleCM(flat_for(m.data))
    assert np.array_equal(o.point_leafnum(pts), g["leafnums"])
    lists, counts, last = o.box_leafnums(mins, maxs, int(g["listsize"]))
    assert np.array_equal(lists, g["lists"]) and np.array_equal(counts, g["counts"]) and np.array_equal(last, g["lastLeafs"])
    ents = o.entity_clusters(absmin, absmax)
    _cmp_ent(ents, g["ent_records"], "golden")
    nc, cb, vis = lc.vis_lump(m.data)
    flood = o.flood_areas(g["portals"])
    assert np.array_equal(o.visible_from_points(views, ents, flood, nc, cb, vis), g["visible"])


def test_loader_reads_clusters_areas_and_visibility():
    from threecore_b200.engine import FlatMap
    for kind in ("room_vis", "arena_small_novis"):
        m = lc.vis_map(kind)
        d = FlatMap.from_bsp(m.data).to_dict()
        nc, cb, vis = lc.vis_lump(m.data)
        assert d["num_areas"] == m.stats["areas"]
        assert d["num_clusters"] == m.stats["clusters"]
        if vis is None:
            assert d["visibility"].size == 0 and d["cluster_bytes"] == (m.stats["clusters"] + 31) & ~31   # cm_load.c:503
        else:
            assert d["cluster_bytes"] == cb and np.array_equal(d["visibility"][:nc * cb], vis)
        if have_ref():
            r = cmref.RefCM().load(m.data)
            assert r.num_clusters() == d["num_clusters"]
            c, a = r.leaf_cluster_area(np.arange(len(d["leafs"])))
            assert np.array_equal(np.stack([c, a], 1), d["leaf_cluster_area"])
            if vis is not None:
                for cl in (0, 7, nc - 1):
                    assert np.array_equal(r.cluster_pvs(cl, cb), d["visibility"][cl * cb:(cl + 1) * cb])


# ---------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def engine():
    from threecore_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


def _upload(engine, kind):
    m = lc.vis_map(kind)
    engine.upload(engine.load_bsp(m.data))
    return m, cmoracle.OracleCM(flat_for(m.data))


@pytest.mark.gpu
@pytest.mark.parametrize("kind", KINDS)
def test_gpu_leaf_lists(engine, kind):
    m, o = _upload(engine, kind)
    pts = lc.viewpoints(m, 200000, 41)
    assert np.array_equal(engine.point_leafnum(pts)[:20000], o.point_leafnum(pts[:20000]))
    mins, maxs = lc.query_boxes(m, 20000, 42)
    for listsize in (0, 1, 8, 128, 700):
        want = o.box_leafnums(mins, maxs, listsize)
        got = engine.box_leafnums(mins, maxs, listsize)
        assert np.array_equal(got[1], want[1]) and np.array_equal(got[2], want[2]), f"listsize {listsize}: counts / lastLeaf"
        assert np.array_equal(got[0], want[0]), f"listsize {listsize}: lists"


@pytest.mark.gpu
@pytest.mark.parametrize("kind", KINDS)
def test_gpu_entity_clusters(engine, kind):
    m, o = _upload(engine, kind)
    mins, maxs = lc.query_boxes(m, 50000, 43)
    want = o.entity_clusters(mins, maxs)
    got = engine.entity_clusters(mins, maxs)
    flat = np.concatenate([got[f].reshape(len(got), -1) for f in got.dtype.names], axis=1)
    assert np.array_equal(flat, want)


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["room_vis", "arena_small_vis", "arena_small_novis"])
def test_gpu_areas_and_visibility(engine, kind):
    m, o = _upload(engine, kind)
    nA = m.stats["areas"]
    nc, cb, vis = lc.vis_lump(m.data)
    if vis is None:
        nc = m.stats["clusters"]
    mins, maxs = lc.query_boxes(m, 333, 44)          # not a multiple of 32: the last word is partial
    ents = engine.entity_clusters(mins, maxs)
    ents21 = o.entity_clusters(mins, maxs)
    views = lc.viewpoints(m, 64, 45)
    portals = np.zeros((nA, nA), np.int32)
    for step, (p, q, open_) in enumerate([(-1, -1, True)] + lc.portal_script(nA, 46)):
        engine.adjust_area_portal_state(p, q, open_)
        if p >= 0:
            d = 1 if open_ else -1
            portals[p, q] += d
            portals[q, p] += d
        flood = o.flood_areas(portals)
        for a in range(nA):
            for b in range(nA):
                assert engine.areas_connected(a, b) == bool(flood[a] == flood[b])
            bits = np.frombuffer(engine.write_area_bits(a), np.uint8)
            assert np.array_equal(np.unpackbits(bits, bitorder="little")[:nA].astype(bool), flood == flood[a])
        if step % 4 == 0:
            got, leafs = engine.visible_from_points(views, ents)
            assert np.array_equal(leafs, o.point_leafnum(views))
            want = o.visible_from_points(views, ents21, flood, nc, cb, vis)
            assert np.array_equal(got, want), f"step {step}: {(got != want).sum()} verdicts differ"
    engine.set_no_areas(True)
    got, _ = engine.visible_from_points(views, ents)
    want = o.visible_from_points(views, ents21, np.zeros(nA, np.int32), nc, cb, vis)
    # cm_noAreas: every pair connected, also from a viewpoint in a solid leaf (area -1) -- cm_test.c:405-409
    lca = o.flat["leaf_cluster_area"]
    inside = lca[o.point_leafnum(views), 1] >= 0
    assert np.array_equal(got[inside], want[inside])
    engine.set_no_areas(False)
    with pytest.raises(Exception):
        engine.adjust_area_portal_state(nA, 0, True)


@pytest.mark.gpu
def test_gpu_golden_leaf_queries(engine):
    g = _golden()
    m, pts, mins, maxs, absmin, absmax, views = _golden_inputs(g)
    engine.upload(engine.load_bsp(m.data))
    assert np.array_equal(engine.point_leafnum(pts), g["leafnums"])
    lists, counts, last = engine.box_leafnums(mins, maxs, int(g["listsize"]))
    assert np.array_equal(lists, g["lists"]) and np.array_equal(counts, g["counts"]) and np.array_equal(last, g["lastLeafs"])
    ents = engine.entity_clusters(absmin, absmax)
    flat = np.concatenate([ents[f].reshape(len(ents), -1) for f in ents.dtype.names], axis=1)
    _cmp_ent(flat, g["ent_records"], "golden")
    nA = g["portals"].shape[0]
    for a in range(nA):
        for b in range(a, nA):
            for _ in range(int(g["portals"][a, b]) // (2 if a == b else 1)):
                engine.adjust_area_portal_state(a, b, True)
    got, _ = engine.visible_from_points(views, ents)
    assert np.array_equal(got, g["visible"])
