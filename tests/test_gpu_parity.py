"""GPU parity: the CUDA engine, called through the C ABI, against the oracle (and the golden vectors).

Bar: byte-identical 56-byte trace_t records and identical contents words -- which implies the stated
tolerance (classification bit-exact, fraction/endpos within 1e-5 relative) with room to spare."""
import numpy as np
import pytest

import cases
from helpers import (GOLDEN_SETS, assert_traces_equal, bspgen, cmoracle, comb_map, flat_for, get_map, load_golden)

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from threecore_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


_loaded = {}


def load(eng, kind):
    """upload `kind` (cached flat dict for the oracle)"""
    m = get_map(kind)
    fm = eng.load_bsp(m.data)
    if kind not in _loaded:
        _loaded[kind] = cmoracle.OracleCM(fm.to_dict())
    return m, _loaded[kind]


def check(eng, oracle, case):
    got = cases.run_case(eng, case)
    want = cases.run_case(oracle, case)
    if case["kind"] == "points":
        assert np.array_equal(got, want), f"{case['name']}: {(got != want).sum()} contents words differ"
    else:
        assert_traces_equal(got, want, case["name"])
    return want


@pytest.mark.parametrize("kind", ["tiny", "room", "arena_small", "arena_models", "patches_small"])
def test_world_traces(eng, kind):
    m, o = load(eng, kind)
    launches = eng.launch_count
    for case in cases.world_cases(kind, n=60000):
        check(eng, o, case)
    assert eng.launch_count > launches, "no kernel launched: not the CUDA path"


@pytest.mark.parametrize("kind", ["arena_models", "patches_small"])
def test_entity_traces_and_point_contents(eng, kind):
    m, o = load(eng, kind)
    for case in cases.entity_cases(kind, n=60000):
        check(eng, o, case)


@pytest.mark.parametrize("name", GOLDEN_SETS)
def test_golden_vectors(eng, name):
    """fixtures produced by the unmodified reference (tests/golden/make_golden.py)"""
    bsp, items = load_golden(name)
    eng.load_bsp(bsp)
    for case, want in items:
        got = cases.run_case(eng, case)
        if case["kind"] == "points":
            assert np.array_equal(got, want), case["name"]
        else:
            assert_traces_equal(got, want, case["name"])


def test_hit_ids_match_oracle(eng):
    m, o = load(eng, "patches_small")
    s, e, mins, maxs = bspgen.make_patch_sweeps(m, 50000, 5)
    got, hit = eng.box_trace(s, e, mins, maxs, mask=bspgen.MASK_ALL, want_hit=True)
    want, info = o.box_trace(s, e, mins, maxs, mask=bspgen.MASK_ALL, want_info=True)
    assert_traces_equal(got, want, "patch sweeps")
    want_hit = np.where(info["hitPatch"] >= 0, -2 - info["hitPatch"], info["hitBrush"])
    assert np.array_equal(hit, want_hit)
    assert (hit <= -2).mean() > 0.1, "patch hits are not being exercised"


@pytest.mark.parametrize("cvars", [dict(noCurves=1, playerCurveClip=1), dict(noCurves=0, playerCurveClip=0)])
def test_patch_cvars(eng, cvars):
    m, _ = load(eng, "patches_small")
    o = cmoracle.OracleCM(eng.map.to_dict(), **cvars)
    eng.set_cvars(**cvars)
    try:
        s, e, mins, maxs = bspgen.make_patch_sweeps(m, 30000, 6)
        got = eng.box_trace(s, e, mins, maxs, mask=bspgen.MASK_ALL)
        want = o.box_trace(s, e, mins, maxs, mask=bspgen.MASK_ALL)
        assert_traces_equal(got, want, f"cvars {cvars}")
    finally:
        eng.set_cvars(0, 1)


def test_counters_match_oracle_visit_counts(eng):
    """the instrumented kernel variant visits exactly what the oracle visits (drives the L2 roofline model)"""
    m, o = load(eng, "arena_small")
    s, e = bspgen.make_sweeps(m, 40000, 8)
    eng.enable_counters(True)
    try:
        eng.read_counters(reset=True)
        got = eng.box_trace(s, e, bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, mask=bspgen.MASK_PLAYERSOLID)
        c = eng.read_counters()
    finally:
        eng.enable_counters(False)
    want, info = o.box_trace(s, e, bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, mask=bspgen.MASK_PLAYERSOLID, want_info=True)
    assert_traces_equal(got, want, "counting variant")
    assert c["traces"] == len(s)
    for k in ("nodes", "leafs", "brushRefs", "splits"):
        assert c[k] == int(info[k].sum()), (k, c[k], int(info[k].sum()))


def test_edge_batches(eng):
    m, o = load(eng, "room")
    # empty batch
    out = eng.box_trace(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32))
    assert out.shape == (0,)
    # ragged sizes around the warp/chunk boundaries
    s, e = bspgen.make_sweeps(m, 1025, 3)
    for n in (1, 2, 31, 32, 33, 255, 256, 257, 1025):
        got = eng.box_trace(s[:n], e[:n], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, mask=bspgen.MASK_PLAYERSOLID)
        want = o.box_trace(s[:n], e[:n], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, mask=bspgen.MASK_PLAYERSOLID)
        assert_traces_equal(got, want, f"n={n}")


def test_error_codes(eng):
    from threecore_b200.engine import CMGPUError, Engine
    m, _ = load(eng, "arena_models")
    s, e = bspgen.make_sweeps(m, 16, 1)
    with pytest.raises(CMGPUError) as ex:
        eng.box_trace(s, e, model=np.full(16, 999, np.int32))          # CM_ClipHandleToModel: bad handle
    assert ex.value.code == 5
    with pytest.raises(CMGPUError) as ex:
        eng.box_trace(s, e, model=np.full(16, -1, np.int32))
    assert ex.value.code == 5
    e2 = Engine(0)
    with pytest.raises(CMGPUError) as ex:
        e2.box_trace(s, e)                                             # no map uploaded
    assert ex.value.code == 4
    e2.close()


@pytest.mark.parametrize("slabs", [50, 150])
def test_deep_unbalanced_trees(eng, slabs):
    """The reference recurses without a limit (cm_trace.c:516-537); here every traversal stack is sized by the depth of
    the uploaded tree.  A comb of 50 slabs is deeper than the 32 pending halves the pool kernel once held, one of 150
    deeper than the 64 entries of the kernels with local-memory stacks (they have a 256-entry variant).  Sweeps along
    the comb in both directions keep a pending half per level; every kernel variant must return the reference's bytes."""
    import torch
    from helpers import cmref
    from threecore_b200.engine import Engine
    m, half = comb_map(slabs)
    assert m.stats["max_depth"] > (64 if slabs > 100 else 32)
    e2 = Engine(0)
    try:
        e2.load_bsp(m.data)
        o = cmoracle.OracleCM(e2.map.to_dict())
        rng = np.random.default_rng(slabs)
        n = 1 << 20                                       # big enough for the binned pool kernel
        s = np.stack([rng.uniform(-half + 70, half - 70, n), rng.uniform(-20, 20, n), rng.uniform(-20, 20, n)], 1).astype(np.float32)
        en = s.copy()
        en[:, 0] = np.where(rng.random(n) < 0.5, half - 70, -half + 70).astype(np.float32)      # to either end of the comb
        en[: n // 8] = s[: n // 8]                         # position tests
        idx = rng.choice(n, 4000, replace=False)
        # masks: solid (stops at the first slab) and one that no brush has (the sweep crosses the whole comb: deepest stacks)
        for mask in (bspgen.MASK_PLAYERSOLID, 0x40000000):
            for mins, maxs in ((bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS), (None, None)):
                want = o.box_trace(s[idx], en[idx], mins, maxs, mask=mask)
                if cmref.available():
                    assert_traces_equal(cmref.RefCM().load(m.data).box_trace(s[idx], en[idx], mins, maxs, mask=mask), want, "oracle vs reference")
                for opts in (dict(), dict(pool=0), dict(fast=0)):
                    try:
                        for k, v in opts.items():
                            e2.set_option(k, v)
                        got = e2.box_trace(s[idx], en[idx], mins, maxs, mask=mask)             # small launch
                        assert_traces_equal(got, want, f"{slabs} slabs, small launch, {opts}, mask {mask:#x}")
                        dev = torch.device("cuda", 0)
                        out = torch.empty((n, 56), dtype=torch.uint8, device=dev)
                        e2.box_trace_device(torch.from_numpy(s).to(dev), torch.from_numpy(en).to(dev), mins, maxs, mask, out)
                        torch.cuda.synchronize()
                        e2.check_status()
                        from threecore_b200.engine import traces_from_tensor
                        assert_traces_equal(traces_from_tensor(out)[idx], want, f"{slabs} slabs, 1 Mi launch, {opts}, mask {mask:#x}")
                    finally:
                        for k in opts:
                            e2.set_option(k, 1)
        # a box over the whole comb: CM_BoxLeafnums_r keeps a pending child per level too (cm_test.c:131-156)
        lists, counts, last = e2.box_leafnums(np.array([[-half, -40, -40]], np.float32), np.array([[half, 40, 40]], np.float32), 1024)
        wl, wc, wlast = o.box_leafnums(np.array([[-half, -40, -40]], np.float32), np.array([[half, 40, 40]], np.float32), 1024)
        assert counts[0] == wc[0] and np.array_equal(lists[0][: counts[0]], wl[0][: wc[0]]) and last[0] == wlast[0]
        assert counts[0] > slabs
    finally:
        e2.close()


def test_tree_deeper_than_the_stacks_is_refused_at_upload():
    """256+ levels: cmgpu_upload says so (CMGPU_ERR_LIMIT) instead of tracing with stacks that could overflow"""
    from threecore_b200.engine import CMGPUError, Engine
    m, _ = comb_map(300)
    assert m.stats["max_depth"] >= 256
    e2 = Engine(0)
    try:
        with pytest.raises(CMGPUError) as ex:
            e2.load_bsp(m.data)
        assert ex.value.code == 8 and "levels deep" in str(ex.value)
    finally:
        e2.close()


def test_bad_model_handle_in_a_device_array_is_reported(eng):
    """CM_ClipHandleToModel drops the frame on a bad handle (cm_load.c:768-798); the host-pointer calls check before they
    launch, a device array is checked by the kernels: no out-of-range read, the item hits nothing, check_status says why"""
    import torch
    from threecore_b200.engine import CMGPUError, traces_from_tensor
    m, o = load(eng, "arena_models")
    n = 4096
    s, e = bspgen.make_sweeps(m, n, 21)
    dev = torch.device("cuda", 0)
    models = np.zeros(n, np.int32)
    models[5], models[77], models[4000] = 999, -3, 1 << 30
    ds, de = torch.from_numpy(s).to(dev), torch.from_numpy(e).to(dev)
    out = torch.empty((n, 56), dtype=torch.uint8, device=dev)
    eng.box_trace_device(ds, de, bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, out,
                         models=torch.from_numpy(models).to(dev))
    torch.cuda.synchronize()
    with pytest.raises(CMGPUError) as ex:
        eng.check_status()
    assert ex.value.code == 5
    got = traces_from_tensor(out)
    ok = np.ones(n, bool)
    ok[[5, 77, 4000]] = False
    want = o.box_trace(s[ok], e[ok], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, mask=bspgen.MASK_PLAYERSOLID)
    assert_traces_equal(got[ok], want, "items next to bad handles")
    assert (got["fraction"][~ok] == 1.0).all() and (got["contents"][~ok] == 0).all()
    eng.check_status()                                    # the status word was cleared
    pc = torch.empty(n, dtype=torch.int32, device=dev)
    eng.point_contents_device(ds, pc, models=torch.from_numpy(models).to(dev))
    torch.cuda.synchronize()
    with pytest.raises(CMGPUError) as ex:
        eng.check_status()
    assert ex.value.code == 5 and (pc.cpu().numpy()[~ok] == 0).all()


# ---- full-size properties (BASELINE.json sizes; the oracle only checks a sample) -----------------
@pytest.fixture(scope="module")
def big(eng):
    import torch
    m = bspgen.arena_map(n_brushes=20000, seed=0xC0FFEE02)
    n = 1 << 22
    s, e = bspgen.make_sweeps(m, n, 0xBADC0DE2, z_range=(24.0, 2048.0))
    eng.load_bsp(m.data)
    dev = torch.device("cuda", 0)
    ds, de = torch.from_numpy(s).to(dev), torch.from_numpy(e).to(dev)
    out = torch.empty((n, 56), dtype=torch.uint8, device=dev)
    eng.box_trace_device(ds, de, bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, out)
    torch.cuda.synchronize()
    eng.check_status()
    return dict(m=m, s=s, e=e, ds=ds, de=de, out=out, n=n)


def test_full_size_sample_against_oracle(eng, big):
    from threecore_b200.engine import traces_from_tensor
    idx = np.random.default_rng(1).choice(big["n"], size=150000, replace=False)
    got = traces_from_tensor(big["out"])[idx]
    o = cmoracle.OracleCM(eng.map.to_dict())
    want = o.box_trace(big["s"][idx], big["e"][idx], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, mask=bspgen.MASK_PLAYERSOLID)
    assert_traces_equal(got, want, "4Mi-ray batch, 150k sample")


def test_full_size_permutation_invariance(eng, big):
    """results belong to items, not to lanes: a permuted batch gives the permuted records (scheduler independence)"""
    import torch
    n = big["n"]
    perm = torch.randperm(n, device=big["ds"].device, generator=torch.Generator(device=big["ds"].device).manual_seed(3))
    out2 = torch.empty_like(big["out"])
    eng.box_trace_device(big["ds"][perm].contiguous(), big["de"][perm].contiguous(), bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS,
                         bspgen.MASK_PLAYERSOLID, out2)
    torch.cuda.synchronize()
    assert torch.equal(out2, big["out"][perm])


def test_full_size_host_api_equals_device_api(eng, big):
    from threecore_b200.engine import traces_from_tensor
    n = 1 << 20
    got = eng.box_trace(big["s"][:n], big["e"][:n], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, mask=bspgen.MASK_PLAYERSOLID)
    assert_traces_equal(got, traces_from_tensor(big["out"][:n]), "host-pointer path vs device-pointer path")


def test_full_size_trace_invariants(eng, big):
    """size-independent properties of trace_t over the whole batch"""
    from threecore_b200.engine import traces_from_tensor
    t = traces_from_tensor(big["out"])
    f = t["fraction"]
    assert ((f >= 0) & (f <= 1)).all()
    assert (t["entityNum"] == 0).all() and (t["plane_pad"] == 0).all()
    assert ((t["allsolid"] == 0) | (t["startsolid"] == 1)).all()            # allsolid implies startsolid
    assert (f[t["allsolid"] == 1] == 0).all()
    miss = f == 1
    assert np.array_equal(t["endpos"][miss], big["e"][miss])                 # untouched sweeps end at `end`
    hit = (f < 1) & (t["allsolid"] == 0)
    nrm = np.linalg.norm(t["plane_normal"][hit].astype(np.float64), axis=1)
    assert np.abs(nrm - 1).max() < 1e-3                                       # the assert the reference compiles out (cm_trace.c:683)
    assert (t["contents"][hit] & bspgen.MASK_PLAYERSOLID != 0).all()
    # endpos lies on the segment at `fraction`
    want = big["s"][hit] + f[hit, None] * (big["e"][hit] - big["s"][hit])
    assert np.array_equal(t["endpos"][hit], want.astype(np.float32))
    # idempotence: re-tracing to the reported endpos is not blocked earlier than the push-off epsilon allows
    assert 0.3 < hit.mean() < 0.7


def test_binned_order_equals_input_order(eng, big):
    """tracing a batch in start-cell order (cmgpu_set_option "bin") must not change a single byte"""
    import torch
    n = 1 << 20
    outs = []
    for on in (1, 0):
        eng.set_option("bin", on)
        o = torch.empty((n, 56), dtype=torch.uint8, device=big["ds"].device)
        eng.box_trace_device(big["ds"][:n], big["de"][:n], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, o)
        torch.cuda.synchronize()
        outs.append(o)
    eng.set_option("bin", 1)
    assert torch.equal(outs[0], outs[1])
    assert torch.equal(outs[0], big["out"][:n])


@pytest.mark.parametrize("opts", [dict(pool=0), dict(pool=0, slab=0), dict(fast=0), dict(pool=1, slab=0), dict(pool=1, pool_reps=5, gang=8),
                                  dict(pool=0, simple_max_items=1 << 21), dict(pool=0, simple_max_items=1 << 21, bin=0, slab=0),
                                  dict(two_phase=0), dict(entry_grid=0), dict(node_thresholds=0), dict(node_thresholds=0, compact_by_item=0), dict(two_phase=0, pool=0), dict(two_phase_small=1, pool=0), dict(two_phase_small=1, pool=0, simple_max_items=1 << 21)])
def test_kernel_variants_agree(eng, big, opts):
    """the general kernel, the one-item-per-lane hot kernel and the pool kernel must return the same bytes"""
    import torch
    n = 1 << 20
    try:
        for k, v in opts.items():
            eng.set_option(k, v)
        o = torch.empty((n, 56), dtype=torch.uint8, device=big["ds"].device)
        eng.box_trace_device(big["ds"][:n], big["de"][:n], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, o)
        torch.cuda.synchronize()
        eng.check_status()
    finally:
        for k, v in dict(pool=1, fast=1, slab=1, pool_reps=1, gang=4, gang_pool=3, simple_max_items=786432, bin=1, two_phase=1, two_phase_small=0, node_thresholds=1, compact_by_item=1, entry_grid=1).items():
            eng.set_option(k, v)
    assert torch.equal(o, big["out"][:n])


@pytest.mark.parametrize("n,shift", [((1 << 20) + 37, 0), ((1 << 20) - 255, 1), ((1 << 20) + 1, 3)])
def test_two_phase_results_odd_sizes_and_unaligned_output(eng, big, n, shift):
    """binned hot launches write 16-byte results by slot and expand them to records in a second pass: item counts that
    are no multiple of the pass's block, hit ids, and an output pointer the pass cannot take (8 mod 16: direct stores)"""
    import torch
    dev = big["ds"].device
    buf = torch.zeros((n + 4) * 56, dtype=torch.uint8, device=dev)
    o = buf[shift * 56:(shift + n) * 56].view(n, 56)
    hit = torch.full((n,), -7, dtype=torch.int32, device=dev)
    eng.box_trace_device(big["ds"][:n], big["de"][:n], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, o, hit_ids=hit)
    torch.cuda.synchronize()
    eng.check_status()
    assert torch.equal(o, big["out"][:n])
    assert not bool(buf[:shift * 56].any()) and not bool(buf[(shift + n) * 56:].any()), "wrote outside the records"
    eng.set_option("two_phase", 0)
    try:
        hit0 = torch.full((n,), -7, dtype=torch.int32, device=dev)
        o0 = torch.empty((n, 56), dtype=torch.uint8, device=dev)
        eng.box_trace_device(big["ds"][:n], big["de"][:n], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, o0, hit_ids=hit0)
        torch.cuda.synchronize()
    finally:
        eng.set_option("two_phase", 1)
    assert torch.equal(hit, hit0) and torch.equal(o0, o)


@pytest.mark.parametrize("which", ["pool", "persistent"])
def test_binning_small_batches_all_query_kinds(eng, which):
    """force the binned path and the big-launch kernels on small batches of every query kind (per-item hulls, models,
    temp boxes, rotations); by default launches this small run the thread-per-item kernel in input order"""
    eng.set_option("bin_min_items", 1)
    eng.set_option("pool_min_items", 1 if which == "pool" else 1 << 30)
    eng.set_option("simple_max_items", 0)
    try:
        for kind in ("arena_models", "patches_small"):
            m, o = load(eng, kind)
            for case in cases.world_cases(kind, n=5000) + cases.entity_cases(kind, n=5000):
                check(eng, o, case)
    finally:
        eng.set_option("bin_min_items", 98304)
        eng.set_option("pool_min_items", 196609)
        eng.set_option("simple_max_items", 786432)


def test_pool_kernel_counters_and_hit_ids(eng, big):
    """the pool kernel's counting variant visits what the oracle visits, and reports the brush that set each result"""
    import torch
    n = 1 << 20                                   # above pool_min_items: this launch runs tracePoolKernel
    eng.load_bsp(big["m"].data)                   # earlier tests left another map in the engine
    o = cmoracle.OracleCM(eng.map.to_dict())
    hit = torch.empty(n, dtype=torch.int32, device=big["ds"].device)
    out = torch.empty((n, 56), dtype=torch.uint8, device=big["ds"].device)
    eng.enable_counters(True)
    try:
        eng.read_counters(reset=True)
        eng.box_trace_device(big["ds"][:n], big["de"][:n], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, out,
                             hit_ids=hit)
        torch.cuda.synchronize()
        c = eng.read_counters()
    finally:
        eng.enable_counters(False)
    assert torch.equal(out, big["out"][:n])
    idx = np.sort(np.random.default_rng(2).choice(n, size=200000, replace=False))
    want, info = o.box_trace(big["s"][idx], big["e"][idx], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS,
                             mask=bspgen.MASK_PLAYERSOLID, want_info=True)
    assert np.array_equal(hit.cpu().numpy()[idx], info["hitBrush"])
    assert c["traces"] == n
    # totals over the whole launch against the sample's means (the oracle walks 200k of the 1Mi items)
    for k in ("nodes", "leafs", "brushRefs", "splits"):
        per_item = c[k] / n
        assert abs(per_item - info[k].mean()) < 0.02 * info[k].mean() + 0.05, (k, per_item, info[k].mean())


def test_unknown_option_is_refused(eng):
    from threecore_b200.engine import CMGPUError
    with pytest.raises(CMGPUError) as ex:
        eng.set_option("no_such_knob", 1)
    assert ex.value.code == 1


def test_node_tables_per_hull_and_per_map(eng, big):
    """the pool kernel keeps one node threshold table per (map, hull): more hulls than the cache holds, the first hull
    again after it was dropped, a hull with a zero extent, and another map in between -- always the general kernel's bytes"""
    import torch
    n = 1 << 20
    dev = big["ds"].device
    eng.load_bsp(big["m"].data)
    rng = np.random.default_rng(77)
    hulls = [(bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS), (None, None),
             (np.array([-8, -8, 0], np.float32), np.array([8, 8, 0], np.float32))]
    hulls += [(-rng.uniform(1, 40, 3).astype(np.float32), rng.uniform(1, 40, 3).astype(np.float32)) for _ in range(8)]
    hulls.append(hulls[0])

    def run(h, **opts):
        for k, v in opts.items():
            eng.set_option(k, v)
        o = torch.empty((n, 56), dtype=torch.uint8, device=dev)
        eng.box_trace_device(big["ds"][:n], big["de"][:n], h[0], h[1], bspgen.MASK_PLAYERSOLID, o)
        torch.cuda.synchronize()
        eng.check_status()
        return o

    try:
        for i, h in enumerate(hulls):
            got = run(h, fast=1)
            want = run(h, fast=0)
            assert torch.equal(got, want), f"hull {i}"
            if i == 5:                                   # another map, then back: the tables of the first load are gone
                m2 = get_map("arena_small")
                eng.load_bsp(m2.data)
                s2, e2 = bspgen.make_sweeps(m2, n, 5)
                a, b = torch.from_numpy(s2).to(dev), torch.from_numpy(e2).to(dev)
                o1 = torch.empty((n, 56), dtype=torch.uint8, device=dev)
                o2 = torch.empty((n, 56), dtype=torch.uint8, device=dev)
                eng.set_option("fast", 1)
                eng.box_trace_device(a, b, h[0], h[1], bspgen.MASK_PLAYERSOLID, o1)
                eng.set_option("fast", 0)
                eng.box_trace_device(a, b, h[0], h[1], bspgen.MASK_PLAYERSOLID, o2)
                torch.cuda.synchronize()
                assert torch.equal(o1, o2), "second map"
                eng.load_bsp(big["m"].data)
    finally:
        eng.set_option("fast", 1)
