"""botlib's AAS tracer: AAS_TraceClientBBox / AAS_PointAreaNum.

CPU: the restatement (oracle/cm_oracle.c) against the unmodified botlib/be_aas_sample.c in libcmref.so and against the
committed fixture the reference produced.  GPU: the kernels, through the C ABI, against both -- all 36 bytes of every
aas_trace_t."""
import os
from types import SimpleNamespace

import numpy as np
import pytest

from helpers import GOLDEN, cmoracle, cmref, have_ref
from threecore_b200 import aasgen


def _golden():
    z = np.load(os.path.join(GOLDEN, "aas_trace.npz"))
    w = SimpleNamespace(planes=z["planes"], nodes=z["nodes"], presence=z["presence"], contents=z["contents"])
    return w, z


def _same_moves(a, b, ra, rb, what):
    a8 = np.ascontiguousarray(a).view(np.uint8).reshape(-1, 84)
    b8 = np.ascontiguousarray(b).view(np.uint8).reshape(-1, 84)
    bad = (a8 != b8).any(axis=1) | (np.asarray(ra) != np.asarray(rb))
    assert not bad.any(), (f"{what}: {int(bad.sum())} of {len(a8)} aas_clientmove_t records differ, first at {int(np.flatnonzero(bad)[0])}:\n"
                           f" got  {a[bad][0]} ret {np.asarray(ra)[bad][0]}\n want {b[bad][0]} ret {np.asarray(rb)[bad][0]}")


def _comb_world(n):
    """planes x = n, n-1, ..., 1, each node's front child an area and its back child the next node: a line from x = 0.5 to
    beyond x = n is split at every node and the near part goes on, one more pending far part per level"""
    planes = np.zeros((2 * n, 5), np.float32)
    nodes = np.zeros((n + 1, 3), np.int32)
    for i in range(1, n + 1):
        d = float(n + 1 - i)
        planes[2 * (i - 1)] = [1, 0, 0, d, 0]
        planes[2 * (i - 1) + 1] = [-1, 0, 0, -d, 0]
        nodes[i] = [2 * (i - 1), -1, i + 1 if i < n else -1]
    w = SimpleNamespace(planes=planes, nodes=nodes, presence=np.array([0, 6], np.int32))
    s = np.array([[0.5, 0, 0], [0.5, 0, 0], [60.5, 0, 0]], np.float32)
    e = np.array([[n + 5.0, 0, 0], [100.5, 0, 0], [0.5, 3, 0]], np.float32)     # overflows / 100 pending parts / the other way round
    return w, s, e


def _same(a, b, what):
    a8 = np.ascontiguousarray(a).view(np.uint8).reshape(-1, 36)
    b8 = np.ascontiguousarray(b).view(np.uint8).reshape(-1, 36)
    bad = (a8 != b8).any(axis=1)
    assert not bad.any(), f"{what}: {int(bad.sum())} of {len(a8)} aas_trace_t records differ, first at {int(np.flatnonzero(bad)[0])}: {a[bad][0]} != {b[bad][0]}"


@pytest.fixture(scope="module")
def world():
    return aasgen.make_world(n_boxes=300, seed=0xAA5003, extent=3072.0, height=768.0)


def test_generated_world_is_an_aas_tree(world):
    assert world.nodes[0].tolist() == [0, 0, 0] and world.stats["areas"] > 500 and world.stats["solid"] > 100
    assert (world.planes[0::2, :4] == -world.planes[1::2, :4]).all(), "planes come in opposite pairs"
    assert set(np.unique(world.presence[1:])) <= {aasgen.PRESENCE_CROUCH, aasgen.PRESENCE_NORMAL | aasgen.PRESENCE_CROUCH}


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref/libcmref.so not built")
def test_restatement_matches_the_reference(world):
    r = cmref.RefCM().aas_load(world)
    o = cmoracle.OracleAAS(world)
    s, e = aasgen.make_lines(world, 120000, 5)
    for pt in (aasgen.PRESENCE_NORMAL, aasgen.PRESENCE_CROUCH, aasgen.PRESENCE_NONE):
        _same(o.trace_client_bbox(s, e, pt), r.aas_trace_client_bbox(s, e, pt), f"presence {pt}")
    per_item = np.random.default_rng(1).choice([2, 4, 6], size=len(s)).astype(np.int32)
    got = o.trace_client_bbox(s, e, per_item)
    _same(got, r.aas_trace_client_bbox(s, e, per_item), "per-item presence types")
    assert 0.2 < (got["fraction"] < 1).mean() < 0.9 and got["startsolid"].mean() > 0.05 and (got["area"] > 0).any()
    assert np.array_equal(o.point_area_num(s), r.aas_point_area_num(s))


def test_restatement_reproduces_the_golden_fixture():
    w, z = _golden()
    o = cmoracle.OracleAAS(w)
    got = o.trace_client_bbox(z["starts"], z["ends"], z["presencetypes"])
    _same(got, z["traces"].view(cmref.AAS_TRACE_DTYPE).reshape(-1), "golden")
    assert np.array_equal(o.point_area_num(z["starts"]), z["areas"])


def test_deep_chain_overflows_the_stack_like_the_reference():
    """more than 127 entries: the reference prints 'stack overflow' and returns the trace as it stands (be_aas_sample.c:596-600)"""
    w, s, e = _comb_world(200)
    o = cmoracle.OracleAAS(w)
    got = o.trace_client_bbox(s, e, 2)
    assert got["fraction"][0] == 0 and got["startsolid"][0] == 0 and got["area"][0] == 0, "the first line must overflow the stack"
    assert got["fraction"][1] == 1 and got["fraction"][2] == 1
    if have_ref():
        _same(got, cmref.RefCM().aas_load(w).aas_trace_client_bbox(s, e, 2), "overflow")


# ---- the CUDA kernels -------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def eng():
    from threecore_b200.engine import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.mark.gpu
def test_gpu_traces_equal_reference_and_oracle(eng, world):
    eng.aas_upload(world)
    o = cmoracle.OracleAAS(world)
    r = cmref.RefCM().aas_load(world) if have_ref() else None
    n = 1 << 20
    s, e = aasgen.make_lines(world, n, 7)
    launches = eng.launch_count
    for pt in (aasgen.PRESENCE_NORMAL, aasgen.PRESENCE_CROUCH):
        got = eng.aas_trace_client_bbox(s, e, pt)
        _same(got, o.trace_client_bbox(s, e, pt), f"gpu vs oracle, presence {pt}")
        if r is not None:
            want = cmref.fork_map(n, cmref.AAS_TRACE_DTYPE, lambda lo, hi: r.aas_trace_client_bbox(s[lo:hi], e[lo:hi], pt))
            _same(got, want, f"gpu vs reference, presence {pt}")
    assert eng.launch_count > launches
    per_item = np.random.default_rng(2).choice([1, 2, 4, 6], size=n).astype(np.int32)
    _same(eng.aas_trace_client_bbox(s, e, per_item), o.trace_client_bbox(s, e, per_item), "per-item presence types")
    assert np.array_equal(eng.aas_point_area_num(s), o.point_area_num(s))


@pytest.mark.gpu
def test_gpu_golden_fixture_device_pointers_and_errors(eng):
    import torch
    from threecore_b200.engine import AAS_TRACE_DTYPE, CMGPUError, Engine
    w, z = _golden()
    eng.aas_upload(w)
    dev = torch.device("cuda", 0)
    out = torch.empty((len(z["starts"]), 36), dtype=torch.uint8, device=dev)
    eng.aas_trace_client_bbox_device(torch.from_numpy(z["starts"]).to(dev), torch.from_numpy(z["ends"]).to(dev),
                                     torch.from_numpy(z["presencetypes"]).to(dev), out)
    torch.cuda.synchronize()
    _same(out.cpu().numpy().view(AAS_TRACE_DTYPE).reshape(-1), z["traces"].view(AAS_TRACE_DTYPE).reshape(-1), "golden, device pointers")
    assert np.array_equal(eng.aas_point_area_num(z["starts"]), z["areas"])
    assert eng.aas_trace_client_bbox(np.zeros((0, 3), np.float32), np.zeros((0, 3), np.float32), 2).shape == (0,)
    e2 = Engine(0)
    try:
        with pytest.raises(CMGPUError) as ex:
            e2.aas_trace_client_bbox(z["starts"][:4], z["ends"][:4], 2)              # nothing uploaded
        assert ex.value.code == 4
        bad = SimpleNamespace(planes=w.planes, nodes=w.nodes.copy(), presence=w.presence)
        bad.nodes[5, 1] = 5                                                          # a node that is its own child
        with pytest.raises(CMGPUError) as ex:
            e2.aas_upload(bad)
        assert ex.value.code == 6                                                    # CMGPU_ERR_BAD_MAP
    finally:
        e2.close()


@pytest.mark.gpu
def test_gpu_stack_overflow_like_the_reference(eng):
    w, s, e = _comb_world(200)
    eng.aas_upload(w)
    _same(eng.aas_trace_client_bbox(s, e, 2), cmoracle.OracleAAS(w).trace_client_bbox(s, e, 2), "overflow")


# ---- AAS_PredictClientMovement (be_aas_move.c:495-979) ------------------------------------------------------------
def _move_world():
    from threecore_b200 import bspgen
    w = aasgen.make_world(n_boxes=300, seed=0xAA5003, extent=3072.0, height=768.0)
    m = bspgen.arena_map(n_brushes=1200, seed=0xC0FFEE22, size_xy=3072.0, size_z=768.0)     # the world whose contents the predictor asks for
    return w, m


@pytest.mark.gpu
def test_gpu_movement_prediction_equals_the_reference(eng):
    """every stop event, both presence types, swimming, jumping, crouching, steps, gaps: all 84 bytes of every aas_clientmove_t
    and the return value, against the unmodified be_aas_move.c"""
    if not have_ref():
        pytest.skip("oracle/_ref/libcmref.so not shipped")
    w, m = _move_world()
    eng.load_bsp(m.data)
    eng.aas_upload(w)
    r = cmref.RefCM().load(m.data)
    r.aas_load(w)
    n = 400000
    p = aasgen.make_moves(w, n, 21)
    got, gres = eng.aas_predict_client_movement(p)
    want = cmref.fork_map(n, cmref.AAS_MOVE_DTYPE, lambda lo, hi: r.aas_predict_client_movement({k: v[lo:hi] for k, v in p.items()})[0])
    wres = cmref.fork_map(n, np.int32, lambda lo, hi: r.aas_predict_client_movement({k: v[lo:hi] for k, v in p.items()})[1])
    _same_moves(got, want, gres, wres, "movement prediction")
    seen = set(np.unique(want["stopevent"]).tolist())
    assert {0, 1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 2048, 4096} <= seen, f"stop events exercised: {sorted(seen)}"
    assert (wres == 0).any() and (want["endcontents"] != 0).any() and (want["frames"] > 10).any()


@pytest.mark.gpu
def test_gpu_movement_prediction_golden_and_device_pointers(eng):
    import torch
    from threecore_b200.engine import AAS_MOVE_DTYPE
    w, z = _golden()
    eng.load_bsp(z["bsp"].tobytes())
    eng.aas_upload(w)
    p = {k[3:]: z[k] for k in z.files if k.startswith("mv_")}
    got, res = eng.aas_predict_client_movement(p)
    _same_moves(got, z["moves"].view(AAS_MOVE_DTYPE).reshape(-1), res, z["move_results"], "golden moves")
    dev = torch.device("cuda", 0)
    dp = {k: torch.from_numpy(np.ascontiguousarray(v)).to(dev) for k, v in p.items()}
    out = torch.empty((len(got), 84), dtype=torch.uint8, device=dev)
    dres = torch.empty(len(got), dtype=torch.int32, device=dev)
    eng.aas_predict_client_movement_device(dp, out, dres)
    torch.cuda.synchronize()
    _same_moves(out.cpu().numpy().view(AAS_MOVE_DTYPE).reshape(-1), got, dres.cpu().numpy(), res, "device pointers")
