"""The product's host loader (cmgpu_flatmap_from_bsp, cm_load.cpp) against the reference's loaded clipMap_t.

Host code only: runs without a GPU."""
import numpy as np
import pytest

from helpers import cmoracle, cmref, get_map, have_ref
from threecore_b200 import bspgen
from threecore_b200.engine import CMGPUError, FlatMap

KINDS = ["tiny", "room", "arena_small", "arena_models", "patches_small"]


@pytest.mark.skipif(not have_ref(), reason="needs oracle/_ref/libcmref.so")
@pytest.mark.parametrize("kind", KINDS)
def test_flatmap_equals_reference_clipmap(kind):
    m = get_map(kind)
    want = cmoracle.flat_from_ref_dump(cmref.RefCM().load(m.data).dump())
    got = FlatMap.from_bsp(m.data).to_dict()
    for key, w in want.items():
        g = got[key]
        if isinstance(w, np.ndarray):
            assert g.shape == w.shape, f"{kind}.{key}: shape {g.shape} != {w.shape}"
            if key in ("brush_bounds", "planes"):
                # the temp-box brush/planes hold whatever the last CM_TempBoxModel left; ours are zero
                bb = want["box_brush"] if key == "brush_bounds" else slice(want["box_first_plane"], None)
                g = g.copy(); w = w.copy()
                if key == "brush_bounds":
                    g[bb] = 0; w[bb] = 0
                else:
                    g[bb, 3] = 0; w[bb, 3] = 0
            assert g.tobytes() == w.tobytes(), f"{kind}.{key} differs"
        else:
            assert g == w, f"{kind}.{key}: {g} != {w}"


def test_loader_rejects_malformed_maps():
    good = get_map("tiny").data
    with pytest.raises(CMGPUError):
        FlatMap.from_bsp(good[:100])                      # truncated header
    bad = bytearray(good); bad[4:8] = (47).to_bytes(4, "little")
    with pytest.raises(CMGPUError):
        FlatMap.from_bsp(bytes(bad))                      # wrong version (cm_load.c:699)
    bad = bytearray(good); bad[8 + 8 * bspgen.LUMP_PLANES + 4: 8 + 8 * bspgen.LUMP_PLANES + 8] = (17).to_bytes(4, "little")
    with pytest.raises(CMGPUError):
        FlatMap.from_bsp(bytes(bad))                      # funny lump size
    bad = bytearray(good); bad[8 + 8 * bspgen.LUMP_NODES: 8 + 8 * bspgen.LUMP_NODES + 4] = (len(good) + 4096).to_bytes(4, "little")
    with pytest.raises(CMGPUError):
        FlatMap.from_bsp(bytes(bad))                      # lump outside the file (cm_load.c:703-709)


def test_flatmap_roundtrip_through_dict():
    fm = FlatMap.from_bsp(get_map("room").data)
    d = fm.to_dict()
    d2 = FlatMap.from_dict(d).to_dict()
    for k, v in d.items():
        if isinstance(v, np.ndarray):
            assert v.tobytes() == d2[k].tobytes(), k
        else:
            assert v == d2[k], k


def test_generator_is_deterministic():
    import hashlib
    a = bspgen.arena_map(n_brushes=400, seed=5, size_xy=2048, size_z=1024, n_models=4).data
    b = bspgen.arena_map(n_brushes=400, seed=5, size_xy=2048, size_z=1024, n_models=4).data
    assert hashlib.sha256(a).hexdigest() == hashlib.sha256(b).hexdigest()
    s1, e1 = bspgen.make_sweeps(get_map("tiny"), 1000, 3)
    s2, e2 = bspgen.make_sweeps(get_map("tiny"), 1000, 3)
    assert s1.tobytes() == s2.tobytes() and e1.tobytes() == e2.tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("kind", ["patches_small", "config3"])
def test_gpu_tessellation_equals_host_and_reference(kind):
    """CM_GeneratePatchCollide as one GPU thread per patch (cmgpu_flatmap_from_bsp_device) -- every
    array of the flattened map, patch planes and facets included, must be the bytes the host loader makes, which
    test_flatmap_equals_reference_clipmap pins to the reference's own loaded clipMap_t"""
    m = get_map("patches_small") if kind == "patches_small" else bspgen.patch_map(n_brushes=2000, n_patches=512, seed=0xC0FFEE03)
    host = FlatMap.from_bsp(m.data).to_dict()
    dev = FlatMap.from_bsp(m.data, tess_device=0).to_dict()
    assert host.keys() == dev.keys()
    for k in host:
        a, b = np.asarray(host[k]), np.asarray(dev[k])
        assert a.shape == b.shape and a.tobytes() == b.tobytes(), f"{kind}: `{k}` differs between host and GPU tessellation"
    assert len(host["patches"]) > 10 and len(host["facets"]) > 100
    if have_ref():
        ref = cmoracle.flat_from_ref_dump(cmref.RefCM().load(m.data).dump())
        for k in ("patch_planes", "patch_plane_signbits", "facets", "borders", "patch_bounds", "patches"):
            assert np.asarray(ref[k]).tobytes() == np.asarray(dev[k]).tobytes(), f"{kind}: `{k}` differs from the reference's clipMap_t"
