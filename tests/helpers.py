"""Shared helpers for the parity tests: map/ray fixtures and record comparison."""
from __future__ import annotations

import functools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import cmoracle, cmref  # noqa: E402  (test infrastructure)
from threecore_b200 import bspgen  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")


def have_ref() -> bool:
    return cmref.available()


@functools.lru_cache(maxsize=None)
def get_map(kind: str):
    """Small deterministic maps used across tests."""
    if kind == "tiny":
        return bspgen.tiny_map()
    if kind == "room":
        return bspgen.room_map(n_brushes=500, seed=0xC0FFEE01)
    if kind == "arena_small":
        return bspgen.arena_map(n_brushes=2500, seed=0xC0FFEE12, size_xy=4096.0, size_z=2048.0,
                                rotated_frac=0.15, nonaxial_node_prob=0.05)
    if kind == "arena_models":
        return bspgen.arena_map(n_brushes=1500, seed=0xC0FFEE14, size_xy=4096.0, size_z=2048.0,
                                rotated_frac=0.15, nonaxial_node_prob=0.05, n_models=48)
    if kind == "patches_small":
        return bspgen.patch_map(n_brushes=300, n_patches=64, seed=0xC0FFEE13, size_xy=3072.0, size_z=1536.0,
                                n_models=6)
    raise KeyError(kind)


def comb_map(slabs):
    """`slabs` thin brushes in a row, split off one at a time: a comb-shaped BSP about as deep as there are slabs"""
    mb = bspgen.MapBuilder()
    half = 50.0 * slabs + 192.0
    bspgen._room_walls(mb, [-half, -64, -64], [half, 64, 64])
    for i in range(slabs):
        x = -50.0 * slabs + 100.0 * i
        mb.box_brush([x, -32, -32], [x + 8, 32, 32], 0)
    return mb.finish(seed=1, leaf_max=1, comb=True, max_depth=4 * slabs + 64), half


def oracle_for(flat: dict, **cvars) -> "cmoracle.OracleCM":
    return cmoracle.OracleCM(flat, **cvars)


def flat_from_ref(data: bytes) -> dict:
    r = cmref.RefCM().load(data)
    return cmoracle.flat_from_ref_dump(r.dump())


def trace_bytes(a: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(a).view(np.uint8).reshape(len(a), 56)


def assert_traces_equal(got: np.ndarray, want: np.ndarray, what: str = ""):
    """Bit-exact comparison of 56-byte trace_t records with a readable first mismatch."""
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    g, w = trace_bytes(got), trace_bytes(want)
    bad = (g != w).any(axis=1)
    if bad.any():
        i = int(np.flatnonzero(bad)[0])
        fields = [f for f in got.dtype.names if not np.array_equal(got[i][f], want[i][f])]
        raise AssertionError(f"{what}: {int(bad.sum())} of {len(got)} records differ; first at {i}, fields {fields}\n"
                             f" got  {got[i]}\n want {want[i]}")


def flat_for(data: bytes, prefer_product: bool = False) -> dict:
    """Flat dict for the oracle: from the reference's own loaded clipMap_t when the reference library
    is available, otherwise from the product's host loader (which test_loader.py checks against it)."""
    if have_ref() and not prefer_product:
        return flat_from_ref(data)
    from threecore_b200.engine import FlatMap
    return FlatMap.from_bsp(data).to_dict()


def load_golden(name: str):
    """-> (bsp bytes, [(case dict, expected output)])"""
    import numpy as np
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    out = []
    for i in range(int(z["num_cases"])):
        c = {"name": str(z[f"c{i}_name"]), "kind": str(z[f"c{i}_kind"]), "mins": None, "maxs": None}
        pre = f"c{i}_in_"
        for k in z.files:
            if k.startswith(pre):
                v = z[k]
                c[k[len(pre):]] = v if v.ndim else v.item()
        want = z[f"c{i}_out"]
        if want.dtype == np.uint8:
            want = np.ascontiguousarray(want).view(cmref.TRACE_DTYPE).reshape(-1)
        out.append((c, want))
    return z["bsp"].tobytes(), out


GOLDEN_SETS = ["tiny_world", "room_world", "models_world", "models_entities", "patches_world", "patches_entities"]


def load_golden_sv():
    """tests/golden/sv_world.npz: moves and points against the scene of sv_scene.make_entities(arena_models, 300, 71),
    answered by the reference's SV_Trace / SV_PointContents (tests/golden/make_golden_sv.py)"""
    z = np.load(os.path.join(GOLDEN, "sv_world.npz"), allow_pickle=False)
    g = {k: np.ascontiguousarray(z[k]) for k in z.files}
    g["traces"] = g["traces"].view(cmref.TRACE_DTYPE).reshape(-1)
    g["bsp_len"] = int(g["bsp_len"].reshape(-1)[0])
    return g
