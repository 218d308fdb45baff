"""Inputs for the leaf / cluster / visibility queries: maps with clusters, areas and a visibility
lump, query boxes from a few units to most of the map, entities and viewpoints.  Shared by the CPU tests (oracle against
the reference), the golden generator and the GPU parity tests."""
from __future__ import annotations

import functools
import struct

import numpy as np

from helpers import bspgen, get_map


@functools.lru_cache(maxsize=None)
def vis_map(kind: str):
    """kind: '<base>' of helpers.get_map + '_vis' (vis lump) or '_novis' (clusters and areas, no lump)"""
    base, _, tail = kind.rpartition("_")
    m = get_map(base)
    return bspgen.with_visibility(m, seed=0x715 + len(base), n_areas=5, vised=(tail == "vis"))


def vis_lump(data: bytes):
    """-> (numClusters, clusterBytes, matrix uint8 [numClusters*clusterBytes]) or (0, 0, None) without a lump"""
    ofs, ln = struct.unpack_from("<ii", data, 8 + 8 * bspgen.LUMP_VISIBILITY)
    if not ln:
        return 0, 0, None
    nc, cb = struct.unpack_from("<ii", data, ofs)
    return nc, cb, np.frombuffer(data, np.uint8, nc * cb, ofs + 8).copy()


def query_boxes(m, n, seed):
    """n boxes inside (and a little outside) the world: 40 % player-sized, 30 % room-sized, 20 % a good part of the
    map (overflows short lists and the 128-leaf list of SV_LinkEntity), 10 % degenerate (a point / a plane)."""
    rng = np.random.default_rng(seed)
    lo, hi = m.world_mins.astype(np.float64), m.world_maxs.astype(np.float64)
    ext = float((hi - lo).max())
    c = lo + rng.random((n, 3)) * (hi - lo)
    kind = rng.random(n)
    half = np.where(kind[:, None] < 0.4, rng.uniform(4, 48, (n, 3)),
                    np.where(kind[:, None] < 0.7, rng.uniform(64, 400, (n, 3)),
                             np.where(kind[:, None] < 0.9, rng.uniform(min(400.0, 0.25 * ext), 0.5 * ext, (n, 3)), 0.0)))
    flat = (kind >= 0.95)
    half[flat, 2] = 0.0
    half[flat, 0:2] = rng.uniform(16, 256, (int(flat.sum()), 2))
    snap = rng.random(n) < 0.3                 # integer coordinates land on node planes exactly
    mins, maxs = c - half, c + half
    mins[snap], maxs[snap] = np.round(mins[snap]), np.round(maxs[snap])
    return mins.astype(np.float32), maxs.astype(np.float32)


def entities(m, n, seed):
    """(mins, maxs, origin) of n box entities; SV_LinkEntity makes absmin/absmax = origin + mins/maxs -/+ 1"""
    rng = np.random.default_rng(seed)
    lo, hi = m.world_mins.astype(np.float64), m.world_maxs.astype(np.float64)
    origin = np.round(lo + rng.random((n, 3)) * (hi - lo))
    big = rng.random(n)
    half = np.where(big[:, None] < 0.6, rng.uniform(8, 40, (n, 3)),
                    np.where(big[:, None] < 0.85, rng.uniform(100, 500, (n, 3)), rng.uniform(600, 1800, (n, 3))))
    half = np.round(half)
    return (-half).astype(np.float32), half.astype(np.float32), origin.astype(np.float32)


def viewpoints(m, n, seed):
    rng = np.random.default_rng(seed)
    lo, hi = m.world_mins.astype(np.float64), m.world_maxs.astype(np.float64)
    return (lo + rng.random((n, 3)) * (hi - lo)).astype(np.float32)


def portal_script(n_areas, seed, steps=12):
    """a deterministic sequence of (area1, area2, open) that never closes a portal that is not open"""
    rng = np.random.default_rng(seed)
    opened = []
    script = []
    for _ in range(steps):
        if opened and rng.random() < 0.3:
            a, b = opened.pop(int(rng.integers(len(opened))))
            script.append((a, b, False))
        else:
            a, b = (int(x) for x in rng.integers(0, n_areas, 2))
            script.append((a, b, True))
            opened.append((a, b))
    return script
