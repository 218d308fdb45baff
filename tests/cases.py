"""Deterministic parity cases shared by the CPU (oracle vs reference) and GPU (engine vs oracle) suites.

Each case is a dict: map kind + a query kind + the numpy inputs.  The same case objects feed
  * oracle/_ref  (the unmodified reference, when built),
  * oracle/libcmoracle (the C restatement),
  * the CUDA engine through the C ABI,
so a parity test is always "same inputs, byte-identical outputs".
"""
from __future__ import annotations

import numpy as np

from helpers import bspgen, get_map

MASKS = np.array([bspgen.MASK_SOLID, bspgen.MASK_PLAYERSOLID, bspgen.MASK_ALL, bspgen.CONTENTS_WATER,
                  bspgen.CONTENTS_PLAYERCLIP | bspgen.CONTENTS_MONSTERCLIP, 0], dtype=np.int32)


def _edge_sweeps(m, rng, n):
    """Hand-picked troublemakers: axis-aligned rays, rays along brush faces, zero-length, tiny, huge."""
    lo = m.world_mins.astype(np.float64) + 70
    hi = m.world_maxs.astype(np.float64) - 70
    s = rng.uniform(lo, hi, size=(n, 3))
    e = s.copy()
    k = n // 8
    # axis aligned
    for a in range(3):
        e[a * k:(a + 1) * k, a] += rng.uniform(-900, 900, size=k)
    # snapped to the integer lattice (brush faces are on it): grazing and coplanar cases
    s[3 * k:5 * k] = np.round(s[3 * k:5 * k])
    e[3 * k:4 * k] = s[3 * k:4 * k] + np.round(rng.uniform(-300, 300, size=(k, 3)))
    e[4 * k:5 * k] = s[4 * k:5 * k]
    e[4 * k:5 * k, rng.integers(0, 3)] += 64.0
    # starting on / in brushes
    nb = m.num_world_brushes
    which = rng.integers(0, nb, size=k)
    s[5 * k:6 * k] = m.brush_maxs[which]          # exactly a brush corner
    e[5 * k:6 * k] = s[5 * k:6 * k] + rng.uniform(-200, 200, size=(k, 3))
    # tiny moves and exact zero-length
    e[6 * k:7 * k] = s[6 * k:7 * k] + rng.uniform(-0.2, 0.2, size=(k, 3))
    e[7 * k:] = s[7 * k:]
    return np.ascontiguousarray(s, np.float32), np.ascontiguousarray(e, np.float32)


def world_cases(kind, n=20000, seed=1):
    """CM_BoxTrace through the world: shared hulls, per-item hulls/masks, edge rays."""
    m = get_map(kind)
    rng = np.random.default_rng(seed)
    out = []
    s, e = bspgen.make_sweeps(m, n, seed + 10, len_short=(1.0, 400.0), len_long=(400.0, 3000.0))
    out.append(dict(name=f"{kind}/point", map=kind, kind="box", starts=s, ends=e, mins=None, maxs=None,
                    mask=bspgen.MASK_SOLID))
    out.append(dict(name=f"{kind}/player", map=kind, kind="box", starts=s, ends=e, mins=bspgen.PLAYER_MINS,
                    maxs=bspgen.PLAYER_MAXS, mask=bspgen.MASK_PLAYERSOLID))
    # odd hull whose symmetrisation is inexact in float
    out.append(dict(name=f"{kind}/oddhull", map=kind, kind="box", starts=s, ends=e,
                    mins=np.array([-7.3, -0.1, -33.7], np.float32), maxs=np.array([12.9, 0.3, 1.1], np.float32),
                    mask=bspgen.MASK_ALL))
    # per-item hull and mask
    half = rng.uniform(0, 40, size=(n, 3))
    half[rng.random(n) < 0.3] = 0.0
    mins = (-half * rng.uniform(0.5, 1.5, size=(n, 3))).astype(np.float32)
    maxs = half.astype(np.float32)
    masks = MASKS[rng.integers(0, len(MASKS), size=n)]
    out.append(dict(name=f"{kind}/peritem", map=kind, kind="box", starts=s, ends=e, mins=mins, maxs=maxs, mask=masks))
    es, ee = _edge_sweeps(m, rng, max(n // 4, 64))
    out.append(dict(name=f"{kind}/edges-point", map=kind, kind="box", starts=es, ends=ee, mins=None, maxs=None,
                    mask=bspgen.MASK_ALL))
    out.append(dict(name=f"{kind}/edges-player", map=kind, kind="box", starts=es, ends=ee, mins=bspgen.PLAYER_MINS,
                    maxs=bspgen.PLAYER_MAXS, mask=bspgen.MASK_PLAYERSOLID))
    # position tests only, huge box (many leafs) and small box
    ps = s[: n // 4]
    out.append(dict(name=f"{kind}/position-big", map=kind, kind="box", starts=ps, ends=ps,
                    mins=np.array([-300, -300, -200], np.float32), maxs=np.array([300, 300, 200], np.float32),
                    mask=bspgen.MASK_ALL))
    out.append(dict(name=f"{kind}/position-point", map=kind, kind="box", starts=ps, ends=ps, mins=None, maxs=None,
                    mask=bspgen.MASK_ALL))
    return out


def entity_cases(kind, n=20000, seed=2):
    """CM_TransformedBoxTrace / CM_BoxTrace against inline models and temp boxes, point contents."""
    m = get_map(kind)
    out = []
    t = bspgen.make_entity_traces(m, n, seed)
    hull = (bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS)
    out.append(dict(name=f"{kind}/transformed", map=kind, kind="transformed", mins=hull[0], maxs=hull[1],
                    mask=bspgen.MASK_ALL, **t))
    t2 = dict(t)
    t2["angles"] = np.zeros_like(t["angles"])
    out.append(dict(name=f"{kind}/translated-point", map=kind, kind="transformed", mins=None, maxs=None,
                    mask=bspgen.MASK_PLAYERSOLID, **t2))
    # untransformed CM_BoxTrace straight against a model handle (sound code does this, snd_dma.c:448)
    local = dict(starts=(t["starts"] - t["origins"]).astype(np.float32), ends=(t["ends"] - t["origins"]).astype(np.float32))
    out.append(dict(name=f"{kind}/model-box", map=kind, kind="box", mins=hull[0], maxs=hull[1], mask=bspgen.MASK_ALL,
                    model=t["models"], boxmins=t["boxmins"], boxmaxs=t["boxmaxs"], **local))
    pts = bspgen.make_points(m, n, seed + 5)
    out.append(dict(name=f"{kind}/pointcontents", map=kind, kind="points", points=pts))
    out.append(dict(name=f"{kind}/pointcontents-entities", map=kind, kind="points",
                    points=(t["origins"] + np.random.default_rng(seed).uniform(-80, 80, size=(n, 3))).astype(np.float32),
                    models=t["models"], origins=t["origins"], angles=t["angles"], boxmins=t["boxmins"],
                    boxmaxs=t["boxmaxs"]))
    out.append(dict(name=f"{kind}/pointcontents-models-local", map=kind, kind="points",
                    points=np.random.default_rng(seed + 1).uniform(-150, 150, size=(n, 3)).astype(np.float32),
                    models=t["models"], boxmins=t["boxmins"], boxmaxs=t["boxmaxs"]))
    return out


def run_case(impl, case):
    """impl: RefCM / OracleCM / Engine -- they share method names and argument order."""
    k = case["kind"]
    if k == "box":
        kw = {}
        if "model" in case:
            kw = dict(model=case["model"], boxmins=case.get("boxmins"), boxmaxs=case.get("boxmaxs"))
        return impl.box_trace(case["starts"], case["ends"], case["mins"], case["maxs"], mask=case["mask"], **kw)
    if k == "transformed":
        return impl.transformed_box_trace(case["starts"], case["ends"], case["mins"], case["maxs"], case["models"],
                                          case["mask"], case["origins"], case["angles"], case["boxmins"], case["boxmaxs"])
    if k == "points":
        if "origins" in case:
            if hasattr(impl, "transformed_point_contents"):
                return impl.transformed_point_contents(case["points"], case["models"], case["origins"], case["angles"],
                                                       case["boxmins"], case["boxmaxs"])
            return impl.point_contents(case["points"], case["models"], case["origins"], case["angles"],
                                       case["boxmins"], case["boxmaxs"])
        if "models" in case:
            return impl.point_contents(case["points"], case["models"], boxmins=case["boxmins"], boxmaxs=case["boxmaxs"])
        return impl.point_contents(case["points"])
    raise KeyError(k)
