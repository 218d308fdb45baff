"""Synthetic AAS worlds for the botlib tracer (AAS_TraceClientBBox, code/botlib/be_aas_sample.c:399-665).

The reference ships neither .aas files nor BSPC, the tool that makes them.  What the tracer reads of an AAS file is
small -- planes (stored in opposite pairs: plane n ^ 1 is plane n flipped, aasfile.h / be_aas_sample.c:492), the BSP
nodes (node 0 is a dummy, the root is node 1; a child > 0 is a node, 0 a solid leaf, < 0 minus an area number,
be_aas_sample.c:225,432,506) and each area's presence type (aasfile.h:32-34) -- so a world of that shape is generated:
random boxes in a room, expanded by the bot's bounding box the way BSPC expands brushes (standing box for
PRESENCE_NORMAL, crouching box for PRESENCE_CROUCH), cut by a k-d tree along the expanded faces; a leaf inside a
crouch-expanded box is solid, one inside only a stand-expanded box is an area bots can only crouch in, the rest are
areas for both.  A share of the splits uses tilted planes and odd (flipped) plane numbers, which real files have too.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

PRESENCE_NONE, PRESENCE_NORMAL, PRESENCE_CROUCH = 1, 2, 4
STAND_MINS, STAND_MAXS = np.array([-15.0, -15.0, -24.0]), np.array([15.0, 15.0, 32.0])
CROUCH_MINS, CROUCH_MAXS = np.array([-15.0, -15.0, -24.0]), np.array([15.0, 15.0, 16.0])


@dataclass
class AasWorld:
    planes: np.ndarray          # [n,5] float32: normal, dist, type
    nodes: np.ndarray           # [n,3] int32: planenum, children[0], children[1]
    presence: np.ndarray        # [numareasettings] int32 (area 0 unused)
    contents: np.ndarray        # [numareasettings] int32 AREACONTENTS_* (aasfile.h:73-89)
    mins: np.ndarray
    maxs: np.ndarray
    stats: dict


def make_world(n_boxes=400, seed=0xAA5001, extent=4096.0, height=1024.0, leaf_depth=40, tilted_frac=0.08) -> AasWorld:
    rng = np.random.default_rng(seed)
    lo = np.array([-extent / 2, -extent / 2, 0.0])
    hi = np.array([extent / 2, extent / 2, height])
    size = rng.uniform([48, 48, 32], [384, 384, 256], size=(n_boxes, 3)).round()
    org = rng.uniform(lo + 64, hi - 64 - size, size=(n_boxes, 3)).round()
    bmin, bmax = org, org + size
    # a point of the bot's origin is "inside" when its box would touch the brush: brush expanded by -maxs / -mins
    smin, smax = bmin - STAND_MAXS, bmax - STAND_MINS
    cmin, cmax = bmin - CROUCH_MAXS, bmax - CROUCH_MINS
    planes, nodes, presence = [], [[0, 0, 0]], [0]
    plane_index = {}

    def plane(normal, dist, flip):
        """index of (normal, dist) -- or of its opposite when flip -- in a table of opposite pairs"""
        key = (tuple(np.float32(normal).tolist()), float(np.float32(dist)))
        if key not in plane_index:
            n32 = np.float32(normal)
            axis = int(np.argmax(np.abs(n32)))
            typ = axis if abs(float(n32[axis])) == 1.0 else 3 + axis
            plane_index[key] = len(planes)
            planes.append([*n32.tolist(), float(np.float32(dist)), typ])
            planes.append([*(-n32).tolist(), float(-np.float32(dist)), typ])
        return plane_index[key] ^ (1 if flip else 0)

    stats = dict(max_depth=0, solid=0, tilted=0)

    def leaf(cmin_, cmax_, ids):
        c = 0.5 * (cmin_ + cmax_)
        in_crouch = bool(np.any(np.all((c > cmin[ids]) & (c < cmax[ids]), axis=1))) if ids.size else False
        in_stand = bool(np.any(np.all((c > smin[ids]) & (c < smax[ids]), axis=1))) if ids.size else False
        if in_crouch:
            stats["solid"] += 1
            return 0
        presence.append(PRESENCE_CROUCH if in_stand else PRESENCE_NORMAL | PRESENCE_CROUCH)
        return -(len(presence) - 1)

    def build(cmin_, cmax_, ids, depth):
        stats["max_depth"] = max(stats["max_depth"], depth)
        if depth >= leaf_depth or ids.size == 0:
            return leaf(cmin_, cmax_, ids)
        # candidate faces of the expanded boxes strictly inside the cell, on its longest axis first
        for a in np.argsort(-(cmax_ - cmin_)):
            cand = np.concatenate([smin[ids, a], smax[ids, a], cmax[ids, a]])
            cand = np.unique(cand[(cand > cmin_[a] + 1e-3) & (cand < cmax_[a] - 1e-3)])
            if cand.size:
                break
        else:
            return leaf(cmin_, cmax_, ids)
        d = float(cand[cand.size // 2])
        me = len(nodes)
        nodes.append([0, 0, 0])
        normal = np.zeros(3)
        normal[a] = 1.0
        tilted = rng.random() < tilted_frac and (cmax_ - cmin_).min() > 64
        if tilted:
            # a slightly tilted plane through the same point: both sides keep every box that straddles the axial cut region
            stats["tilted"] += 1
            t = rng.normal(0, 0.12, 3)
            t[a] = 1.0
            normal = t / np.linalg.norm(t)
            pt = 0.5 * (cmin_ + cmax_)
            pt[a] = d
            dist = float(np.float32(normal) @ np.float32(pt))
        else:
            dist = d
        flip = rng.random() < 0.3
        pn = plane(normal, dist, flip)
        nodes[me][0] = pn
        fmin, bmax_ = cmin_.copy(), cmax_.copy()
        if tilted:
            # children keep the whole cell as their bounds (conservative); boxes are not filtered
            front_ids = back_ids = ids
            fmin_c, fmax_c, bmin_c, bmax_c = cmin_, cmax_, cmin_, cmax_
            depth_next = max(depth + 1, leaf_depth - 6)                # tilted cells end soon: their leaves are classified by centre
        else:
            fmin[a] = d
            bmax_[a] = d
            front_ids = ids[smax[ids, a] > d]
            back_ids = ids[smin[ids, a] < d]
            fmin_c, fmax_c, bmin_c, bmax_c = fmin, cmax_, cmin_, bmax_
            depth_next = depth + 1
        front = build(fmin_c, fmax_c, front_ids, depth_next)
        back = build(bmin_c, bmax_c, back_ids, depth_next)
        # children[0] is the side the plane's normal points to (be_aas_sample.c:247,590-598)
        nodes[me][1], nodes[me][2] = (back, front) if flip else (front, back)
        return me

    import sys
    old = sys.getrecursionlimit()
    sys.setrecursionlimit(max(old, 10000))
    try:
        root = build(lo.copy(), hi.copy(), np.arange(n_boxes), 0)
    finally:
        sys.setrecursionlimit(old)
    assert root == 1, "the root must be node 1"
    stats.update(nodes=len(nodes), planes=len(planes), areas=len(presence) - 1)
    # a few areas are liquids, jump pads, teleporters and cluster portals: what the movement predictor's stop events look at
    crng = np.random.default_rng(seed ^ 0xC0)
    contents = crng.choice([0, 1, 2, 4, 8, 64, 128], size=len(presence), p=[0.82, 0.05, 0.02, 0.02, 0.03, 0.03, 0.03]).astype(np.int32)
    contents[0] = 0
    return AasWorld(np.asarray(planes, np.float32).reshape(-1, 5), np.asarray(nodes, np.int32).reshape(-1, 3),
                    np.asarray(presence, np.int32), contents, lo.astype(np.float32), hi.astype(np.float32), stats)


def make_lines(w: AasWorld, n, seed, length=(8.0, 1200.0), zero_frac=0.02):
    rng = np.random.default_rng(seed)
    s = rng.uniform(w.mins + 8, w.maxs - 8, size=(n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    e = s + d * rng.uniform(length[0], length[1], size=(n, 1))
    axis = rng.random(n) < 0.15                                   # axis-parallel moves: ties with axial planes
    e[axis, 1:] = s[axis, 1:]
    snap = rng.random(n) < 0.2                                    # lattice-snapped starts land exactly on planes
    s[snap] = np.round(s[snap])
    z = rng.random(n) < zero_frac
    e[z] = s[z]
    return s.astype(np.float32), e.astype(np.float32)


SE_ALL = [1, 2, 4, 8, 16, 32, 64, 128, 256, 512, 1024, 2048, 4096]


def make_moves(w: AasWorld, n, seed, maxframes=(1, 40)):
    """inputs of n calls of AAS_PredictClientMovement (be_aas_move.c:965): where the bot is, how it moves, what the command is,
    and which events stop the prediction -- the combinations be_ai_move.c and be_aas_reach.c use, and random ones"""
    rng = np.random.default_rng(seed)
    org = rng.uniform(w.mins + 32, w.maxs - 32, size=(n, 3))
    low = rng.random(n) < 0.6                                          # most bots are near the floor
    org[low, 2] = rng.uniform(w.mins[2] + 24, w.mins[2] + 96, size=int(low.sum()))
    yaw = rng.uniform(0, 2 * np.pi, n)
    speed = rng.choice([0.0, 100.0, 320.0, 400.0, 600.0], size=n)
    vel = np.stack([np.cos(yaw) * speed, np.sin(yaw) * speed, rng.choice([0.0, -60.0, 270.0, -400.0], size=n)], 1)
    cyaw = yaw + rng.normal(0, 0.5, n)
    cmd = np.stack([np.cos(cyaw) * 400.0, np.sin(cyaw) * 400.0, rng.choice([0.0, 0.0, 400.0, -400.0], size=n)], 1)
    cmd[rng.random(n) < 0.1] = 0.0
    typical = [1 | 32 | 4 | 8 | 16 | 64, 4 | 8 | 16 | 32 | 64, 128 | 1024, 512 | 1 | 32, 1 | 2, 2048 | 1, 4096 | 256 | 1, 0]
    stop = np.where(rng.random(n) < 0.6, rng.choice(typical, size=n),
                    [int(sum(rng.choice(SE_ALL, size=rng.integers(1, 5), replace=False))) for _ in range(n)])
    # make the rarer stop events reachable: some bots head for the area a few frames ahead of them (SE_ENTERAREA,
    # SE_HITGROUNDAREA), some pass the 8-unit box around the world origin that SE_HITBOUNDINGBOX tests (be_aas_move.c:973)
    stoparea = rng.integers(0, len(w.presence), n)
    ahead = rng.random(n) < 0.25
    stoparea[ahead] = point_area_num(w, (org + vel * 0.3)[ahead])
    box = rng.random(n) < 0.05
    org[box] = rng.uniform(-40, 40, size=(int(box.sum()), 3)) + [0, 0, 0]
    vel[box] = -org[box] * rng.uniform(1.0, 6.0, size=(int(box.sum()), 1))
    stop = np.asarray(stop, np.int64)
    stop[box] |= 2048
    return dict(origins=org.astype(np.float32), presencetypes=rng.choice([PRESENCE_NORMAL, PRESENCE_NORMAL, PRESENCE_CROUCH], size=n).astype(np.int32),
                ongrounds=rng.integers(0, 2, n).astype(np.int32), velocities=vel.astype(np.float32), cmdmoves=cmd.astype(np.float32),
                cmdframes=rng.integers(0, 12, n).astype(np.int32), maxframes=rng.integers(maxframes[0], maxframes[1], n).astype(np.int32),
                frametimes=rng.choice([0.1, 0.1, 0.05, 0.0], size=n).astype(np.float32), stopevents=np.asarray(stop, np.int32),
                stopareanums=stoparea.astype(np.int32))


def point_area_num(w: AasWorld, pts) -> np.ndarray:
    """AAS_PointAreaNum for generator purposes (float64 here: only used to pick interesting targets)"""
    pts = np.asarray(pts, np.float64)
    node = np.ones(len(pts), np.int64)
    live = node > 0
    while live.any():
        nd = w.nodes[node[live]]
        pl = w.planes[nd[:, 0]].astype(np.float64)
        d = (pts[live] * pl[:, :3]).sum(1) - pl[:, 3]
        node[live] = np.where(d > 0, nd[:, 1], nd[:, 2])
        live = node > 0
    return np.where(node < 0, -node, 0)
