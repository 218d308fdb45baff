"""Synthetic IBSP v46 map and trace-stream generator.

The reference ships no map assets, so every workload in
BASELINE.json is a procedurally generated clip map.  This module emits byte
images that satisfy exactly what the reference loader consumes
(code/qcommon/cm_load.c:606-745, on-disk records code/qcommon/qfiles.h:173-299):

* 144-byte header, version 46, 17 lumps, each lump a whole number of records;
* brushes whose first six sides are the axial planes -x,+x,-y,+y,-z,+z
  (cm_load.c:230-239 derives the brush bounds from them), further sides are
  arbitrary outward-facing planes;
* a BSP of +axis (and optionally non-axial) node planes built here by
  recursive splitting, children >= 0 nodes, < 0 leaf -(leaf+1)
  (cm_trace.c:454-455); leafs list every brush / patch whose bounds touch them;
* inline models (dmodel_t) whose brushes/surfaces sit outside the world tree;
* MST_PATCH surfaces with odd control-grid sizes (cm_load.c:525-580).

Everything is numpy; all randomness comes from numpy's PCG64 with explicit seeds
so the same (kind, seed) gives the same bytes on every machine.
"""
from __future__ import annotations

import math
import struct
from dataclasses import dataclass, field

import numpy as np

# content / surface flags (code/qcommon/surfaceflags.h:30-82, code/game/bg_public.h:53-59)
CONTENTS_SOLID = 1
CONTENTS_LAVA = 8
CONTENTS_SLIME = 16
CONTENTS_WATER = 32
CONTENTS_PLAYERCLIP = 0x10000
CONTENTS_MONSTERCLIP = 0x20000
CONTENTS_BODY = 0x2000000
CONTENTS_TRIGGER = 0x40000000
MASK_SOLID = CONTENTS_SOLID
MASK_PLAYERSOLID = CONTENTS_SOLID | CONTENTS_PLAYERCLIP | CONTENTS_BODY
MASK_ALL = -1
SURF_NODAMAGE = 0x1
SURF_SLICK = 0x2
SURF_METALSTEPS = 0x1000
SURF_NOSTEPS = 0x2000
BOX_MODEL_HANDLE = 1023
MST_PLANAR = 1
MST_PATCH = 2

PLAYER_MINS = np.array([-15.0, -15.0, -24.0], dtype=np.float32)
PLAYER_MAXS = np.array([15.0, 15.0, 32.0], dtype=np.float32)

(LUMP_ENTITIES, LUMP_SHADERS, LUMP_PLANES, LUMP_NODES, LUMP_LEAFS, LUMP_LEAFSURFACES,
 LUMP_LEAFBRUSHES, LUMP_MODELS, LUMP_BRUSHES, LUMP_BRUSHSIDES, LUMP_DRAWVERTS,
 LUMP_DRAWINDEXES, LUMP_FOGS, LUMP_SURFACES, LUMP_LIGHTMAPS, LUMP_LIGHTGRID,
 LUMP_VISIBILITY) = range(17)

DEFAULT_SHADERS = [
    # name, surfaceFlags, contentFlags
    ("textures/synth/solid", 0, CONTENTS_SOLID),
    ("textures/synth/metal", SURF_METALSTEPS, CONTENTS_SOLID),
    ("textures/synth/slick", SURF_SLICK | SURF_NODAMAGE, CONTENTS_SOLID),
    ("textures/synth/clip", SURF_NOSTEPS, CONTENTS_PLAYERCLIP),
    ("textures/synth/water", 0, CONTENTS_WATER),
    ("textures/synth/monsterclip", 0, CONTENTS_MONSTERCLIP),
    ("textures/synth/curve", SURF_NODAMAGE, CONTENTS_SOLID),
    ("textures/synth/mover", SURF_METALSTEPS, CONTENTS_SOLID),
]


# ----------------------------------------------------------------------------
# map under construction
# ----------------------------------------------------------------------------
@dataclass
class _Brush:
    sides: list          # [(plane index, shader)]
    shader: int
    mins: np.ndarray
    maxs: np.ndarray
    rotated: bool = False


@dataclass
class _Patch:
    width: int
    height: int
    points: np.ndarray   # [height*width, 3] float32, row-major (cm_load.c:567-572)
    shader: int
    mins: np.ndarray = None
    maxs: np.ndarray = None


@dataclass
class SynthMap:
    """A generated map: the byte image plus what the ray generators need."""
    data: bytes
    world_mins: np.ndarray
    world_maxs: np.ndarray
    brush_mins: np.ndarray       # world brushes only
    brush_maxs: np.ndarray
    num_world_brushes: int
    num_brushes: int
    num_models: int
    model_mins: np.ndarray       # [num_models,3] raw dmodel bounds
    model_maxs: np.ndarray
    patch_mins: np.ndarray
    patch_maxs: np.ndarray
    stats: dict = field(default_factory=dict)


class MapBuilder:
    def __init__(self, shaders=None):
        self.shaders = list(shaders or DEFAULT_SHADERS)
        self.planes: list[tuple[float, float, float, float]] = []
        self._plane_lut: dict = {}
        self.brushes: list[_Brush] = []
        self.patches: list[_Patch] = []
        self.planar_surfaces = 0       # non-patch surfaces interleaved (ignored by CM)
        self.models: list[dict] = []   # submodels (index 1..)
        self.world_mins = None
        self.world_maxs = None

    # -- planes -------------------------------------------------------------
    def plane(self, normal, dist) -> int:
        n = np.asarray(normal, dtype=np.float32)
        key = (float(n[0]), float(n[1]), float(n[2]), float(np.float32(dist)))
        idx = self._plane_lut.get(key)
        if idx is None:
            idx = len(self.planes)
            self.planes.append(key)
            self._plane_lut[key] = idx
        return idx

    def _axial_sides(self, mins, maxs, shader):
        mins = np.asarray(mins, dtype=np.float32)
        maxs = np.asarray(maxs, dtype=np.float32)
        sides = []
        for a in range(3):
            nneg = [0.0, 0.0, 0.0]
            nneg[a] = -1.0
            npos = [0.0, 0.0, 0.0]
            npos[a] = 1.0
            sides.append((self.plane(nneg, -mins[a]), shader))
            sides.append((self.plane(npos, maxs[a]), shader))
        return sides

    # -- brushes ------------------------------------------------------------
    def box_brush(self, mins, maxs, shader=0, side_shaders=None) -> int:
        mins = np.asarray(mins, dtype=np.float32)
        maxs = np.asarray(maxs, dtype=np.float32)
        sides = self._axial_sides(mins, maxs, shader)
        if side_shaders is not None:
            sides = [(p, side_shaders[i]) for i, (p, _) in enumerate(sides)]
        self.brushes.append(_Brush(sides, shader, mins.copy(), maxs.copy()))
        return len(self.brushes) - 1

    def hull_brush(self, normals, dists, shader=0, snap=True, verts=None) -> int:
        """Convex brush from outward planes; the six axial bounding sides come first."""
        normals = np.asarray(normals, dtype=np.float64)
        dists = np.asarray(dists, dtype=np.float64)
        if verts is None:
            verts = _hull_vertices(normals, dists)
        if verts.shape[0] < 4:
            raise ValueError("degenerate hull brush")
        mins = verts.min(axis=0)
        maxs = verts.max(axis=0)
        if snap:
            mins = np.floor(mins)
            maxs = np.ceil(maxs)
        mins = mins.astype(np.float32)
        maxs = maxs.astype(np.float32)
        sides = self._axial_sides(mins, maxs, shader)
        for n, d in zip(normals, dists):
            nf = n.astype(np.float32)
            nz = np.flatnonzero(nf)
            if nz.size == 1 and abs(float(nf[nz[0]])) == 1.0:
                continue   # already covered by an axial side
            sides.append((self.plane(nf, np.float32(d)), shader))
        self.brushes.append(_Brush(sides, shader, mins, maxs, rotated=True))
        return len(self.brushes) - 1

    def rotated_box_brush(self, center, half, yaw_deg, pitch_deg=0.0, roll_deg=0.0, shader=0) -> int:
        m = _rotation(yaw_deg, pitch_deg, roll_deg)
        center = np.asarray(center, dtype=np.float64)
        normals, dists = [], []
        for a in range(3):
            for s in (1.0, -1.0):
                n = s * m[:, a]
                normals.append(n)
                dists.append(float(n @ center) + half[a])
        signs = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], dtype=np.float64)
        verts = center[None, :] + (signs * np.asarray(half, dtype=np.float64)[None, :]) @ m.T
        return self.hull_brush(normals, dists, shader, verts=verts)

    def wedge_brush(self, mins, maxs, axis=0, flip=False, shader=0) -> int:
        """Ramp: the box cut by a slanted plane rising along `axis`."""
        mins = np.asarray(mins, dtype=np.float64)
        maxs = np.asarray(maxs, dtype=np.float64)
        normals, dists = [], []
        for a in range(3):
            e = np.zeros(3)
            e[a] = 1.0
            normals.append(e.copy())
            dists.append(maxs[a])
            normals.append(-e)
            dists.append(-mins[a])
        run = maxs[axis] - mins[axis]
        rise = maxs[2] - mins[2]
        n = np.zeros(3)
        n[axis] = rise if flip else -rise
        n[2] = run
        n /= np.linalg.norm(n)
        p = mins.copy()
        if flip:
            p[axis] = maxs[axis]
        normals.append(n)
        dists.append(float(n @ p))
        return self.hull_brush(normals, dists, shader)

    # -- patches ------------------------------------------------------------
    def patch(self, width, height, points, shader=6) -> int:
        pts = np.asarray(points, dtype=np.float32).reshape(height * width, 3)
        p = _Patch(width, height, pts, shader, pts.min(axis=0), pts.max(axis=0))
        self.patches.append(p)
        return len(self.patches) - 1

    # -- submodels ----------------------------------------------------------
    def begin_model(self):
        self._model_first_brush = len(self.brushes)
        self._model_first_patch = len(self.patches)

    def end_model(self):
        self.models.append(dict(first_brush=self._model_first_brush,
                                num_brushes=len(self.brushes) - self._model_first_brush,
                                first_patch=self._model_first_patch,
                                num_patches=len(self.patches) - self._model_first_patch))

    # -- finalisation -------------------------------------------------------
    def finish(self, seed=0, leaf_max=4, nonaxial_node_prob=0.0, max_depth=48,
               planar_every=0, comb=False) -> SynthMap:
        """Build the BSP over the world brushes/patches and serialise all lumps.  comb=True peels one item per split
        (the lowest usable plane on the longest axis): a degenerate tree about as deep as the map has brushes."""
        rng = np.random.default_rng(seed ^ 0x5EED)
        n_world_brushes = self.models[0]["first_brush"] if self.models else len(self.brushes)
        n_world_patches = self.models[0]["first_patch"] if self.models else len(self.patches)

        # surface numbering: optionally interleave ignored planar surfaces so that
        # cm.surfaces[] has NULL holes (cm_trace.c:409-411)
        surf_of_patch = []
        surfaces = []          # ("patch", idx) or ("planar", None)
        for i in range(len(self.patches)):
            if planar_every and i % planar_every == 0:
                surfaces.append(("planar", None))
            surf_of_patch.append(len(surfaces))
            surfaces.append(("patch", i))

        bmins = np.array([b.mins for b in self.brushes], dtype=np.float64).reshape(-1, 3)
        bmaxs = np.array([b.maxs for b in self.brushes], dtype=np.float64).reshape(-1, 3)
        pmins = np.array([p.mins for p in self.patches], dtype=np.float64).reshape(-1, 3)
        pmaxs = np.array([p.maxs for p in self.patches], dtype=np.float64).reshape(-1, 3)

        imins = np.concatenate([bmins[:n_world_brushes], pmins[:n_world_patches]], axis=0)
        imaxs = np.concatenate([bmaxs[:n_world_brushes], pmaxs[:n_world_patches]], axis=0)
        # patches are thin: pad them so they are listed in every leaf they may touch
        imins[n_world_brushes:] -= 1.0
        imaxs[n_world_brushes:] += 1.0
        if imins.shape[0] == 0:
            raise ValueError("empty world")
        wmins = imins.min(axis=0) if self.world_mins is None else np.asarray(self.world_mins, dtype=np.float64)
        wmaxs = imaxs.max(axis=0) if self.world_maxs is None else np.asarray(self.world_maxs, dtype=np.float64)

        # candidate non-axial splitters: extra sides of rotated world brushes
        rotated_ids = [i for i in range(n_world_brushes) if self.brushes[i].rotated
                       and len(self.brushes[i].sides) > 6]

        nodes = []      # [plane, child0, child1]
        leafs = []      # [first_brush, n_brush, first_surf, n_surf]
        leafbrushes = []
        leafsurfaces = []
        stats = dict(max_depth=0, nonaxial_nodes=0)

        def make_leaf(ids):
            ids = np.sort(ids)
            br = ids[ids < n_world_brushes]
            pa = ids[ids >= n_world_brushes] - n_world_brushes
            # leaf order of brushes is semantics (first hit wins ties): shuffle deterministically
            if br.size > 1:
                br = br[rng.permutation(br.size)]
            leafs.append([len(leafbrushes), int(br.size), len(leafsurfaces), 0])
            leafbrushes.extend(int(x) for x in br)
            ns = 0
            for p in pa:
                s = surf_of_patch[int(p)]
                if planar_every and int(p) % planar_every == 0:
                    leafsurfaces.append(s - 1)   # the ignored planar neighbour
                    ns += 1
                leafsurfaces.append(s)
                ns += 1
            leafs[-1][3] = ns
            return -(len(leafs) - 1) - 1

        def choose_axial(ids, cmin, cmax):
            n = ids.size
            best = None
            order = np.argsort(-(cmax - cmin))
            mn = imins[ids]
            mx = imaxs[ids]
            for a in order:
                cand = np.concatenate([mn[:, a], mx[:, a]])
                cand = cand[(cand > cmin[a]) & (cand < cmax[a])]
                if cand.size == 0:
                    continue
                cand = np.unique(cand)
                if comb:
                    for c in cand:
                        nf = int(np.count_nonzero(mx[:, a] > c))
                        nb = int(np.count_nonzero(mn[:, a] < c))
                        if max(nf, nb) < n:
                            return (0.0, int(a), float(c))
                    continue
                centre = np.median(0.5 * (mn[:, a] + mx[:, a]))
                # a handful of candidates around the median
                pos = np.searchsorted(cand, centre)
                lo = max(0, pos - 3)
                sel = cand[lo:lo + 6]
                for c in sel:
                    nf = int(np.count_nonzero(mx[:, a] > c))
                    nb = int(np.count_nonzero(mn[:, a] < c))
                    if max(nf, nb) >= n:
                        continue
                    cost = max(nf, nb) + 0.25 * (nf + nb - n)
                    if best is None or cost < best[0]:
                        best = (cost, int(a), float(c))
                if best is not None and best[0] <= 0.75 * n:
                    break
            return best

        def classify_plane(ids, normal, dist):
            mn = imins[ids]
            mx = imaxs[ids]
            pos = np.where(normal >= 0, mx, mn) @ normal - dist   # farthest corner along +n
            neg = np.where(normal >= 0, mn, mx) @ normal - dist   # farthest along -n
            return pos > 0.0, neg < 0.0                           # touches front / back

        def build(ids, cmin, cmax, depth):
            stats["max_depth"] = max(stats["max_depth"], depth)
            n = ids.size
            if n <= leaf_max or depth >= max_depth:
                return make_leaf(ids)
            # occasionally split on a rotated brush face to exercise non-axial nodes
            if nonaxial_node_prob > 0 and rng.random() < nonaxial_node_prob:
                cands = [i for i in ids[ids < n_world_brushes] if self.brushes[int(i)].rotated
                         and len(self.brushes[int(i)].sides) > 6]
                if cands:
                    b = self.brushes[int(cands[int(rng.integers(len(cands)))])]
                    pl = b.sides[6 + int(rng.integers(len(b.sides) - 6))][0]
                    nx, ny, nz, d = self.planes[pl]
                    normal = np.array([nx, ny, nz], dtype=np.float64)
                    f, bk = classify_plane(ids, normal, float(d))
                    # items exactly on the plane belong to both sides
                    both = ~f & ~bk
                    f = f | both
                    bk = bk | both
                    if 0 < np.count_nonzero(f) < n and 0 < np.count_nonzero(bk) < n:
                        me = len(nodes)
                        nodes.append([pl, 0, 0])
                        stats["nonaxial_nodes"] += 1
                        c0 = build(ids[f], cmin, cmax, depth + 1)
                        c1 = build(ids[bk], cmin, cmax, depth + 1)
                        nodes[me][1] = c0
                        nodes[me][2] = c1
                        return me
            best = choose_axial(ids, cmin, cmax)
            if best is None:
                return make_leaf(ids)
            _, a, c = best
            normal = [0.0, 0.0, 0.0]
            normal[a] = 1.0
            pl = self.plane(normal, c)
            me = len(nodes)
            nodes.append([pl, 0, 0])
            front = imaxs[ids, a] > c
            back = imins[ids, a] < c
            fmin = cmin.copy()
            fmin[a] = c
            bmax = cmax.copy()
            bmax[a] = c
            c0 = build(ids[front], fmin, cmax, depth + 1)
            c1 = build(ids[back], cmin, bmax, depth + 1)
            nodes[me][1] = c0
            nodes[me][2] = c1
            return me

        import sys
        old = sys.getrecursionlimit()
        sys.setrecursionlimit(max(old, 10000))
        try:
            root = build(np.arange(imins.shape[0]), wmins.copy(), wmaxs.copy(), 0)
        finally:
            sys.setrecursionlimit(old)
        if root < 0:
            # the loader needs >= 1 node (cm_load.c:205): wrap the single leaf
            pl = self.plane([1.0, 0.0, 0.0], float(wmins[0]) - 64.0)
            leafs.append([len(leafbrushes), 0, len(leafsurfaces), 0])
            nodes.append([pl, root, -(len(leafs) - 1) - 1])
        stats.update(nodes=len(nodes), leafs=len(leafs), leafbrushes=len(leafbrushes),
                     leafsurfaces=len(leafsurfaces), planes=len(self.planes),
                     brushes=len(self.brushes), patches=len(self.patches))

        # ---- serialise ----------------------------------------------------
        lumps = [b""] * 17
        lumps[LUMP_ENTITIES] = b'{\n"classname" "worldspawn"\n}\n\x00'
        sh = bytearray()
        for name, sflags, cflags in self.shaders:
            nm = name.encode()[:63]
            sh += nm + b"\x00" * (64 - len(nm)) + struct.pack("<ii", sflags, cflags)
        lumps[LUMP_SHADERS] = bytes(sh)
        lumps[LUMP_PLANES] = np.array(self.planes, dtype="<f4").reshape(-1, 4).tobytes()

        nd = np.zeros((len(nodes), 9), dtype="<i4")
        nd[:, 0:3] = np.array(nodes, dtype=np.int64).reshape(-1, 3)
        lumps[LUMP_NODES] = nd.tobytes()

        lf = np.zeros((len(leafs), 12), dtype="<i4")
        la = np.array(leafs, dtype=np.int64).reshape(-1, 4)
        lf[:, 0] = np.arange(len(leafs))         # cluster
        lf[:, 1] = 0                             # area (numAreas^2 ints are allocated, cm_load.c:319)
        lf[:, 8] = la[:, 2]
        lf[:, 9] = la[:, 3]
        lf[:, 10] = la[:, 0]
        lf[:, 11] = la[:, 1]
        lumps[LUMP_LEAFS] = lf.tobytes()
        lumps[LUMP_LEAFSURFACES] = np.array(leafsurfaces, dtype="<i4").tobytes()
        lumps[LUMP_LEAFBRUSHES] = np.array(leafbrushes, dtype="<i4").tobytes()

        sides = []
        br = np.zeros((len(self.brushes), 3), dtype="<i4")
        for i, b in enumerate(self.brushes):
            br[i] = (len(sides), len(b.sides), b.shader)
            sides.extend(b.sides)
        lumps[LUMP_BRUSHES] = br.tobytes()
        lumps[LUMP_BRUSHSIDES] = np.array(sides, dtype="<i4").reshape(-1, 2).tobytes()

        # surfaces + drawverts
        dverts = []
        surf = np.zeros((len(surfaces), 26), dtype="<i4")
        for s, (kind, idx) in enumerate(surfaces):
            if kind == "planar":
                surf[s, 2] = MST_PLANAR
                continue
            p = self.patches[idx]
            surf[s, 0] = p.shader
            surf[s, 1] = -1
            surf[s, 2] = MST_PATCH
            surf[s, 3] = sum(v.shape[0] for v in dverts)
            surf[s, 4] = p.points.shape[0]
            surf[s, 24] = p.width
            surf[s, 25] = p.height
            dverts.append(p.points)
        lumps[LUMP_SURFACES] = surf.tobytes()
        if dverts:
            allv = np.concatenate(dverts, axis=0)
            dv = np.zeros((allv.shape[0], 11), dtype="<f4")
            dv[:, 0:3] = allv
            dv[:, 9] = 1.0
            lumps[LUMP_DRAWVERTS] = dv.tobytes()

        # models
        def surf_range(first_patch, num):
            if num == 0:
                return 0, 0
            lo = surf_of_patch[first_patch]
            if planar_every and first_patch % planar_every == 0:
                lo -= 1
            hi = surf_of_patch[first_patch + num - 1] + 1
            return lo, hi - lo

        models = []
        fs, ns = surf_range(0, n_world_patches)
        models.append((wmins, wmaxs, fs, ns, 0, n_world_brushes))
        for m in self.models:
            fb, nb = m["first_brush"], m["num_brushes"]
            mn = bmins[fb:fb + nb].min(axis=0) if nb else np.zeros(3)
            mx = bmaxs[fb:fb + nb].max(axis=0) if nb else np.zeros(3)
            if m["num_patches"]:
                fp, npp = m["first_patch"], m["num_patches"]
                mn = np.minimum(mn, pmins[fp:fp + npp].min(axis=0)) if nb else pmins[fp:fp + npp].min(axis=0)
                mx = np.maximum(mx, pmaxs[fp:fp + npp].max(axis=0)) if nb else pmaxs[fp:fp + npp].max(axis=0)
            fs, ns = surf_range(m["first_patch"], m["num_patches"])
            models.append((mn, mx, fs, ns, fb, nb))
        mb = bytearray()
        for mn, mx, fs, ns, fb, nb in models:
            mb += struct.pack("<6f4i", *[float(x) for x in mn], *[float(x) for x in mx], fs, ns, fb, nb)
        lumps[LUMP_MODELS] = bytes(mb)

        # header + 4-byte aligned lumps
        ofs = 8 + 17 * 8
        table = []
        body = bytearray()
        for l in lumps:
            table.append((ofs, len(l)))
            body += l
            pad = (-len(l)) % 4
            body += b"\x00" * pad
            ofs += len(l) + pad
        head = bytearray(b"IBSP" + struct.pack("<i", 46))
        for o, ln in table:
            head += struct.pack("<ii", o, ln)
        data = bytes(head + body)

        return SynthMap(
            data=data,
            world_mins=wmins.astype(np.float32), world_maxs=wmaxs.astype(np.float32),
            brush_mins=bmins[:n_world_brushes].astype(np.float32),
            brush_maxs=bmaxs[:n_world_brushes].astype(np.float32),
            num_world_brushes=n_world_brushes, num_brushes=len(self.brushes),
            num_models=len(models),
            model_mins=np.array([m[0] for m in models], dtype=np.float32),
            model_maxs=np.array([m[1] for m in models], dtype=np.float32),
            patch_mins=pmins.astype(np.float32), patch_maxs=pmaxs.astype(np.float32),
            stats=stats)


# ----------------------------------------------------------------------------
# geometry helpers
# ----------------------------------------------------------------------------
def _rotation(yaw_deg, pitch_deg=0.0, roll_deg=0.0) -> np.ndarray:
    y, p, r = (math.radians(v) for v in (yaw_deg, pitch_deg, roll_deg))
    rz = np.array([[math.cos(y), -math.sin(y), 0], [math.sin(y), math.cos(y), 0], [0, 0, 1]])
    ry = np.array([[math.cos(p), 0, math.sin(p)], [0, 1, 0], [-math.sin(p), 0, math.cos(p)]])
    rx = np.array([[1, 0, 0], [0, math.cos(r), -math.sin(r)], [0, math.sin(r), math.cos(r)]])
    return rz @ ry @ rx


def _hull_vertices(normals, dists) -> np.ndarray:
    """Vertices of the convex polytope {x : n_i.x <= d_i} by triple-plane intersection."""
    n = normals.shape[0]
    verts = []
    for i in range(n):
        for j in range(i + 1, n):
            for k in range(j + 1, n):
                a = np.stack([normals[i], normals[j], normals[k]])
                det = np.linalg.det(a)
                if abs(det) < 1e-9:
                    continue
                x = np.linalg.solve(a, np.array([dists[i], dists[j], dists[k]]))
                if np.all(normals @ x - dists <= 1e-6):
                    verts.append(x)
    return np.array(verts).reshape(-1, 3)


# ----------------------------------------------------------------------------
# map recipes
# ----------------------------------------------------------------------------
def _room_walls(mb: MapBuilder, mins, maxs, thick=64.0, shader=0):
    mins = np.asarray(mins, dtype=np.float64)
    maxs = np.asarray(maxs, dtype=np.float64)
    for a in range(3):
        lo = mins - thick
        hi = maxs + thick
        hi_a = hi.copy()
        hi_a[a] = mins[a]
        mb.box_brush(lo, hi_a, shader)
        lo_a = lo.copy()
        lo_a[a] = maxs[a]
        mb.box_brush(lo_a, hi, shader)
    mb.world_mins = mins - thick
    mb.world_maxs = maxs + thick


def _scatter_brushes(mb, rng, n, lo, hi, size_lo, size_hi, rotated_frac, shaders_p):
    lo = np.asarray(lo, dtype=np.float64)
    hi = np.asarray(hi, dtype=np.float64)
    shader_ids = np.array([s for s, _ in shaders_p])
    shader_w = np.array([w for _, w in shaders_p], dtype=np.float64)
    shader_w /= shader_w.sum()
    for _ in range(n):
        size = np.floor(rng.uniform(size_lo, size_hi, size=3))
        c = np.floor(rng.uniform(lo + size * 0.5, hi - size * 0.5))
        shader = int(rng.choice(shader_ids, p=shader_w))
        u = rng.random()
        if u < rotated_frac * 0.6:
            pitch = float(rng.uniform(-30, 30)) if rng.random() < 0.5 else 0.0
            roll = float(rng.uniform(-20, 20)) if rng.random() < 0.25 else 0.0
            mb.rotated_box_brush(c, size * 0.5, float(rng.uniform(0, 360)), pitch, roll, shader)
        elif u < rotated_frac:
            mn = c - np.floor(size * 0.5)
            mb.wedge_brush(mn, mn + size, axis=int(rng.integers(2)), flip=bool(rng.integers(2)), shader=shader)
        else:
            mn = c - np.floor(size * 0.5)
            mb.box_brush(mn, mn + size, shader)


SOLID_MIX = [(0, 0.55), (1, 0.2), (2, 0.1), (3, 0.07), (4, 0.05), (5, 0.03)]


def room_map(n_brushes=500, seed=0xC0FFEE01, extent=4096.0) -> SynthMap:
    """Config 1: closed axial room with ~n axial boxes, point traces, MASK_SOLID."""
    rng = np.random.default_rng(seed)
    mb = MapBuilder()
    h = extent / 2
    _room_walls(mb, [-h, -h, -h], [h, h, h])
    _scatter_brushes(mb, rng, n_brushes - 6, [-h, -h, -h], [h, h, h], 32, 256, 0.0, [(0, 0.7), (1, 0.3)])
    return mb.finish(seed=seed, leaf_max=4)


def arena_map(n_brushes=20000, seed=0xC0FFEE02, size_xy=16384.0, size_z=4096.0,
              rotated_frac=0.10, nonaxial_node_prob=0.02, n_models=0, model_seed=None) -> SynthMap:
    """Configs 2/4/5: arena with floor/walls, pillars, platforms and rotated/ramp brushes."""
    rng = np.random.default_rng(seed)
    mb = MapBuilder()
    hx = size_xy / 2
    _room_walls(mb, [-hx, -hx, 0.0], [hx, hx, size_z])
    n = n_brushes - 6
    n_pillars = n // 5
    # pillars: tall thin boxes standing on the floor
    for _ in range(n_pillars):
        w = np.floor(rng.uniform(24, 96, size=2))
        hgt = float(np.floor(rng.uniform(128, 768)))
        c = np.floor(rng.uniform([-hx + 128, -hx + 128], [hx - 128, hx - 128]))
        mb.box_brush([c[0] - w[0] / 2, c[1] - w[1] / 2, 0.0], [c[0] + w[0] / 2, c[1] + w[1] / 2, hgt],
                     int(rng.choice([0, 1])))
    # everything else within the lower slab where movement happens
    _scatter_brushes(mb, rng, n - n_pillars, [-hx, -hx, 0.0], [hx, hx, min(size_z, 2048.0)],
                     24, 224, rotated_frac, SOLID_MIX)
    if n_models:
        _add_models(mb, np.random.default_rng(model_seed if model_seed is not None else seed + 77), n_models)
    return mb.finish(seed=seed, leaf_max=4, nonaxial_node_prob=nonaxial_node_prob)


def _add_models(mb: MapBuilder, rng, n_models):
    """Inline models (doors / platforms) of 1-8 brushes in model-local coordinates."""
    for _ in range(n_models):
        mb.begin_model()
        nb = int(rng.integers(1, 9))
        for _ in range(nb):
            size = np.floor(rng.uniform(16, 192, size=3))
            c = np.floor(rng.uniform(-128, 128, size=3))
            u = rng.random()
            shader = int(rng.choice([7, 0, 3]))
            if u < 0.2:
                mb.rotated_box_brush(c, size * 0.5, float(rng.uniform(0, 360)), 0.0, 0.0, shader)
            elif u < 0.3:
                mn = c - np.floor(size * 0.5)
                mb.wedge_brush(mn, mn + size, axis=int(rng.integers(2)), flip=bool(rng.integers(2)), shader=shader)
            else:
                mn = c - np.floor(size * 0.5)
                mb.box_brush(mn, mn + size, shader)
        mb.end_model()


def _patch_points(rng, kind, origin, scale):
    """Control grids: bumpy terrain, cylinder (wraps), arch.  Returns (w, h, pts[h*w,3])."""
    if kind == 0:     # terrain sheet
        w = int(rng.choice([3, 5, 7, 9]))
        h = int(rng.choice([3, 5, 7, 9]))
        xs = np.linspace(-scale, scale, w)
        ys = np.linspace(-scale, scale, h)
        pts = np.zeros((h, w, 3))
        pts[:, :, 0] = xs[None, :]
        pts[:, :, 1] = ys[:, None]
        pts[:, :, 2] = rng.uniform(-0.35 * scale, 0.35 * scale, size=(h, w))
    elif kind == 1:   # cylinder: 9 columns around, first == last (wrapWidth)
        w, h = 9, int(rng.choice([3, 5]))
        ang = np.linspace(0, 2 * math.pi, 9)
        # quadratic-bezier circle: odd columns are pushed out to the corners
        rad = np.where(np.arange(9) % 2 == 1, scale * 0.5 * math.sqrt(2.0), scale * 0.5)
        zs = np.linspace(-scale, scale, h)
        pts = np.zeros((h, w, 3))
        pts[:, :, 0] = (rad * np.cos(ang))[None, :]
        pts[:, :, 1] = (rad * np.sin(ang))[None, :]
        pts[:, :, 2] = zs[:, None]
        pts[:, 8, :] = pts[:, 0, :]
    else:             # arch / half pipe
        w, h = 5, int(rng.choice([3, 5, 7]))
        ang = np.linspace(0, math.pi, 5)
        rad = np.where(np.arange(5) % 2 == 1, scale * math.sqrt(2.0), scale)
        ys = np.linspace(-scale, scale, h)
        pts = np.zeros((h, w, 3))
        pts[:, :, 0] = (rad * np.cos(ang))[None, :]
        pts[:, :, 2] = (rad * np.sin(ang))[None, :]
        pts[:, :, 1] = ys[:, None]
    pts = np.round(pts + np.asarray(origin)[None, None, :])
    return w, h, pts.reshape(-1, 3)


def patch_map(n_brushes=2000, n_patches=512, seed=0xC0FFEE03, size_xy=8192.0, size_z=2048.0,
              n_models=0) -> SynthMap:
    """Config 3: brushes plus Bezier patches (terrain sheets, cylinders, arches)."""
    rng = np.random.default_rng(seed)
    mb = MapBuilder()
    hx = size_xy / 2
    _room_walls(mb, [-hx, -hx, 0.0], [hx, hx, size_z])
    _scatter_brushes(mb, rng, n_brushes - 6, [-hx, -hx, 0.0], [hx, hx, size_z * 0.75],
                     32, 256, 0.1, SOLID_MIX)
    for _ in range(n_patches):
        scale = float(rng.uniform(48, 256))
        origin = np.floor(rng.uniform([-hx + 300, -hx + 300, 300], [hx - 300, hx - 300, size_z * 0.75]))
        w, h, pts = _patch_points(rng, int(rng.integers(3)), origin, scale)
        mb.patch(w, h, pts, shader=int(rng.choice([6, 0, 3])))
    if n_models:
        mrng = np.random.default_rng(seed + 77)
        for m in range(n_models):
            mb.begin_model()
            size = np.floor(rng.uniform(32, 128, size=3))
            mb.box_brush(-size, size, 7)
            if m % 2 == 0:
                w, h, pts = _patch_points(mrng, 0, np.zeros(3), float(mrng.uniform(64, 160)))
                mb.patch(w, h, pts, shader=6)
            mb.end_model()
    return mb.finish(seed=seed, leaf_max=4, nonaxial_node_prob=0.02, planar_every=5)


def tiny_map(seed=1) -> SynthMap:
    """A handful of brushes; the whole tree is one or two nodes. For edge-case tests."""
    mb = MapBuilder()
    _room_walls(mb, [-256, -256, -256], [256, 256, 256])
    mb.box_brush([-32, -32, -32], [32, 32, 32], 0)
    mb.rotated_box_brush([128, 0, 0], [24, 40, 16], 30.0, 15.0, 0.0, 1)
    mb.wedge_brush([-200, -64, -256], [-100, 64, -200], axis=0, shader=2)
    return mb.finish(seed=seed, leaf_max=2, nonaxial_node_prob=0.5)


def with_visibility(m: SynthMap, seed=0x715, n_areas=5, solid_frac=0.3, density=0.25, vised=True) -> SynthMap:
    """The same map with clusters, areas and a visibility lump (finish() gives every leaf its own cluster,
    area 0 and no lump).  A fraction of the leafs becomes "solid" (cluster -1, area -1, as q3map marks leafs inside
    brushes), the others get consecutive clusters and one of n_areas areas; the matrix is random with the given
    density plus the diagonal.  Nothing here is spatially meaningful -- it feeds CM_BoxLeafnums' lastLeaf rule,
    SV_LinkEntity's cluster/area bookkeeping and the PVS test (cm_test.c:73-88, sv_world.c:255-303,
    sv_snapshot.c:316-355), none of which looks at geometry."""
    rng = np.random.default_rng(seed)
    data = bytearray(m.data)
    lofs, llen = struct.unpack_from("<ii", data, 8 + 8 * LUMP_LEAFS)
    lf = np.frombuffer(bytes(data[lofs:lofs + llen]), dtype="<i4").reshape(-1, 12).copy()
    n = lf.shape[0]
    solid = rng.random(n) < solid_frac
    solid[0] = False
    open_ = np.flatnonzero(~solid)
    lf[:, 0] = -1
    lf[:, 1] = -1
    lf[open_, 0] = np.arange(open_.size)
    # areas in runs of leaf numbers, so that boxes usually touch one or two and sometimes three or more
    run = max(1, open_.size // max(4 * n_areas, 1))
    lf[open_, 1] = (np.arange(open_.size) // run) % n_areas
    data[lofs:lofs + llen] = lf.astype("<i4").tobytes()
    if vised:
        nc = int(open_.size)
        cb = ((nc + 63) & ~63) >> 3
        vis = np.zeros((nc, cb), np.uint8)
        for r0 in range(0, nc, 512):                  # in slabs of rows: the matrix of a 20 k-brush arena is ~50 MB
            r1 = min(nc, r0 + 512)
            bits = rng.random((r1 - r0, cb * 8), dtype=np.float32) < density
            bits[np.arange(r1 - r0), np.arange(r0, r1)] = True
            bits[:, nc:] = False
            vis[r0:r1] = np.packbits(bits, axis=1, bitorder="little")
        lump = struct.pack("<ii", nc, cb) + vis.tobytes()
        pad = (-len(data)) % 4
        data += b"\x00" * pad
        struct.pack_into("<ii", data, 8 + 8 * LUMP_VISIBILITY, len(data), len(lump))
        data += lump
    out = SynthMap(**{k: getattr(m, k) for k in m.__dataclass_fields__})
    out.data = bytes(data)
    out.stats = dict(m.stats, clusters=int(open_.size), areas=n_areas, vised=bool(vised))
    return out


# ----------------------------------------------------------------------------
# trace streams
# ----------------------------------------------------------------------------
def _unit_dirs(rng, n):
    v = rng.normal(size=(n, 3))
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    return v


def make_sweeps(m: SynthMap, n, seed, len_short=(1.0, 512.0), len_long=(512.0, 4096.0), long_frac=0.3,
                in_solid_frac=0.05, zero_len_frac=0.02, z_range=None, margin=32.0):
    """Config-2 style sweep stream: returns float32 starts[n,3], ends[n,3]."""
    rng = np.random.default_rng(seed)
    lo = m.world_mins.astype(np.float64) + 64.0 + margin
    hi = m.world_maxs.astype(np.float64) - 64.0 - margin
    if z_range is not None:
        lo[2], hi[2] = z_range
    starts = rng.uniform(lo, hi, size=(n, 3))
    # some starts inside solids (startsolid / allsolid paths)
    k = int(n * in_solid_frac)
    if k and m.num_world_brushes:
        which = rng.integers(0, m.num_world_brushes, size=k)
        t = rng.random(size=(k, 3))
        starts[:k] = m.brush_mins[which] + t * (m.brush_maxs[which] - m.brush_mins[which])
        starts[:k] = np.clip(starts[:k], m.world_mins + 1.0, m.world_maxs - 1.0)
    length = np.where(rng.random(n) < long_frac, rng.uniform(*len_long, size=n), rng.uniform(*len_short, size=n))
    ends = starts + _unit_dirs(rng, n) * length[:, None]
    perm = rng.permutation(n)
    starts = starts[perm].astype(np.float32)
    ends = ends[perm].astype(np.float32)
    z = int(n * zero_len_frac)
    if z:
        idx = rng.choice(n, size=z, replace=False)
        ends[idx] = starts[idx]
    return np.ascontiguousarray(starts), np.ascontiguousarray(ends)


def make_point_sweeps(m: SynthMap, n, seed, length=(64.0, 2048.0)):
    """Config-1 style point traces inside the room."""
    return make_sweeps(m, n, seed, len_short=length, len_long=length, long_frac=0.0,
                       in_solid_frac=0.0, zero_len_frac=0.0)


def make_patch_sweeps(m: SynthMap, n, seed, aim_frac=0.6):
    """Config-3 style: a share of the traces is aimed at patches so that they get hit."""
    rng = np.random.default_rng(seed)
    starts, ends = make_sweeps(m, n, seed + 1, len_short=(16.0, 512.0), len_long=(512.0, 2048.0))
    k = int(n * aim_frac)
    if k and m.patch_mins.shape[0]:
        which = rng.integers(0, m.patch_mins.shape[0], size=k)
        c = 0.5 * (m.patch_mins[which] + m.patch_maxs[which])
        r = 0.5 * np.linalg.norm(m.patch_maxs[which] - m.patch_mins[which], axis=1)
        d = _unit_dirs(rng, k)
        jitter = rng.uniform(-0.4, 0.4, size=(k, 3)) * r[:, None]
        s = c + jitter - d * (r * rng.uniform(0.2, 2.0, size=k))[:, None]
        e = c + jitter + d * (r * rng.uniform(0.0, 1.5, size=k))[:, None]
        idx = rng.choice(n, size=k, replace=False)
        starts[idx] = s.astype(np.float32)
        ends[idx] = e.astype(np.float32)
    hull = (rng.random(n) < 0.5)
    mins = np.where(hull[:, None], PLAYER_MINS[None, :], 0.0).astype(np.float32)
    maxs = np.where(hull[:, None], PLAYER_MAXS[None, :], 0.0).astype(np.float32)
    return starts, ends, np.ascontiguousarray(mins), np.ascontiguousarray(maxs)


def make_entity_traces(m: SynthMap, n, seed, box_frac=0.1, rotated_frac=0.5):
    """Config-4 style transformed traces against inline models / temp boxes.

    Returns dict of arrays: starts, ends, models, origins, angles, boxmins, boxmaxs.
    """
    rng = np.random.default_rng(seed)
    nm = m.num_models
    models = rng.integers(1, max(nm, 2), size=n).astype(np.int32) if nm > 1 else np.full(n, BOX_MODEL_HANDLE, np.int32)
    isbox = rng.random(n) < (box_frac if nm > 1 else 1.0)
    models[isbox] = BOX_MODEL_HANDLE
    lo = m.world_mins.astype(np.float64) + 256
    hi = m.world_maxs.astype(np.float64) - 256
    origins = np.floor(rng.uniform(lo, hi, size=(n, 3)))
    angles = np.zeros((n, 3))
    rot = rng.random(n) < rotated_frac
    full = rng.random(n) < 0.5
    angles[rot, 1] = rng.uniform(0, 360, size=int(rot.sum()))
    both = rot & full
    angles[both, 0] = rng.uniform(-180, 180, size=int(both.sum()))
    angles[both, 2] = rng.uniform(-180, 180, size=int(both.sum()))
    half = rng.uniform(8, 64, size=(n, 3))
    boxmins = np.floor(-half)
    boxmaxs = np.floor(half * rng.uniform(0.5, 1.5, size=(n, 3))) + 1
    # aim the sweep at the entity
    d = _unit_dirs(rng, n)
    reach = rng.uniform(32, 400, size=n)
    off = rng.uniform(-96, 96, size=(n, 3))
    starts = origins + off - d * reach[:, None]
    ends = origins + off + d * (reach * rng.uniform(-0.2, 1.2, size=n))[:, None]
    z = rng.random(n) < 0.03
    ends[z] = starts[z]
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    return dict(starts=f32(starts), ends=f32(ends), models=np.ascontiguousarray(models),
                origins=f32(origins), angles=f32(angles), boxmins=f32(boxmins), boxmaxs=f32(boxmaxs))


def make_points(m: SynthMap, n, seed, in_solid_frac=0.3):
    """Point-contents queries: uniform in the world plus a share inside brushes."""
    rng = np.random.default_rng(seed)
    pts = rng.uniform(m.world_mins.astype(np.float64), m.world_maxs.astype(np.float64), size=(n, 3))
    k = int(n * in_solid_frac)
    if k and m.num_world_brushes:
        which = rng.integers(0, m.num_world_brushes, size=k)
        t = rng.random(size=(k, 3))
        pts[:k] = m.brush_mins[which] + t * (m.brush_maxs[which] - m.brush_mins[which])
    # a few exactly on brush faces / integer lattice (ties in d > dist)
    j = n // 20
    if j:
        pts[-j:] = np.floor(pts[-j:])
    return np.ascontiguousarray(pts[rng.permutation(n)], dtype=np.float32)
