"""ctypes front-end for oracle/libcmoracle.so (cm_oracle.c) -- TEST INFRASTRUCTURE.

The oracle consumes the clip map as a dict of index-based numpy arrays ("flat
dict").  ``flat_from_ref_dump`` builds that dict from the reference library's
own loaded clipMap_t (oracle/cmref.py:RefCM.dump), so the oracle can be fed
without touching any product code.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from .cmref import TRACE_DTYPE

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcmoracle.so")

INFO_DTYPE = np.dtype([(k, "<i4") for k in (
    "hitBrush", "hitPatch", "nodes", "leafs", "brushRefs", "brushBounds", "brushSides",
    "crossings", "splits", "patches", "patchPlanes", "facets")])


class _Map(C.Structure):
    _fields_ = [
        ("numPlanes", C.c_int), ("planes", C.c_void_p), ("planeType", C.c_void_p), ("planeSignbits", C.c_void_p),
        ("numNodes", C.c_int), ("nodes", C.c_void_p),
        ("numLeafs", C.c_int), ("leafs", C.c_void_p),
        ("numLeafBrushes", C.c_int), ("leafbrushes", C.c_void_p),
        ("numLeafSurfaces", C.c_int), ("leafsurfaces", C.c_void_p),
        ("numBrushes", C.c_int), ("brushes", C.c_void_p), ("brushBounds", C.c_void_p),
        ("numBrushSides", C.c_int), ("brushsides", C.c_void_p),
        ("numModels", C.c_int), ("modelLeafs", C.c_void_p),
        ("numSurfaces", C.c_int), ("surf2patch", C.c_void_p),
        ("numPatches", C.c_int), ("patches", C.c_void_p), ("patchBounds", C.c_void_p),
        ("numPatchPlanes", C.c_int), ("patchPlanes", C.c_void_p), ("patchPlaneSignbits", C.c_void_p),
        ("numFacets", C.c_int), ("facets", C.c_void_p),
        ("numBorders", C.c_int), ("borders", C.c_void_p),
        ("boxBrush", C.c_int), ("boxFirstPlane", C.c_int), ("boxFirstSide", C.c_int), ("boxLeafBrush", C.c_int),
        ("noCurves", C.c_int), ("playerCurveClip", C.c_int),
    ]


def build(quiet=True) -> bool:
    r = subprocess.run(["make", "-C", HERE, "oracle"], capture_output=quiet, text=True)
    return r.returncode == 0 and os.path.exists(LIB_PATH)


def available() -> bool:
    return os.path.exists(LIB_PATH)


def flat_from_ref_dump(d: dict) -> dict:
    """Reference clipMap_t dump -> flat dict (submodel index lists appended to the shared arrays)."""
    n = d["counts"]
    leafbrushes = list(d["leafbrushes"])            # includes the box entry at the end
    leafsurfaces = list(d["leafsurfaces"])
    model_leafs = np.zeros((n["models"], 4), np.int32)
    bo = so = 0
    for i in range(1, n["models"]):
        nb, ns = (int(x) for x in d["model_nums"][i])
        model_leafs[i] = (len(leafbrushes), nb, len(leafsurfaces), ns)
        leafbrushes.extend(int(x) for x in d["model_brushlist"][bo:bo + nb])
        leafsurfaces.extend(int(x) for x in d["model_surflist"][so:so + ns])
        bo += nb
        so += ns
    patches = np.zeros((n["patches"], 6), np.int32)
    facets = np.zeros((n["facets"], 3), np.int32)
    fp = ff = fb = 0
    for p in range(n["patches"]):
        sf, ct, npl, nf = (int(x) for x in d["patch_meta"][p])
        patches[p] = (sf, ct, fp, npl, ff, nf)
        for f in range(nf):
            sp, nb = (int(x) for x in d["facet_meta"][ff + f])
            facets[ff + f] = (sp, nb, fb)
            fb += nb
        fp += npl
        ff += nf
    return dict(
        planes=np.ascontiguousarray(d["planes"], np.float32),
        plane_type=np.ascontiguousarray(d["plane_type"], np.uint8),
        plane_signbits=np.ascontiguousarray(d["plane_signbits"], np.uint8),
        nodes=np.ascontiguousarray(d["nodes"], np.int32),
        leafs=np.ascontiguousarray(d["leafs"][:, 2:6], np.int32),
        leaf_cluster_area=np.ascontiguousarray(d["leafs"][:, 0:2], np.int32),
        leafbrushes=np.array(leafbrushes, np.int32),
        leafsurfaces=np.array(leafsurfaces, np.int32),
        brushes=np.ascontiguousarray(d["brush_meta"][:, 1:4], np.int32),
        brush_bounds=np.ascontiguousarray(d["brush_bounds"], np.float32),
        brushsides=np.ascontiguousarray(d["brushsides"][:, 0:2], np.int32),
        model_leafs=model_leafs,
        model_bounds=np.ascontiguousarray(d["model_bounds"], np.float32),
        surf2patch=np.ascontiguousarray(d["surf2patch"], np.int32),
        patches=patches,
        patch_bounds=np.ascontiguousarray(d["patch_bounds"], np.float32),
        patch_planes=np.ascontiguousarray(d["patch_planes"], np.float32),
        patch_plane_signbits=np.ascontiguousarray(d["patch_plane_signbits"], np.int32),
        facets=facets,
        borders=np.ascontiguousarray(d["borders"][:, 0:2], np.int32),
        box_brush=n["brushes"] - 1,
        box_first_plane=n["planes"] - 12,
        box_first_side=n["brushsides"] - 6,
        box_leaf_brush=len(d["leafbrushes"]) - 1,
    )


_f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
_i = lambda a: np.ascontiguousarray(a, dtype=np.int32)
_p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None


class OracleCM:
    def __init__(self, flat: dict, noCurves=0, playerCurveClip=1):
        if not available() and not build():
            raise RuntimeError("oracle/libcmoracle.so missing and could not be built")
        self.lib = C.CDLL(LIB_PATH)
        self.flat = flat
        m = _Map()
        keep = []

        def arr(key, dtype):
            a = np.ascontiguousarray(flat[key], dtype=dtype)
            if a.size == 0:
                a = np.zeros(8, dtype=dtype)
            keep.append(a)
            return a.ctypes.data

        m.numPlanes = len(flat["planes"]); m.planes = arr("planes", np.float32)
        m.planeType = arr("plane_type", np.uint8); m.planeSignbits = arr("plane_signbits", np.uint8)
        m.numNodes = len(flat["nodes"]); m.nodes = arr("nodes", np.int32)
        m.numLeafs = len(flat["leafs"]); m.leafs = arr("leafs", np.int32)
        m.numLeafBrushes = len(flat["leafbrushes"]); m.leafbrushes = arr("leafbrushes", np.int32)
        m.numLeafSurfaces = len(flat["leafsurfaces"]); m.leafsurfaces = arr("leafsurfaces", np.int32)
        m.numBrushes = len(flat["brushes"]); m.brushes = arr("brushes", np.int32)
        m.brushBounds = arr("brush_bounds", np.float32)
        m.numBrushSides = len(flat["brushsides"]); m.brushsides = arr("brushsides", np.int32)
        m.numModels = len(flat["model_leafs"]); m.modelLeafs = arr("model_leafs", np.int32)
        m.numSurfaces = len(flat["surf2patch"]); m.surf2patch = arr("surf2patch", np.int32)
        m.numPatches = len(flat["patches"]); m.patches = arr("patches", np.int32)
        m.patchBounds = arr("patch_bounds", np.float32)
        m.numPatchPlanes = len(flat["patch_planes"]); m.patchPlanes = arr("patch_planes", np.float32)
        m.patchPlaneSignbits = arr("patch_plane_signbits", np.int32)
        m.numFacets = len(flat["facets"]); m.facets = arr("facets", np.int32)
        m.numBorders = len(flat["borders"]); m.borders = arr("borders", np.int32)
        m.boxBrush = int(flat["box_brush"]); m.boxFirstPlane = int(flat["box_first_plane"])
        m.boxFirstSide = int(flat["box_first_side"]); m.boxLeafBrush = int(flat["box_leaf_brush"])
        m.noCurves = int(noCurves); m.playerCurveClip = int(playerCurveClip)
        self._m = m
        self._keep = keep

    def set_cvars(self, noCurves=0, playerCurveClip=1):
        self._m.noCurves = int(noCurves)
        self._m.playerCurveClip = int(playerCurveClip)

    @staticmethod
    def _hull(n, mins, maxs):
        if mins is None:
            return None, None, 0
        mins, maxs = _f(mins), _f(maxs)
        if mins.ndim == 1:
            return mins, maxs, 0
        assert mins.shape == (n, 3)
        return mins, maxs, 1

    @staticmethod
    def _mask(n, mask):
        m = np.asarray(mask)
        if m.ndim == 0:
            return np.array([int(m)], dtype=np.int64).astype(np.int32), 0
        assert m.shape == (n,)
        return _i(m), 1

    def box_trace(self, starts, ends, mins=None, maxs=None, model=0, mask=1, boxmins=None, boxmaxs=None,
                  want_info=False):
        starts, ends = _f(starts), _f(ends)
        n = starts.shape[0]
        mins, maxs, hs = self._hull(n, mins, maxs)
        m, ms = self._mask(n, mask)
        models = None
        if not np.isscalar(model) or model != 0:
            models = _i(np.broadcast_to(_i(model), (n,)))
        boxmins = _f(boxmins) if boxmins is not None else None
        boxmaxs = _f(boxmaxs) if boxmaxs is not None else None
        out = np.zeros(n, dtype=TRACE_DTYPE)
        info = np.zeros(n, dtype=INFO_DTYPE) if want_info else None
        self.lib.cmo_box_trace_batch(C.byref(self._m), n, _p(starts), _p(ends), _p(mins), _p(maxs), hs,
                                     _p(models), _p(m), ms, _p(boxmins), _p(boxmaxs), _p(out), _p(info))
        return (out, info) if want_info else out

    def transformed_box_trace(self, starts, ends, mins, maxs, models, mask, origins, angles,
                              boxmins=None, boxmaxs=None, want_info=False):
        starts, ends = _f(starts), _f(ends)
        n = starts.shape[0]
        mins, maxs, hs = self._hull(n, mins, maxs)
        m, ms = self._mask(n, mask)
        models = _i(np.broadcast_to(_i(models), (n,)))
        origins, angles = _f(origins), _f(angles)
        boxmins = _f(boxmins) if boxmins is not None else None
        boxmaxs = _f(boxmaxs) if boxmaxs is not None else None
        out = np.zeros(n, dtype=TRACE_DTYPE)
        info = np.zeros(n, dtype=INFO_DTYPE) if want_info else None
        self.lib.cmo_transformed_box_trace_batch(C.byref(self._m), n, _p(starts), _p(ends), _p(mins), _p(maxs), hs,
                                                 _p(models), _p(m), ms, _p(origins), _p(angles),
                                                 _p(boxmins), _p(boxmaxs), _p(out), _p(info))
        return (out, info) if want_info else out

    def point_contents(self, pts, models=None, origins=None, angles=None, boxmins=None, boxmaxs=None):
        pts = _f(pts)
        n = pts.shape[0]
        if models is not None:
            models = _i(np.broadcast_to(_i(models), (n,)))
        origins = _f(origins) if origins is not None else None
        angles = _f(angles) if angles is not None else None
        boxmins = _f(boxmins) if boxmins is not None else None
        boxmaxs = _f(boxmaxs) if boxmaxs is not None else None
        out = np.zeros(n, dtype=np.int32)
        self.lib.cmo_point_contents_batch(C.byref(self._m), n, _p(pts), _p(models), _p(origins), _p(angles),
                                          _p(boxmins), _p(boxmaxs), _p(out))
        return out

    def point_leafnum(self, pts):
        pts = _f(pts)
        self.lib.cmo_point_leafnum.restype = C.c_int
        return np.array([self.lib.cmo_point_leafnum(C.byref(self._m), _p(pts[i])) for i in range(pts.shape[0])],
                        dtype=np.int32)

    # ---- leaf lists, entity clusters, areas, visibility (restated in cm_oracle.c) ----
    def _lca(self):
        a = self.flat.get("leaf_cluster_area")
        if a is None:
            return None
        if not hasattr(self, "_lca_arr"):
            self._lca_arr = np.ascontiguousarray(a, np.int32)
        return self._lca_arr

    def box_leafnums(self, mins, maxs, listsize=64):
        mins, maxs = _f(mins).reshape(-1, 3), _f(maxs).reshape(-1, 3)
        n = mins.shape[0]
        lists = np.full((n, listsize), -1, np.int32)
        counts, last = np.zeros(n, np.int32), np.zeros(n, np.int32)
        lca = self._lca()
        self.lib.cmo_box_leafnums_batch(C.byref(self._m), _p(lca), n, _p(mins), _p(maxs), int(listsize),
                                        _p(lists), _p(counts), _p(last))
        return lists, counts, last

    def entity_clusters(self, absmins, absmaxs):
        """-> int32 [n][21]: numLeafs, areanum, areanum2, numClusters, lastCluster, clusternums[16]"""
        absmins, absmaxs = _f(absmins).reshape(-1, 3), _f(absmaxs).reshape(-1, 3)
        out = np.zeros((absmins.shape[0], 21), np.int32)
        lca = self._lca()
        self.lib.cmo_entity_clusters_batch(C.byref(self._m), _p(lca), absmins.shape[0], _p(absmins), _p(absmaxs), _p(out))
        return out

    def flood_areas(self, area_portals):
        ap = np.ascontiguousarray(area_portals, np.int32)
        n = ap.shape[0]
        out = np.zeros(max(n, 1), np.int32)
        self.lib.cmo_flood_areas(n, _p(ap), _p(out))
        return out[:n]

    def visible_from_points(self, origins, ents, floodnum, num_clusters, cluster_bytes, vis):
        """ents: int32 [n][21] from entity_clusters; vis: uint8 matrix or None -> bool [nViews][nEnts]"""
        origins = _f(origins).reshape(-1, 3)
        ents = np.ascontiguousarray(ents, np.int32)
        floodnum = np.ascontiguousarray(floodnum, np.int32)
        lca = self._lca()
        leafs = self.point_leafnum(origins)
        visb = np.ascontiguousarray(vis, np.uint8) if vis is not None and np.size(vis) else None
        out = np.zeros((origins.shape[0], ents.shape[0]), bool)
        self.lib.cmo_entity_visible.restype = C.c_int
        for v in range(origins.shape[0]):
            cl, ar = int(lca[leafs[v], 0]), int(lca[leafs[v], 1])
            for e in range(ents.shape[0]):
                out[v, e] = bool(self.lib.cmo_entity_visible(floodnum.shape[0], _p(floodnum), int(num_clusters), int(cluster_bytes),
                                                             _p(visb), ar, cl, _p(np.ascontiguousarray(ents[e]))))
        return out

    def angle_vectors(self, angles):
        a = _f(angles)
        f = np.zeros(3, np.float32); r = np.zeros(3, np.float32); u = np.zeros(3, np.float32)
        self.lib.cmo_angle_vectors(_p(a), _p(f), _p(r), _p(u))
        return f, r, u


class OracleAAS:
    """The restated AAS tracer of cm_oracle.c (AAS_TraceClientBBox / AAS_PointAreaNum) over an aasgen.AasWorld."""

    def __init__(self, world):
        if not available() and not build():
            raise RuntimeError("oracle/libcmoracle.so missing and could not be built")
        self.lib = C.CDLL(LIB_PATH)
        self.planes, self.nodes, self.presence = _f(world.planes), _i(world.nodes), _i(world.presence)

    def trace_client_bbox(self, starts, ends, presencetype):
        from .cmref import AAS_TRACE_DTYPE
        starts, ends = _f(starts), _f(ends)
        n = starts.shape[0]
        pt = np.asarray(presencetype)
        pt, ps = (_i(pt.reshape(1)), 0) if pt.ndim == 0 else (_i(pt), 1)
        out = np.zeros(n, dtype=AAS_TRACE_DTYPE)
        self.lib.cmo_aas_trace_client_bbox_batch(_p(self.planes), _p(self.nodes), _p(self.presence), n, _p(starts), _p(ends),
                                                 _p(pt), ps, _p(out))
        return out

    def point_area_num(self, points):
        points = _f(points)
        out = np.zeros(points.shape[0], np.int32)
        self.lib.cmo_aas_point_area_num_batch(_p(self.planes), _p(self.nodes), points.shape[0], _p(points), _p(out))
        return out
