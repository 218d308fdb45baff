"""ctypes front-end for oracle/_ref/libcmref.so -- TEST INFRASTRUCTURE.

libcmref.so is the UNMODIFIED reference clip-map code (code/qcommon/cm_*.c,
q_math.c, q_shared.c, and code/server/sv_world.c) built by oracle/Makefile plus
oracle/ref_shim.c and oracle/sv_shim.c.  It is
the checker for parity tests and the timed subject of ``bench.py --impl
reference`` / the ``cpu_baseline`` leg.  Product code never imports this.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "_ref", "libcmref.so")

TRACE_DTYPE = np.dtype([
    ("allsolid", "<i4"), ("startsolid", "<i4"), ("fraction", "<f4"), ("endpos", "<f4", (3,)),
    ("plane_normal", "<f4", (3,)), ("plane_dist", "<f4"), ("plane_type", "u1"),
    ("plane_signbits", "u1"), ("plane_pad", "u1", (2,)),
    ("surfaceFlags", "<i4"), ("contents", "<i4"), ("entityNum", "<i4")])
assert TRACE_DTYPE.itemsize == 56     # q_shared.h:976-986

# aas_trace_t, botlib/be_aas.h:84-93
AAS_TRACE_DTYPE = np.dtype([("startsolid", "<i4"), ("fraction", "<f4"), ("endpos", "<f4", (3,)), ("ent", "<i4"),
                            ("lastarea", "<i4"), ("area", "<i4"), ("planenum", "<i4")])
assert AAS_TRACE_DTYPE.itemsize == 36
# aas_clientmove_t, botlib/be_aas.h:178-189
AAS_MOVE_DTYPE = np.dtype([("endpos", "<f4", (3,)), ("endarea", "<i4"), ("velocity", "<f4", (3,)), ("trace", AAS_TRACE_DTYPE),
                           ("presencetype", "<i4"), ("stopevent", "<i4"), ("endcontents", "<i4"), ("time", "<f4"), ("frames", "<i4")])
assert AAS_MOVE_DTYPE.itemsize == 84


def available() -> bool:
    return os.path.exists(LIB_PATH)


def build(quiet=True) -> bool:
    """(Re)build via the Makefile when the reference tree is present."""
    import subprocess
    r = subprocess.run(["make", "-C", HERE, "ref"], capture_output=quiet, text=True)
    return r.returncode == 0 and available()


def fork_map(n: int, out_dtype, work, nworkers: int | None = None) -> np.ndarray:
    """Run ``work(lo, hi) -> array[hi - lo]`` over contiguous slices of range(n) in fork()ed children, one per host
    core, and return the concatenated result.  The reference's CM layer is not re-entrant (global checkcount, shared temp
    box: cm_trace.c:553, cm_load.c:927-949), so parallel checking needs processes, not threads; a child inherits the map
    the parent loaded.  Results come back through an anonymous shared mapping."""
    import mmap
    import traceback
    out_dtype = np.dtype(out_dtype)
    if nworkers is None:
        try:
            nworkers = len(os.sched_getaffinity(0))
        except Exception:
            nworkers = os.cpu_count() or 1
    nworkers = max(1, min(int(nworkers), (n + 4095) // 4096))
    if nworkers == 1 or n == 0:
        return np.asarray(work(0, n), dtype=out_dtype).reshape(n)
    buf = mmap.mmap(-1, n * out_dtype.itemsize)
    out = np.frombuffer(buf, dtype=out_dtype, count=n)
    pids = []
    for w in range(nworkers):
        lo, hi = n * w // nworkers, n * (w + 1) // nworkers
        pid = os.fork()
        if pid == 0:
            rc = 1
            try:
                out[lo:hi] = np.asarray(work(lo, hi), dtype=out_dtype).reshape(hi - lo)
                rc = 0
            except BaseException:
                traceback.print_exc()
            finally:
                os._exit(rc)
        pids.append(pid)
    bad = [p for p in pids if os.waitpid(p, 0)[1] != 0]
    res = out.copy()
    del out
    buf.close()
    if bad:
        raise RuntimeError(f"fork_map: {len(bad)} of {nworkers} workers failed")
    return res


_f = lambda a: np.ascontiguousarray(a, dtype=np.float32)
_i = lambda a: np.ascontiguousarray(a, dtype=np.int32)
_p = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None


class RefCM:
    """One loaded map in the reference library (the library is a process-wide singleton)."""

    def __init__(self):
        if not available():
            raise RuntimeError("oracle/_ref/libcmref.so missing: run `make -C oracle ref` where /root/reference exists")
        self.lib = C.CDLL(LIB_PATH)
        self.lib.cmref_last_error.restype = C.c_char_p
        self.lib.cmref_shared_alloc.restype = C.c_void_p
        self.lib.cmref_shared_alloc.argtypes = [C.c_size_t]
        self.lib.cmref_shared_free.argtypes = [C.c_void_p, C.c_size_t]
        assert self.lib.cmref_sizeof_trace() == 56

    def _check(self, rc, what):
        if rc != 0:
            raise RuntimeError(f"{what}: {self.lib.cmref_last_error().decode(errors='replace')}")

    def load(self, data: bytes):
        self._data = bytes(data)
        self._check(self.lib.cmref_load(self._data, len(self._data)), "CM_LoadMap")
        return self

    def set_cvars(self, noCurves=0, playerCurveClip=1):
        self.lib.cmref_set_cvars(int(noCurves), int(playerCurveClip))

    def num_inline_models(self):
        return self.lib.cmref_num_inline_models()

    @staticmethod
    def _hull(n, mins, maxs):
        if mins is None:
            return None, None, 0
        mins = _f(mins)
        maxs = _f(maxs)
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

    def box_trace(self, starts, ends, mins=None, maxs=None, model=0, mask=1, boxmins=None, boxmaxs=None):
        starts, ends = _f(starts), _f(ends)
        n = starts.shape[0]
        mins, maxs, hs = self._hull(n, mins, maxs)
        m, ms = self._mask(n, mask)
        models = None
        if not np.isscalar(model) or model != 0:
            models = _i(np.broadcast_to(_i(model), (n,)))
        out = np.zeros(n, dtype=TRACE_DTYPE)
        boxmins = _f(boxmins) if boxmins is not None else None
        boxmaxs = _f(boxmaxs) if boxmaxs is not None else None
        self._check(self.lib.cmref_box_trace_batch(n, _p(starts), _p(ends), _p(mins), _p(maxs), hs,
                                                   _p(models), _p(m), ms, _p(boxmins), _p(boxmaxs), _p(out)),
                    "CM_BoxTrace")
        return out

    def transformed_box_trace(self, starts, ends, mins, maxs, models, mask, origins, angles,
                              boxmins=None, boxmaxs=None):
        starts, ends = _f(starts), _f(ends)
        n = starts.shape[0]
        mins, maxs, hs = self._hull(n, mins, maxs)
        m, ms = self._mask(n, mask)
        models = _i(np.broadcast_to(_i(models), (n,)))
        origins, angles = _f(origins), _f(angles)
        if boxmins is None:
            boxmins = np.zeros((n, 3), np.float32)
            boxmaxs = np.zeros((n, 3), np.float32)
        boxmins, boxmaxs = _f(boxmins), _f(boxmaxs)
        out = np.zeros(n, dtype=TRACE_DTYPE)
        self._check(self.lib.cmref_transformed_box_trace_batch(
            n, _p(starts), _p(ends), _p(mins), _p(maxs), hs, _p(models), _p(m), ms,
            _p(origins), _p(angles), _p(boxmins), _p(boxmaxs), _p(out)), "CM_TransformedBoxTrace")
        return out

    def point_contents(self, pts, models=None, boxmins=None, boxmaxs=None):
        pts = _f(pts)
        n = pts.shape[0]
        if models is not None:
            models = _i(np.broadcast_to(_i(models), (n,)))
            if boxmins is None:
                boxmins = np.zeros((n, 3), np.float32)
                boxmaxs = np.zeros((n, 3), np.float32)
            boxmins, boxmaxs = _f(boxmins), _f(boxmaxs)
        out = np.zeros(n, dtype=np.int32)
        self._check(self.lib.cmref_point_contents_batch(n, _p(pts), _p(models), _p(boxmins), _p(boxmaxs), _p(out)),
                    "CM_PointContents")
        return out

    def transformed_point_contents(self, pts, models, origins, angles, boxmins=None, boxmaxs=None):
        pts = _f(pts)
        n = pts.shape[0]
        models = _i(np.broadcast_to(_i(models), (n,)))
        origins, angles = _f(origins), _f(angles)
        if boxmins is None:
            boxmins = np.zeros((n, 3), np.float32)
            boxmaxs = np.zeros((n, 3), np.float32)
        boxmins, boxmaxs = _f(boxmins), _f(boxmaxs)
        out = np.zeros(n, dtype=np.int32)
        self._check(self.lib.cmref_transformed_point_contents_batch(
            n, _p(pts), _p(models), _p(origins), _p(angles), _p(boxmins), _p(boxmaxs), _p(out)),
            "CM_TransformedPointContents")
        return out

    def point_leafnum(self, pts):
        pts = _f(pts)
        out = np.zeros(pts.shape[0], dtype=np.int32)
        self._check(self.lib.cmref_point_leafnum_batch(pts.shape[0], _p(pts), _p(out)), "CM_PointLeafnum")
        return out

    def box_leafnums(self, mins, maxs, listsize=64):
        mins, maxs = _f(mins), _f(maxs)
        n = mins.shape[0]
        lists = np.full((n, listsize), -1, dtype=np.int32)
        counts = np.zeros(n, np.int32)
        last = np.zeros(n, np.int32)
        self._check(self.lib.cmref_box_leafnums_batch(n, _p(mins), _p(maxs), listsize, _p(lists), _p(counts), _p(last)),
                    "CM_BoxLeafnums")
        return lists, counts, last

    def angle_vectors(self, angles):
        a = _f(angles)
        f = np.zeros(3, np.float32)
        r = np.zeros(3, np.float32)
        u = np.zeros(3, np.float32)
        self.lib.cmref_angle_vectors(_p(a), _p(f), _p(r), _p(u))
        return f, r, u

    def bench_box_trace(self, starts, ends, mins, maxs, mask, nworkers, repeats=1, want_results=False):
        """All-core throughput run: fork()ed workers, one per core, slowest worker's time."""
        starts, ends = _f(starts), _f(ends)
        n = starts.shape[0]
        mins = _f(mins) if mins is not None else None
        maxs = _f(maxs) if maxs is not None else None
        secs = C.c_double(0.0)
        wsecs = np.zeros(max(nworkers, 1), dtype=np.float64)
        shared = None
        nbytes = n * 56
        if want_results:
            shared = self.lib.cmref_shared_alloc(nbytes)
        rc = self.lib.cmref_box_trace_bench(n, _p(starts), _p(ends), _p(mins), _p(maxs), int(mask),
                                            int(nworkers), int(repeats), C.c_void_p(shared),
                                            C.byref(secs), _p(wsecs))
        out = None
        if want_results and shared:
            buf = (C.c_char * nbytes).from_address(shared)
            out = np.frombuffer(bytes(buf), dtype=TRACE_DTYPE).copy()
            self.lib.cmref_shared_free(C.c_void_p(shared), nbytes)
        self._check(rc, "bench")
        return secs.value, wsecs[:nworkers], out

    # ---- the reference's server world (server/sv_world.c, compiled verbatim; oracle/sv_shim.c) ------------
    def sv_clear_world(self):
        """SV_ClearWorld after zeroing all MAX_GENTITIES (4096) gentities (ownerNum = ENTITYNUM_NONE)"""
        self.lib.svref_clear_world()

    def sv_link_entity(self, num, bmodel, modelindex, mins, maxs, origin, angles, contents, owner=4095):
        """set the fields the game owns, then SV_LinkEntity; returns r.linked"""
        mins, maxs, origin, angles = _f(mins), _f(maxs), _f(origin), _f(angles)
        return self.lib.svref_link_entity(int(num), int(bool(bmodel)), int(modelindex), _p(mins), _p(maxs), _p(origin),
                                          _p(angles), int(contents), int(owner))

    def sv_set_owner(self, num, owner):
        self.lib.svref_set_owner(int(num), int(owner))

    def sv_unlink_entity(self, num):
        self.lib.svref_unlink_entity(int(num))

    def sv_entity_bounds(self, num):
        a, b = np.zeros(3, np.float32), np.zeros(3, np.float32)
        linked = self.lib.svref_entity_bounds(int(num), _p(a), _p(b))
        return a, b, linked

    def sv_area_entities(self, mins, maxs, maxcount=4096):
        mins, maxs = _f(mins), _f(maxs)
        lst = np.zeros(maxcount, np.int32)
        n = self.lib.svref_area_entities(_p(mins), _p(maxs), _p(lst), int(maxcount))
        return lst[:n].copy()

    def sv_trace(self, starts, ends, mins, maxs, pass_ents, masks):
        """SV_Trace per item (sv_world.c:562-609)"""
        starts, ends = _f(starts), _f(ends)
        n = starts.shape[0]
        mins, maxs, hs = self._hull(n, mins, maxs)
        pe = np.asarray(pass_ents)
        pe, ps = (_i(pe.reshape(1)), 0) if pe.ndim == 0 else (_i(pe), 1)
        m, ms = self._mask(n, masks)
        out = np.zeros(n, dtype=TRACE_DTYPE)
        self.lib.svref_trace_batch(_p(out), n, _p(starts), _p(ends), _p(mins), _p(maxs), hs, _p(pe), ps, _p(m), ms)
        return out

    def sv_point_contents(self, pts, pass_ents):
        pts = _f(pts)
        pe = np.asarray(pass_ents)
        pe, ps = (_i(pe.reshape(1)), 0) if pe.ndim == 0 else (_i(pe), 1)
        out = np.zeros(pts.shape[0], np.int32)
        self.lib.svref_point_contents_batch(_p(out), pts.shape[0], _p(pts), _p(pe), ps)
        return out

    def sv_time_traces(self, starts, ends, mins, maxs, pass_ents, mask):
        starts, ends, mins, maxs, pe = _f(starts), _f(ends), _f(mins), _f(maxs), _i(pass_ents)
        self.lib.svref_time_traces.restype = C.c_double
        return self.lib.svref_time_traces(starts.shape[0], _p(starts), _p(ends), _p(mins), _p(maxs), _p(pe), int(mask))

    # ---- clusters, areas, visibility (cm_load.c:838-860, cm_test.c:306-470) and what uses them ---------------
    def leaf_cluster_area(self, leafs):
        leafs = _i(leafs)
        c, a = np.zeros(leafs.shape[0], np.int32), np.zeros(leafs.shape[0], np.int32)
        self._check(self.lib.cmref_leaf_cluster_area(leafs.shape[0], _p(leafs), _p(c), _p(a)), "CM_LeafCluster")
        return c, a

    def num_clusters(self):
        return int(self.lib.cmref_num_clusters())

    def cluster_pvs(self, cluster, nbytes):
        out = np.zeros(nbytes, np.uint8)
        self._check(self.lib.cmref_cluster_pvs(int(cluster), int(nbytes), _p(out)), "CM_ClusterPVS")
        return out

    def adjust_area_portal_state(self, a1, a2, open_):
        self._check(self.lib.cmref_adjust_area_portal(int(a1), int(a2), int(bool(open_))), "CM_AdjustAreaPortalState")

    def areas_connected(self, a1, a2):
        a1, a2 = _i(a1), _i(a2)
        out = np.zeros(a1.shape[0], np.int32)
        self._check(self.lib.cmref_areas_connected(a1.shape[0], _p(a1), _p(a2), _p(out)), "CM_AreasConnected")
        return out

    def write_area_bits(self, area, nbytes=32):
        buf = np.zeros(nbytes, np.uint8)
        n = C.c_int(0)
        self._check(self.lib.cmref_write_area_bits(int(area), _p(buf), C.byref(n)), "CM_WriteAreaBits")
        return bytes(buf[:n.value])

    def sv_entity_clusters(self, num):
        """-> (areanum, areanum2, numClusters, lastCluster, clusternums[16]) of the linked entity's svEntity_t"""
        out = np.zeros(20, np.int32)
        linked = self.lib.svref_entity_clusters(int(num), _p(out))
        return out, linked

    def sv_visible_from_points(self, origins, nums):
        origins, nums = _f(origins).reshape(-1, 3), _i(nums)
        vis = np.zeros((origins.shape[0], nums.shape[0]), np.uint8)
        self.lib.svref_visible_from_points(origins.shape[0], _p(origins), nums.shape[0], _p(nums), _p(vis))
        return vis.astype(bool)

    # ---- botlib's AAS sampling (botlib/be_aas_sample.c, compiled verbatim; oracle/aas_shim.c) ---------------------
    def aas_load(self, world):
        """world: threecore_b200.aasgen.AasWorld (planes [n,5], nodes [n,3], presence [n])"""
        planes, nodes, pres = _f(world.planes), _i(world.nodes), _i(world.presence)
        contents = _i(getattr(world, "contents", None)) if getattr(world, "contents", None) is not None else None
        assert self.lib.aasref_sizeof_trace() == AAS_TRACE_DTYPE.itemsize
        assert self.lib.aasref_sizeof_clientmove() == AAS_MOVE_DTYPE.itemsize
        self._check(self.lib.aasref_load(planes.shape[0], _p(planes), nodes.shape[0], _p(nodes), pres.shape[0], _p(pres), _p(contents)),
                    "aas load")
        return self

    def aas_predict_client_movement(self, p):
        """n x AAS_PredictClientMovement(&move, -1, ...) (be_aas_move.c:965-979); p: dict of arrays as aasgen.make_moves returns.
        The world's contents (AAS_PointContents) come from the clip map loaded with load(), if any.  -> (moves, return values)"""
        n = len(p["origins"])
        moves = np.zeros(n, dtype=AAS_MOVE_DTYPE)
        res = np.zeros(n, np.int32)
        a = {k: (_f(v) if k in ("origins", "velocities", "cmdmoves", "frametimes") else _i(v)) for k, v in p.items()}
        self.lib.aasref_predict_client_movement_batch(n, _p(a["origins"]), _p(a["presencetypes"]), _p(a["ongrounds"]), _p(a["velocities"]),
                                                      _p(a["cmdmoves"]), _p(a["cmdframes"]), _p(a["maxframes"]), _p(a["frametimes"]),
                                                      _p(a["stopevents"]), _p(a["stopareanums"]), _p(moves), _p(res))
        return moves, res

    def aas_trace_client_bbox(self, starts, ends, presencetype):
        """n x AAS_TraceClientBBox(start, end, presencetype, passent = -1) (be_aas_sample.c:399-665)"""
        starts, ends = _f(starts), _f(ends)
        n = starts.shape[0]
        pt = np.asarray(presencetype)
        pt, ps = (_i(pt.reshape(1)), 0) if pt.ndim == 0 else (_i(pt), 1)
        out = np.zeros(n, dtype=AAS_TRACE_DTYPE)
        self.lib.aasref_trace_client_bbox_batch(n, _p(starts), _p(ends), _p(pt), ps, _p(out))
        return out

    def aas_point_area_num(self, points):
        points = _f(points)
        out = np.zeros(points.shape[0], np.int32)
        self.lib.aasref_point_area_num_batch(points.shape[0], _p(points), _p(out))
        return out

    # ---- dumps of the loaded clipMap_t -----------------------------------
    COUNT_NAMES = ["planes", "nodes", "leafs", "leafbrushes", "leafsurfaces", "brushes", "brushsides",
                   "models", "surfaces", "shaders", "patches", "patchplanes", "facets", "borders"]

    def counts(self):
        c = np.zeros(len(self.COUNT_NAMES), np.int32)
        self.lib.cmref_counts(_p(c))
        return dict(zip(self.COUNT_NAMES, (int(x) for x in c)))

    def dump(self):
        n = self.counts()
        d = dict(counts=n)
        d["planes"] = np.zeros((n["planes"], 4), np.float32)
        d["plane_type"] = np.zeros(n["planes"], np.uint8)
        d["plane_signbits"] = np.zeros(n["planes"], np.uint8)
        self.lib.cmref_dump_planes(_p(d["planes"]), _p(d["plane_type"]), _p(d["plane_signbits"]))
        d["nodes"] = np.zeros((n["nodes"], 3), np.int32)
        self.lib.cmref_dump_nodes(_p(d["nodes"]))
        d["leafs"] = np.zeros((n["leafs"], 6), np.int32)
        self.lib.cmref_dump_leafs(_p(d["leafs"]))
        d["leafbrushes"] = np.zeros(n["leafbrushes"], np.int32)
        self.lib.cmref_dump_leafbrushes(_p(d["leafbrushes"]))
        d["leafsurfaces"] = np.zeros(max(n["leafsurfaces"], 1), np.int32)
        self.lib.cmref_dump_leafsurfaces(_p(d["leafsurfaces"]))
        d["leafsurfaces"] = d["leafsurfaces"][:n["leafsurfaces"]]
        d["brush_meta"] = np.zeros((n["brushes"], 4), np.int32)
        d["brush_bounds"] = np.zeros((n["brushes"], 6), np.float32)
        self.lib.cmref_dump_brushes(_p(d["brush_meta"]), _p(d["brush_bounds"]))
        d["brushsides"] = np.zeros((n["brushsides"], 3), np.int32)
        self.lib.cmref_dump_brushsides(_p(d["brushsides"]))
        d["model_bounds"] = np.zeros((n["models"], 6), np.float32)
        d["model_nums"] = np.zeros((n["models"], 2), np.int32)
        self.lib.cmref_dump_models(_p(d["model_bounds"]), _p(d["model_nums"]))
        d["model_brushlist"] = np.zeros(max(int(d["model_nums"][:, 0].sum()), 1), np.int32)
        d["model_surflist"] = np.zeros(max(int(d["model_nums"][:, 1].sum()), 1), np.int32)
        self.lib.cmref_dump_model_lists(_p(d["model_brushlist"]), _p(d["model_surflist"]))
        d["model_brushlist"] = d["model_brushlist"][:int(d["model_nums"][:, 0].sum())]
        d["model_surflist"] = d["model_surflist"][:int(d["model_nums"][:, 1].sum())]
        d["surf2patch"] = np.zeros(max(n["surfaces"], 1), np.int32)
        d["patch_meta"] = np.zeros((max(n["patches"], 1), 4), np.int32)
        d["patch_bounds"] = np.zeros((max(n["patches"], 1), 6), np.float32)
        d["patch_planes"] = np.zeros((max(n["patchplanes"], 1), 4), np.float32)
        d["patch_plane_signbits"] = np.zeros(max(n["patchplanes"], 1), np.int32)
        d["facet_meta"] = np.zeros((max(n["facets"], 1), 2), np.int32)
        d["borders"] = np.zeros((max(n["borders"], 1), 3), np.int32)
        self.lib.cmref_dump_patches(_p(d["surf2patch"]), _p(d["patch_meta"]), _p(d["patch_bounds"]),
                                    _p(d["patch_planes"]), _p(d["patch_plane_signbits"]),
                                    _p(d["facet_meta"]), _p(d["borders"]))
        d["surf2patch"] = d["surf2patch"][:n["surfaces"]]
        d["patch_meta"] = d["patch_meta"][:n["patches"]]
        d["patch_bounds"] = d["patch_bounds"][:n["patches"]]
        d["patch_planes"] = d["patch_planes"][:n["patchplanes"]]
        d["patch_plane_signbits"] = d["patch_plane_signbits"][:n["patchplanes"]]
        d["facet_meta"] = d["facet_meta"][:n["facets"]]
        d["borders"] = d["borders"][:n["borders"]]
        return d
