"""Python host binding of libcmgpu.so (include/cm_gpu.h) -- plumbing only.

The product is the CUDA library; this module loads it with ctypes, mirrors the
C structs and offers numpy (host-pointer) and torch (device-pointer) call
styles.  PyTorch is used for device memory, streams and torch.distributed.
There is NO CPU implementation here: if the library is missing or no B200 is
visible, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcmgpu.so")
BOX_MODEL_HANDLE = 1023

TRACE_DTYPE = np.dtype([
    ("allsolid", "<i4"), ("startsolid", "<i4"), ("fraction", "<f4"), ("endpos", "<f4", (3,)),
    ("plane_normal", "<f4", (3,)), ("plane_dist", "<f4"), ("plane_type", "u1"),
    ("plane_signbits", "u1"), ("plane_pad", "u1", (2,)),
    ("surfaceFlags", "<i4"), ("contents", "<i4"), ("entityNum", "<i4")])
assert TRACE_DTYPE.itemsize == 56

ERR_NAMES = {0: "OK", 1: "INVALID", 2: "NO_DEVICE", 3: "CUDA", 4: "NO_MAP", 5: "BAD_HANDLE", 6: "BAD_MAP",
             7: "NOMEM", 8: "LIMIT"}


class CMGPUError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"cmgpu error {code} ({ERR_NAMES.get(code, '?')}): {msg}")
        self.code = code


class _FlatMap(C.Structure):
    _fields_ = [
        ("numPlanes", C.c_int32), ("planes", C.c_void_p), ("planeType", C.c_void_p), ("planeSignbits", C.c_void_p),
        ("numNodes", C.c_int32), ("nodes", C.c_void_p),
        ("numLeafs", C.c_int32), ("leafs", C.c_void_p),
        ("numLeafBrushes", C.c_int32), ("leafbrushes", C.c_void_p),
        ("numLeafSurfaces", C.c_int32), ("leafsurfaces", C.c_void_p),
        ("numBrushes", C.c_int32), ("brushes", C.c_void_p), ("brushBounds", C.c_void_p),
        ("numBrushSides", C.c_int32), ("brushsides", C.c_void_p),
        ("numModels", C.c_int32), ("modelLeafs", C.c_void_p), ("modelBounds", C.c_void_p),
        ("numSurfaces", C.c_int32), ("surf2patch", C.c_void_p),
        ("numPatches", C.c_int32), ("patches", C.c_void_p), ("patchBounds", C.c_void_p),
        ("numPatchPlanes", C.c_int32), ("patchPlanes", C.c_void_p), ("patchPlaneSignbits", C.c_void_p),
        ("numFacets", C.c_int32), ("facets", C.c_void_p),
        ("numBorders", C.c_int32), ("borders", C.c_void_p),
        ("leafClusterArea", C.c_void_p),
        ("boxBrush", C.c_int32), ("boxFirstPlane", C.c_int32), ("boxFirstSide", C.c_int32), ("boxLeafBrush", C.c_int32),
        ("numClusters", C.c_int32), ("numAreas", C.c_int32), ("vised", C.c_int32), ("clusterBytes", C.c_int32),
        ("visibility", C.c_void_p),
    ]


class _Counters(C.Structure):
    _fields_ = [(k, C.c_ulonglong) for k in ("traces", "nodes", "leafs", "brushRefs", "brushBounds", "brushSides",
                                             "crossings", "splits", "patches", "patchPlanes", "facets")]


# flat-dict key -> (struct count field, struct pointer field, dtype, columns)
_ARRAYS = [
    ("planes", "numPlanes", "planes", np.float32, 4),
    ("plane_type", "numPlanes", "planeType", np.uint8, 0),
    ("plane_signbits", "numPlanes", "planeSignbits", np.uint8, 0),
    ("nodes", "numNodes", "nodes", np.int32, 3),
    ("leafs", "numLeafs", "leafs", np.int32, 4),
    ("leaf_cluster_area", "numLeafs", "leafClusterArea", np.int32, 2),
    ("leafbrushes", "numLeafBrushes", "leafbrushes", np.int32, 0),
    ("leafsurfaces", "numLeafSurfaces", "leafsurfaces", np.int32, 0),
    ("brushes", "numBrushes", "brushes", np.int32, 3),
    ("brush_bounds", "numBrushes", "brushBounds", np.float32, 6),
    ("brushsides", "numBrushSides", "brushsides", np.int32, 2),
    ("model_leafs", "numModels", "modelLeafs", np.int32, 4),
    ("model_bounds", "numModels", "modelBounds", np.float32, 6),
    ("surf2patch", "numSurfaces", "surf2patch", np.int32, 0),
    ("patches", "numPatches", "patches", np.int32, 6),
    ("patch_bounds", "numPatches", "patchBounds", np.float32, 6),
    ("patch_planes", "numPatchPlanes", "patchPlanes", np.float32, 4),
    ("patch_plane_signbits", "numPatchPlanes", "patchPlaneSignbits", np.int32, 0),
    ("facets", "numFacets", "facets", np.int32, 3),
    ("borders", "numBorders", "borders", np.int32, 2),
]
_SCALARS = [("box_brush", "boxBrush"), ("box_first_plane", "boxFirstPlane"), ("box_first_side", "boxFirstSide"),
            ("box_leaf_brush", "boxLeafBrush")]

_lib = None


# cmgpu_sv_entity (include/cm_gpu.h): one linked entity as SV_ClipMoveToEntities reads it
SV_ENTITY_DTYPE = np.dtype([("number", "<i4"), ("ownerNum", "<i4"), ("contents", "<i4"), ("bmodel", "<i4"), ("modelindex", "<i4"),
                            ("mins", "<f4", (3,)), ("maxs", "<f4", (3,)), ("absmin", "<f4", (3,)), ("absmax", "<f4", (3,)),
                            ("origin", "<f4", (3,)), ("angles", "<f4", (3,))])
assert SV_ENTITY_DTYPE.itemsize == 92
# cmgpu_entity_clusters_t: what the PVS part of SV_LinkEntity stores (server.h:42-45)
# aas_trace_t (botlib/be_aas.h:84-93) == cmgpu_aas_trace_t
AAS_TRACE_DTYPE = np.dtype([("startsolid", "<i4"), ("fraction", "<f4"), ("endpos", "<f4", (3,)), ("ent", "<i4"),
                            ("lastarea", "<i4"), ("area", "<i4"), ("planenum", "<i4")])
# aas_clientmove_t (botlib/be_aas.h:178-189) == cmgpu_aas_clientmove_t
AAS_MOVE_DTYPE = np.dtype([("endpos", "<f4", (3,)), ("endarea", "<i4"), ("velocity", "<f4", (3,)), ("trace", AAS_TRACE_DTYPE),
                           ("presencetype", "<i4"), ("stopevent", "<i4"), ("endcontents", "<i4"), ("time", "<f4"), ("frames", "<i4")])


class _AasMoves(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("origins", "velocities", "cmdmoves", "frametimes", "presencetypes", "ongrounds", "cmdframes",
                                          "maxframes", "stopevents", "stopareanums")]


ENT_CLUSTERS_DTYPE = np.dtype([("numLeafs", "<i4"), ("areanum", "<i4"), ("areanum2", "<i4"), ("numClusters", "<i4"),
                               ("lastCluster", "<i4"), ("clusternums", "<i4", (16,))])
assert ENT_CLUSTERS_DTYPE.itemsize == 84


def load_library() -> C.CDLL:
    """Load libcmgpu.so; loud failure when it has not been built (no fallback exists)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("CMGPU_LIB", LIB_PATH)      # experiments may point at another build of the same library
    if not os.path.exists(path):
        raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(make -C threecore_b200/csrc). The engine has no CPU path.")
    lib = C.CDLL(path)
    lib.cmgpu_last_error.restype = C.c_char_p
    lib.cmgpu_last_error.argtypes = [C.c_void_p]
    lib.cmgpu_create.argtypes = [C.POINTER(C.c_void_p), C.c_int]
    lib.cmgpu_destroy.argtypes = [C.c_void_p]
    lib.cmgpu_destroy.restype = None
    lib.cmgpu_flatmap_from_bsp.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.POINTER(_FlatMap)), C.c_char_p, C.c_size_t]
    lib.cmgpu_flatmap_from_bsp_device.argtypes = [C.c_int, C.c_char_p, C.c_size_t, C.POINTER(C.POINTER(_FlatMap)), C.c_char_p, C.c_size_t]
    lib.cmgpu_flatmap_free.argtypes = [C.POINTER(_FlatMap)]
    lib.cmgpu_flatmap_free.restype = None
    lib.cmgpu_upload.argtypes = [C.c_void_p, C.POINTER(_FlatMap)]
    lib.cmgpu_clear.argtypes = [C.c_void_p]
    lib.cmgpu_set_cvars.argtypes = [C.c_void_p, C.c_int, C.c_int]
    lib.cmgpu_num_inline_models.argtypes = [C.c_void_p]
    lib.cmgpu_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
    lib.cmgpu_launch_count.argtypes = [C.c_void_p]
    lib.cmgpu_launch_count.restype = C.c_ulonglong
    lib.cmgpu_ipc_alloc.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p), C.c_char_p]
    lib.cmgpu_ipc_open.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_void_p)]
    lib.cmgpu_ipc_close.argtypes = [C.c_void_p, C.c_void_p]
    lib.cmgpu_ipc_free.argtypes = [C.c_void_p, C.c_void_p]
    lib.cmgpu_host_register.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    lib.cmgpu_aas_upload.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.cmgpu_aas_predict_client_movement_batch.argtypes = [C.c_void_p, C.c_int, C.POINTER(_AasMoves), C.c_void_p, C.c_void_p]
    lib.cmgpu_aas_predict_client_movement_batch_device.argtypes = [C.c_void_p, C.c_int, C.POINTER(_AasMoves), C.c_void_p, C.c_void_p, C.c_void_p]
    lib.cmgpu_aas_set_settings.argtypes = [C.c_void_p, C.c_void_p]
    lib.cmgpu_aas_default_settings.argtypes = [C.c_void_p]
    lib.cmgpu_aas_default_settings.restype = None
    lib.cmgpu_aas_trace_client_bbox_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    lib.cmgpu_aas_trace_client_bbox_batch_device.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.cmgpu_aas_point_area_num_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    lib.cmgpu_aas_point_area_num_batch_device.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.cmgpu_host_unregister.argtypes = [C.c_void_p, C.c_void_p]
    lib.cmgpu_last_batch_split.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
    lib.cmgpu_format_records.argtypes = [C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3 + [C.c_int] + [C.c_void_p] * 3
    lib.cmgpu_format_records_threads.argtypes = [C.c_int] + lib.cmgpu_format_records.argtypes
    lib.cmgpu_kernel_timing.argtypes = [C.c_void_p, C.c_int]
    lib.cmgpu_kernel_time_ms.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int)]
    lib.cmgpu_enable_counters.argtypes = [C.c_void_p, C.c_int]
    lib.cmgpu_read_counters.argtypes = [C.c_void_p, C.POINTER(_Counters), C.c_int]
    lib.cmgpu_check_status.argtypes = [C.c_void_p, C.c_void_p]
    vp = C.c_void_p
    lib.cmgpu_box_trace_batch.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.c_int, vp, vp, C.c_int, vp, vp, vp, vp]
    lib.cmgpu_transformed_box_trace_batch.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.c_int, vp, vp, C.c_int,
                                                      vp, vp, vp, vp, vp, vp]
    lib.cmgpu_point_contents_batch.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp, vp]
    lib.cmgpu_box_trace_batch_device.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.c_int, vp, vp, C.c_int,
                                                 vp, vp, vp, vp, vp, vp, vp, vp]
    lib.cmgpu_point_contents_batch_device.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.cmgpu_rotation_matrices.argtypes = [C.c_int, vp, vp, vp]
    lib.cmgpu_sv_set_entities.argtypes = [vp, C.c_int, vp, vp, C.c_int]
    lib.cmgpu_sv_trace_batch.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.c_int, vp, C.c_int, vp, C.c_int, vp]
    lib.cmgpu_sv_trace_batch_device.argtypes = [vp, C.c_int, vp, vp, vp, vp, C.c_int, vp, C.c_int, vp, C.c_int, vp, vp]
    lib.cmgpu_sv_point_contents_batch.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp]
    lib.cmgpu_sv_point_contents_batch_device.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp, vp]
    lib.cmgpu_sv_last_candidates.argtypes = [vp]
    lib.cmgpu_sv_last_candidates.restype = C.c_ulonglong
    lib.cmgpu_rotation_matrices.restype = None
    lib.cmgpu_measure_l2_read.argtypes = [vp, C.c_size_t, C.c_int, C.POINTER(C.c_double)]
    lib.cmgpu_point_leafnum_batch.argtypes = [vp, C.c_int, vp, vp]
    lib.cmgpu_box_leafnums_batch.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp, vp, vp]
    lib.cmgpu_point_leafnum_batch_device.argtypes = [vp, C.c_int, vp, vp, vp]
    lib.cmgpu_box_leafnums_batch_device.argtypes = [vp, C.c_int, vp, vp, C.c_int, vp, vp, vp, vp]
    lib.cmgpu_entity_clusters_batch.argtypes = [vp, C.c_int, vp, vp, vp]
    lib.cmgpu_entity_clusters_batch_device.argtypes = [vp, C.c_int, vp, vp, vp, vp]
    lib.cmgpu_adjust_area_portal_state.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    lib.cmgpu_areas_connected.argtypes = [vp, C.c_int, C.c_int]
    lib.cmgpu_write_area_bits.argtypes = [vp, vp, C.c_int]
    lib.cmgpu_set_no_areas.argtypes = [vp, C.c_int]
    lib.cmgpu_visible_from_points_batch.argtypes = [vp, C.c_int, vp, C.c_int, vp, vp, vp]
    lib.cmgpu_visible_from_points_batch_device.argtypes = [vp, C.c_int, vp, C.c_int, vp, vp, vp, vp]
    if lib.cmgpu_abi_version() != 2:
        raise ImportError("libcmgpu.so ABI version mismatch")
    _lib = lib
    return lib


class FlatMap:
    """A cmgpu_flatmap: either produced by the library's own BSP loader or assembled from arrays."""

    def __init__(self):
        self._ptr = None          # POINTER(_FlatMap) owned by the library
        self._struct = None       # _FlatMap assembled here
        self._keep = []

    @classmethod
    def from_bsp(cls, data: bytes, tess_device: int | None = None) -> "FlatMap":
        """tess_device: CUDA device that tessellates the Bezier patches (cmgpu_flatmap_from_bsp_device); None: the host does"""
        lib = load_library()
        self = cls()
        out = C.POINTER(_FlatMap)()
        err = C.create_string_buffer(512)
        if tess_device is None:
            rc = lib.cmgpu_flatmap_from_bsp(bytes(data), len(data), C.byref(out), err, len(err))
        else:
            rc = lib.cmgpu_flatmap_from_bsp_device(int(tess_device), bytes(data), len(data), C.byref(out), err, len(err))
        if rc:
            raise CMGPUError(rc, err.value.decode(errors="replace"))
        self._ptr = out
        return self

    @classmethod
    def from_dict(cls, d: dict) -> "FlatMap":
        """From a flat dict (same keys as to_dict); used to upload a map flattened elsewhere,
        e.g. by the engine-side shim out of a loaded clipMap_t."""
        self = cls()
        s = _FlatMap()
        for key, cnt, ptr, dtype, cols in _ARRAYS:
            if key not in d:
                if key == "leaf_cluster_area":
                    continue
                raise KeyError(key)
            a = np.ascontiguousarray(d[key], dtype=dtype)
            n = a.shape[0]
            if a.size == 0:
                a = np.zeros(8, dtype=dtype)
            self._keep.append(a)
            setattr(s, cnt, n)
            setattr(s, ptr, a.ctypes.data)
        for key, fld in _SCALARS:
            setattr(s, fld, int(d[key]))
        # clusters, areas, visibility matrix (optional; derived like cm_load.c:313-316,503-505 when absent)
        lca = d.get("leaf_cluster_area")
        s.numClusters = int(d.get("num_clusters", (int(np.max(lca[:, 0])) + 1) if lca is not None and len(lca) else 0))
        s.numAreas = int(d.get("num_areas", (int(np.max(lca[:, 1])) + 1) if lca is not None and len(lca) else 0))
        vis = d.get("visibility")
        if vis is not None and np.size(vis):
            v = np.ascontiguousarray(vis, dtype=np.uint8)
            self._keep.append(v)
            s.vised = 1
            s.clusterBytes = int(d["cluster_bytes"])
            s.visibility = v.ctypes.data
        else:
            s.vised = 0
            s.clusterBytes = (s.numClusters + 31) & ~31
        self._struct = s
        return self

    @property
    def struct(self) -> _FlatMap:
        return self._ptr.contents if self._ptr is not None else self._struct

    def pointer(self):
        return self._ptr if self._ptr is not None else C.pointer(self._struct)

    def to_dict(self) -> dict:
        s = self.struct
        d = {}
        for key, cnt, ptr, dtype, cols in _ARRAYS:
            n = getattr(s, cnt)
            addr = getattr(s, ptr)
            shape = (n, cols) if cols else (n,)
            if not addr or n == 0:
                d[key] = np.zeros(shape, dtype=dtype)
                continue
            nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
            buf = (C.c_char * nbytes).from_address(addr)
            d[key] = np.frombuffer(buf, dtype=dtype).reshape(shape).copy()
        for key, fld in _SCALARS:
            d[key] = int(getattr(s, fld))
        d["num_clusters"], d["num_areas"], d["cluster_bytes"] = int(s.numClusters), int(s.numAreas), int(s.clusterBytes)
        if s.vised and s.visibility:
            nbytes = int(s.numClusters) * int(s.clusterBytes)
            d["visibility"] = np.frombuffer((C.c_char * nbytes).from_address(s.visibility), dtype=np.uint8).copy()
        else:
            d["visibility"] = np.zeros(0, np.uint8)
        return d

    def __del__(self):
        if self._ptr is not None and _lib is not None:
            _lib.cmgpu_flatmap_free(self._ptr)
            self._ptr = None


def _f(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float32)


def format_records(results16, starts, ends, side_planes, side_type_signbits, side_surface_flags, brush_contents, want_hit=False,
                   threads=0):
    """cmgpu_format_records: the host threads' record formatter on its own (no device).  results16: uint32 [n][4].
    threads > 0: through the worker pool (cmgpu_format_records_threads)."""
    lib = load_library()
    r16 = np.ascontiguousarray(results16, dtype=np.uint32).reshape(-1, 4)
    n = r16.shape[0]
    starts, ends = _f(starts), _f(ends)
    sp, st, sf, bc = _f(side_planes), _i(side_type_signbits), _i(side_surface_flags), _i(brush_contents)
    out = np.zeros(n, dtype=TRACE_DTYPE)
    hit = np.zeros(n, dtype=np.int32) if want_hit else None
    args = (n, _p(r16), _p(starts), _p(ends), len(st), _p(sp), _p(st), _p(sf), len(bc), _p(bc), _p(out), _p(hit))
    rc = lib.cmgpu_format_records_threads(int(threads), *args) if threads else lib.cmgpu_format_records(*args)
    if rc:
        raise CMGPUError(rc, "cmgpu_format_records: bad argument or index out of range")
    return (out, hit) if want_hit else out


def _i(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.int32)


def _p(a):
    return None if a is None else a.ctypes.data


def rotation_matrices(angles):
    """Host AngleVectors + CreateRotationMatrix (q_math.c:999, cm_trace.c:65): ([n,9] float32, [n] int32)."""
    lib = load_library()
    a = _f(angles).reshape(-1, 3)
    m = np.zeros((a.shape[0], 9), np.float32)
    idx = np.zeros(a.shape[0], np.int32)
    lib.cmgpu_rotation_matrices(a.shape[0], _p(a), _p(m), _p(idx))
    return m, idx


class Engine:
    """One cmgpu context on one GPU."""

    def __init__(self, device: int = 0):
        self.lib = load_library()
        h = C.c_void_p()
        rc = self.lib.cmgpu_create(C.byref(h), int(device))
        if rc:
            raise CMGPUError(rc, self.lib.cmgpu_last_error(None).decode(errors="replace"))
        self.h = h
        self.device = int(device)
        self.map = None

    def close(self):
        if getattr(self, "h", None):
            self.lib.cmgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc:
            raise CMGPUError(rc, self.lib.cmgpu_last_error(self.h).decode(errors="replace"))

    # ---- map ---------------------------------------------------------------
    def load_bsp(self, data: bytes, gpu_patches: bool = False) -> FlatMap:
        """parse the .bsp image (CM_LoadMap) and upload it; gpu_patches: tessellate the Bezier patches on this GPU"""
        fm = FlatMap.from_bsp(data, self.device if gpu_patches else None)
        self.upload(fm)
        return fm

    def upload(self, fm: FlatMap):
        self._check(self.lib.cmgpu_upload(self.h, fm.pointer()))
        self.map = fm

    def clear(self):
        self._check(self.lib.cmgpu_clear(self.h))
        self.map = None

    def set_cvars(self, noCurves=0, playerCurveClip=1):
        self._check(self.lib.cmgpu_set_cvars(self.h, int(noCurves), int(playerCurveClip)))

    def num_inline_models(self):
        return self.lib.cmgpu_num_inline_models(self.h)

    def set_option(self, name: str, value: float):
        """Tuning knobs of cmgpu_set_option (cm_gpu.h); none changes results."""
        self._check(self.lib.cmgpu_set_option(self.h, name.encode(), float(value)))

    @property
    def launch_count(self) -> int:
        return int(self.lib.cmgpu_launch_count(self.h))

    # ---- result buffers shared between the ranks of one node (cmgpu_ipc_*) ----------------
    def ipc_alloc(self, nbytes: int):
        """-> (device pointer, 64-byte handle) of a buffer other processes of this node can open"""
        p = C.c_void_p()
        h = C.create_string_buffer(64)
        self._check(self.lib.cmgpu_ipc_alloc(self.h, int(nbytes), C.byref(p), h))
        return int(p.value), h.raw

    def ipc_open(self, handle: bytes) -> int:
        p = C.c_void_p()
        self._check(self.lib.cmgpu_ipc_open(self.h, handle, C.byref(p)))
        return int(p.value)

    def ipc_close(self, ptr: int):
        self._check(self.lib.cmgpu_ipc_close(self.h, C.c_void_p(ptr)))

    def ipc_free(self, ptr: int):
        self._check(self.lib.cmgpu_ipc_free(self.h, C.c_void_p(ptr)))

    def host_register(self, ptr: int, nbytes: int):
        """pin a long-lived host region (cmgpu_host_register) so that the host-pointer calls copy straight from / to it"""
        self._check(self.lib.cmgpu_host_register(self.h, C.c_void_p(ptr), int(nbytes)))

    def host_unregister(self, ptr: int):
        self._check(self.lib.cmgpu_host_unregister(self.h, C.c_void_p(ptr)))

    def last_batch_split(self):
        """(records written by the host threads, records written by the device) of the last big host batch (cmgpu_last_batch_split)"""
        a, b = C.c_ulonglong(0), C.c_ulonglong(0)
        self._check(self.lib.cmgpu_last_batch_split(self.h, C.byref(a), C.byref(b)))
        return int(a.value), int(b.value)

    def kernel_timing(self, on=True):
        """bracket every trace-kernel launch with CUDA events on its own stream (cmgpu_kernel_timing)"""
        self._check(self.lib.cmgpu_kernel_timing(self.h, int(bool(on))))

    def kernel_time_ms(self):
        """(sum of the bracketed trace-kernel durations in ms, number of launches) since the last call"""
        ms, n = C.c_double(), C.c_int()
        self._check(self.lib.cmgpu_kernel_time_ms(self.h, C.byref(ms), C.byref(n)))
        return float(ms.value), int(n.value)

    def measure_l2_read(self, nbytes=24 << 20, passes=50) -> float:
        """GB/s at which the SMs read an L2-resident buffer (denominator of bench.py's roofline.l2)"""
        g = C.c_double()
        self._check(self.lib.cmgpu_measure_l2_read(self.h, int(nbytes), int(passes), C.byref(g)))
        return float(g.value)

    def enable_counters(self, on=True):
        self._check(self.lib.cmgpu_enable_counters(self.h, int(bool(on))))

    def read_counters(self, reset=True) -> dict:
        c = _Counters()
        self._check(self.lib.cmgpu_read_counters(self.h, C.byref(c), int(reset)))
        return {k: int(getattr(c, k)) for k, _ in _Counters._fields_}

    # ---- host-pointer calls (numpy) -------------------------------------------
    @staticmethod
    def _hull(n, mins, maxs):
        if mins is None:
            return None, None, 0
        mins, maxs = _f(mins), _f(maxs)
        if mins.ndim == 1:
            return mins, maxs, 0
        assert mins.shape == (n, 3) and maxs.shape == (n, 3)
        return mins, maxs, 1

    @staticmethod
    def _mask(n, mask):
        m = np.asarray(mask)
        if m.ndim == 0:
            return np.array([int(m)], dtype=np.int64).astype(np.int32), 0
        assert m.shape == (n,)
        return _i(m), 1

    def box_trace(self, starts, ends, mins=None, maxs=None, model=0, mask=1, boxmins=None, boxmaxs=None,
                  want_hit=False, out=None):
        """cmgpu_box_trace_batch on host arrays: n x CM_BoxTrace."""
        starts, ends = _f(starts), _f(ends)
        n = starts.shape[0]
        mins, maxs, hs = self._hull(n, mins, maxs)
        m, ms = self._mask(n, mask)
        models = None
        if not np.isscalar(model) or model != 0:
            models = _i(np.broadcast_to(_i(model), (n,)))
        boxmins, boxmaxs = _f(boxmins), _f(boxmaxs)
        if out is None:
            out = np.zeros(n, dtype=TRACE_DTYPE)
        hit = np.zeros(n, np.int32) if want_hit else None
        self._check(self.lib.cmgpu_box_trace_batch(self.h, n, _p(starts), _p(ends), _p(mins), _p(maxs), hs,
                                                   _p(models), _p(m), ms, _p(boxmins), _p(boxmaxs), _p(out), _p(hit)))
        return (out, hit) if want_hit else out

    def transformed_box_trace(self, starts, ends, mins, maxs, models, mask, origins, angles,
                              boxmins=None, boxmaxs=None, want_hit=False):
        starts, ends = _f(starts), _f(ends)
        n = starts.shape[0]
        mins, maxs, hs = self._hull(n, mins, maxs)
        m, ms = self._mask(n, mask)
        models = _i(np.broadcast_to(_i(models), (n,)))
        origins, angles = _f(origins), _f(angles)
        boxmins, boxmaxs = _f(boxmins), _f(boxmaxs)
        out = np.zeros(n, dtype=TRACE_DTYPE)
        hit = np.zeros(n, np.int32) if want_hit else None
        self._check(self.lib.cmgpu_transformed_box_trace_batch(
            self.h, n, _p(starts), _p(ends), _p(mins), _p(maxs), hs, _p(models), _p(m), ms,
            _p(origins), _p(angles), _p(boxmins), _p(boxmaxs), _p(out), _p(hit)))
        return (out, hit) if want_hit else out

    def point_contents(self, pts, models=None, origins=None, angles=None, boxmins=None, boxmaxs=None):
        pts = _f(pts)
        n = pts.shape[0]
        if models is not None:
            models = _i(np.broadcast_to(_i(models), (n,)))
        out = np.zeros(n, np.int32)
        self._check(self.lib.cmgpu_point_contents_batch(self.h, n, _p(pts), _p(models), _p(_f(origins)), _p(_f(angles)),
                                                        _p(_f(boxmins)), _p(_f(boxmaxs)), _p(out)))
        return out

    def box_trace_hostptr(self, n, starts_ptr, ends_ptr, mins, maxs, mask, out_ptr):
        """Raw host-pointer call (pinned torch tensors): shared hull and mask, world model."""
        mins, maxs = _f(mins), _f(maxs)
        m = np.array([int(mask)], dtype=np.int64).astype(np.int32)
        self._check(self.lib.cmgpu_box_trace_batch(self.h, int(n), starts_ptr, ends_ptr, _p(mins), _p(maxs), 0,
                                                   None, _p(m), 0, None, None, out_ptr, None))

    # ---- device-pointer calls (torch tensors on this GPU) -----------------------
    def box_trace_device(self, starts, ends, mins, maxs, mask, out, models=None, origins=None, rotations=None,
                         rot_index=None, boxmins=None, boxmaxs=None, hit_ids=None, stream=None):
        """cmgpu_box_trace_batch_device: per-item arguments are CUDA tensors; only enqueues work.

        starts/ends float32 [n,3] CUDA tensors.  mins/maxs: a shared hull as 3 host floats (numpy / list /
        None) or per-item CUDA tensors [n,3].  mask: a shared int or a per-item int32 CUDA tensor [n].
        out uint8 [n,56] (a CUDA tensor, or a raw device pointer such as PeerResults.pointer()).  Uses torch's current stream unless `stream` (a cudaStream_t int) is given.
        """
        import torch
        n = starts.shape[0]
        if stream is None:
            stream = torch.cuda.current_stream(starts.device).cuda_stream
        ptr = lambda t: None if t is None else (int(t) if isinstance(t, int) else t.data_ptr())
        keep = []
        if mins is None or not isinstance(mins, torch.Tensor):
            hs = 0
            pmins = pmaxs = None
            if mins is not None:
                hm, hx = _f(mins).reshape(3), _f(maxs).reshape(3)
                keep += [hm, hx]
                pmins, pmaxs = hm.ctypes.data, hx.ctypes.data
        else:
            hs = 1
            pmins, pmaxs = mins.data_ptr(), maxs.data_ptr()
        if isinstance(mask, torch.Tensor):
            ms, pmask = 1, mask.data_ptr()
        else:
            hmask = np.array([int(mask)], dtype=np.int64).astype(np.int32)
            keep.append(hmask)
            ms, pmask = 0, hmask.ctypes.data
        self._check(self.lib.cmgpu_box_trace_batch_device(
            self.h, n, ptr(starts), ptr(ends), pmins, pmaxs, hs, ptr(models), pmask, ms,
            ptr(origins), ptr(rotations), ptr(rot_index), ptr(boxmins), ptr(boxmaxs), ptr(out), ptr(hit_ids),
            C.c_void_p(stream)))

    def point_contents_device(self, points, out, models=None, origins=None, rotations=None, rot_index=None,
                              boxmins=None, boxmaxs=None, stream=None):
        import torch
        n = points.shape[0]
        if stream is None:
            stream = torch.cuda.current_stream(points.device).cuda_stream
        ptr = lambda t: None if t is None else t.data_ptr()
        self._check(self.lib.cmgpu_point_contents_batch_device(
            self.h, n, ptr(points), ptr(models), ptr(origins), ptr(rotations), ptr(rot_index),
            ptr(boxmins), ptr(boxmaxs), ptr(out), C.c_void_p(stream)))

    # ---- botlib's AAS tracer (botlib/be_aas_sample.c:212-258, 399-665) -----------------
    def aas_upload(self, world):
        """world: planes [n,5] float32, nodes [n,3] int32, presence [numareas] int32 (aasgen.AasWorld or the like)"""
        planes, nodes, pres = _f(world.planes), _i(world.nodes), _i(world.presence)
        cont = getattr(world, "contents", None)
        cont = _i(cont) if cont is not None else None
        self._check(self.lib.cmgpu_aas_upload(self.h, planes.shape[0], _p(planes), nodes.shape[0], _p(nodes), pres.shape[0], _p(pres), _p(cont)))

    AAS_MOVE_KEYS = ("origins", "velocities", "cmdmoves", "frametimes", "presencetypes", "ongrounds", "cmdframes", "maxframes",
                     "stopevents", "stopareanums")

    def aas_predict_client_movement(self, p):
        """n x AAS_PredictClientMovement(&move, -1, ...) on host arrays; p: dict with AAS_MOVE_KEYS -> (aas_clientmove_t records, return values)"""
        a = {k: (_f(p[k]) if k in ("origins", "velocities", "cmdmoves", "frametimes") else _i(p[k])) for k in self.AAS_MOVE_KEYS}
        n = a["origins"].shape[0]
        m = _AasMoves(*[a[k].ctypes.data for k in self.AAS_MOVE_KEYS])
        moves = np.zeros(n, dtype=AAS_MOVE_DTYPE)
        res = np.zeros(n, np.int32)
        self._check(self.lib.cmgpu_aas_predict_client_movement_batch(self.h, n, C.byref(m), _p(moves), _p(res)))
        return moves, res

    def aas_predict_client_movement_device(self, p, moves, results=None, stream=None):
        """p: dict of CUDA tensors with AAS_MOVE_KEYS; moves: uint8 [n,84]; results: int32 [n] or None"""
        import torch
        if stream is None:
            stream = torch.cuda.current_stream(moves.device).cuda_stream
        m = _AasMoves(*[p[k].data_ptr() for k in self.AAS_MOVE_KEYS])
        self._check(self.lib.cmgpu_aas_predict_client_movement_batch_device(self.h, p["origins"].shape[0], C.byref(m), moves.data_ptr(),
                                                                           None if results is None else results.data_ptr(), C.c_void_p(stream)))

    def aas_trace_client_bbox(self, starts, ends, presencetype) -> np.ndarray:
        """n x AAS_TraceClientBBox(start, end, presencetype, -1) on host arrays -> aas_trace_t records"""
        starts, ends = _f(starts), _f(ends)
        n = starts.shape[0]
        pt = np.asarray(presencetype)
        pt, ps = (_i(pt.reshape(1)), 0) if pt.ndim == 0 else (_i(pt), 1)
        out = np.zeros(n, dtype=AAS_TRACE_DTYPE)
        self._check(self.lib.cmgpu_aas_trace_client_bbox_batch(self.h, n, _p(starts), _p(ends), _p(pt), ps, _p(out)))
        return out

    def aas_trace_client_bbox_device(self, starts, ends, presencetype, out, stream=None):
        """CUDA tensors; presencetype: an int (shared) or an int32 CUDA tensor [n]; out: uint8 [n,36]"""
        import torch
        if stream is None:
            stream = torch.cuda.current_stream(starts.device).cuda_stream
        if isinstance(presencetype, int):
            pt = np.array([presencetype], np.int32)
            ptr, ps, keep = pt.ctypes.data, 0, pt
        else:
            ptr, ps, keep = presencetype.data_ptr(), 1, presencetype
        self._check(self.lib.cmgpu_aas_trace_client_bbox_batch_device(self.h, starts.shape[0], starts.data_ptr(), ends.data_ptr(), ptr, ps,
                                                                      out.data_ptr(), C.c_void_p(stream)))
        del keep

    def aas_point_area_num(self, points) -> np.ndarray:
        points = _f(points)
        out = np.zeros(points.shape[0], np.int32)
        self._check(self.lib.cmgpu_aas_point_area_num_batch(self.h, points.shape[0], _p(points), _p(out)))
        return out

    # ---- whole SV_Trace on the device (server/sv_world.c:562-609) -----------------
    def sv_set_entities(self, ents: np.ndarray, owner_nums=None):
        """ents: structured array of SV_ENTITY_DTYPE (the linked entities, in SV_AreaEntities order);
        owner_nums: r.ownerNum of every gentity by number (for the pass entity's owner), or None"""
        ents = np.ascontiguousarray(ents, dtype=SV_ENTITY_DTYPE)
        own = None if owner_nums is None else _i(owner_nums)
        self._check(self.lib.cmgpu_sv_set_entities(self.h, len(ents), _p(ents), _p(own), 0 if own is None else len(own)))

    def sv_trace(self, starts, ends, mins, maxs, pass_ents, masks):
        """host arrays in, structured trace_t array out; hull / pass entity / mask shared (1-d, scalar) or per move"""
        starts, ends = _f(starts), _f(ends)
        n = starts.shape[0]
        mins, maxs, hs = self._hull(n, mins, maxs)
        pe = np.asarray(pass_ents)
        pe, ps = (_i(pe.reshape(1)), 0) if pe.ndim == 0 else (_i(pe), 1)
        m, ms = self._mask(n, masks)
        out = np.zeros(n, dtype=TRACE_DTYPE)
        self._check(self.lib.cmgpu_sv_trace_batch(self.h, n, _p(starts), _p(ends), _p(mins), _p(maxs), hs, _p(pe), ps,
                                                  _p(m), ms, _p(out)))
        return out

    def sv_trace_device(self, starts, ends, mins, maxs, pass_ent, mask, out, stream=None):
        """resident moves (CUDA tensors [n,3]); shared hull (3 host floats each), pass entity and mask (ints),
        or per-move CUDA tensors for any of them"""
        import torch
        n = starts.shape[0]
        if stream is None:
            stream = torch.cuda.current_stream(starts.device).cuda_stream
        keep = []

        def shared_or_tensor(v, dtype):
            if isinstance(v, torch.Tensor):
                return v.data_ptr(), 1
            if v is None:
                return None, 0
            h = np.ascontiguousarray(np.asarray(v).reshape(-1), dtype=dtype)
            keep.append(h)
            return h.ctypes.data, 0
        pmins, hs = shared_or_tensor(mins, np.float32)
        pmaxs, _ = shared_or_tensor(maxs, np.float32)
        ppass, ps = shared_or_tensor(np.int64(pass_ent).astype(np.int32) if not isinstance(pass_ent, torch.Tensor) else pass_ent, np.int32)
        pmask, ms = shared_or_tensor(np.int64(mask).astype(np.int32) if not isinstance(mask, torch.Tensor) else mask, np.int32)
        self._check(self.lib.cmgpu_sv_trace_batch_device(self.h, n, starts.data_ptr(), ends.data_ptr(), pmins, pmaxs, hs,
                                                         ppass, ps, pmask, ms, out.data_ptr(), C.c_void_p(stream)))

    def sv_point_contents(self, pts, pass_ents):
        pts = _f(pts)
        pe = np.asarray(pass_ents)
        pe, ps = (_i(pe.reshape(1)), 0) if pe.ndim == 0 else (_i(pe), 1)
        out = np.zeros(pts.shape[0], np.int32)
        self._check(self.lib.cmgpu_sv_point_contents_batch(self.h, pts.shape[0], _p(pts), _p(pe), ps, _p(out)))
        return out

    # ---- leaf queries, entity clusters, visibility (cm_test.c, sv_world.c:255-303, sv_snapshot.c:316-355) ----
    def point_leafnum(self, pts):
        pts = _f(pts).reshape(-1, 3)
        out = np.zeros(pts.shape[0], np.int32)
        self._check(self.lib.cmgpu_point_leafnum_batch(self.h, pts.shape[0], _p(pts), _p(out)))
        return out

    def box_leafnums(self, mins, maxs, listsize=64):
        mins, maxs = _f(mins).reshape(-1, 3), _f(maxs).reshape(-1, 3)
        n = mins.shape[0]
        lists = np.full((n, max(listsize, 0)), -1, np.int32)
        counts = np.zeros(n, np.int32)
        last = np.zeros(n, np.int32)
        self._check(self.lib.cmgpu_box_leafnums_batch(self.h, n, _p(mins), _p(maxs), listsize, _p(lists) if lists.size else None,
                                                      _p(counts), _p(last)))
        return lists, counts, last

    def entity_clusters(self, absmins, absmaxs) -> np.ndarray:
        absmins, absmaxs = _f(absmins).reshape(-1, 3), _f(absmaxs).reshape(-1, 3)
        out = np.zeros(absmins.shape[0], ENT_CLUSTERS_DTYPE)
        self._check(self.lib.cmgpu_entity_clusters_batch(self.h, absmins.shape[0], _p(absmins), _p(absmaxs), _p(out)))
        return out

    def adjust_area_portal_state(self, area1, area2, open_):
        self._check(self.lib.cmgpu_adjust_area_portal_state(self.h, int(area1), int(area2), int(bool(open_))))

    def areas_connected(self, area1, area2) -> bool:
        r = self.lib.cmgpu_areas_connected(self.h, int(area1), int(area2))
        if r < 0:
            self._check(r)
        return bool(r)

    def write_area_bits(self, area, nbytes=32) -> bytes:
        buf = np.zeros(nbytes, np.uint8)
        r = self.lib.cmgpu_write_area_bits(self.h, _p(buf), int(area))
        if r < 0:
            self._check(r)
        return bytes(buf[:r])

    def set_no_areas(self, on):
        self._check(self.lib.cmgpu_set_no_areas(self.h, int(bool(on))))

    def visible_from_points(self, origins, ents: np.ndarray):
        """-> (bool [nViews][nEnts], view leafs) : the PVS stage of SV_AddEntitiesVisibleFromPoint per pair"""
        origins = _f(origins).reshape(-1, 3)
        ents = np.ascontiguousarray(ents, ENT_CLUSTERS_DTYPE)
        nv, ne = origins.shape[0], ents.shape[0]
        words = (ne + 31) // 32
        bits = np.zeros((nv, max(words, 1)), np.uint32)
        leafs = np.zeros(nv, np.int32)
        self._check(self.lib.cmgpu_visible_from_points_batch(self.h, nv, _p(origins), ne, _p(ents) if ne else None,
                                                             _p(bits), _p(leafs)))
        vis = np.unpackbits(bits.view(np.uint8).reshape(nv, -1), axis=1, bitorder="little")[:, :ne].astype(bool)
        return vis, leafs

    # resident variants: CUDA tensors in and out, work only enqueued on the tensors' current stream
    def _stream(self, t, stream):
        import torch
        return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream if stream is None else stream)

    def point_leafnum_device(self, points, out, stream=None):
        self._check(self.lib.cmgpu_point_leafnum_batch_device(self.h, points.shape[0], points.data_ptr(), out.data_ptr(),
                                                              self._stream(points, stream)))

    def box_leafnums_device(self, mins, maxs, listsize, lists, counts, last_leafs, stream=None):
        self._check(self.lib.cmgpu_box_leafnums_batch_device(self.h, mins.shape[0], mins.data_ptr(), maxs.data_ptr(), int(listsize),
                                                             lists.data_ptr(), counts.data_ptr(), last_leafs.data_ptr(),
                                                             self._stream(mins, stream)))

    def entity_clusters_device(self, absmins, absmaxs, out, stream=None):
        """out: uint8/int32 CUDA tensor of n * 84 bytes (cmgpu_entity_clusters_t records)"""
        self._check(self.lib.cmgpu_entity_clusters_batch_device(self.h, absmins.shape[0], absmins.data_ptr(), absmaxs.data_ptr(),
                                                                out.data_ptr(), self._stream(absmins, stream)))

    def visible_from_points_device(self, origins, n_ents, ents, visible, view_leafs=None, stream=None):
        """visible: int32 CUDA tensor [nViews][(n_ents+31)//32]"""
        self._check(self.lib.cmgpu_visible_from_points_batch_device(
            self.h, origins.shape[0], origins.data_ptr(), int(n_ents), ents.data_ptr(), visible.data_ptr(),
            None if view_leafs is None else view_leafs.data_ptr(), self._stream(origins, stream)))

    @property
    def sv_last_candidates(self) -> int:
        return int(self.lib.cmgpu_sv_last_candidates(self.h))

    def check_status(self, stream=None):
        import torch
        if stream is None:
            stream = torch.cuda.current_stream(self.device).cuda_stream
        self._check(self.lib.cmgpu_check_status(self.h, C.c_void_p(stream)))


def traces_from_tensor(t) -> np.ndarray:
    """uint8 [n,56] CUDA/CPU tensor -> numpy structured trace_t array."""
    return t.detach().cpu().numpy().view(TRACE_DTYPE).reshape(-1)
