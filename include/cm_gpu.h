/*
 * cm_gpu.h -- C ABI of the B200 batched collision-trace engine (libcmgpu.so).
 *
 * This is the boundary the ThreeCore host code binds to.  Every entry point is
 * extern "C", takes plain pointers and sizes, and returns an int status
 * (CMGPU_OK == 0).  Nothing here longjmp()s: the engine-side shim turns a
 * non-zero status into Com_Error(ERR_DROP, ...) exactly where the reference
 * raises it (code/qcommon/cm_load.c:770,792-795).
 *
 * Reference interfaces each group replaces (paths relative to the reference
 * tree, code/qcommon unless noted):
 *
 *   cmgpu_flatmap, cmgpu_upload      the upload hook after cm_load.c:734; the
 *                                    arrays are clipMap_t (cm_local.h:120-172)
 *                                    with every pointer turned into an index
 *   cmgpu_flatmap_from_bsp           CM_LoadMap's lump parsing, cm_load.c:606-745
 *   cmgpu_box_trace_batch            CM_BoxTrace / CM_Trace, cm_public.h:45-48
 *                                    (cm_trace.c:545-698), one call per item
 *   cmgpu_transformed_box_trace_batch  CM_TransformedBoxTrace, cm_public.h:49-52
 *                                    (cm_trace.c:708-805) incl. CM_TempBoxModel
 *                                    (cm_load.c:927-949) for handle 1023
 *   cmgpu_point_contents_batch       CM_PointContents / CM_TransformedPointContents,
 *                                    cm_public.h:42-43 (cm_test.c:217-294)
 *   cmgpu_clear                      CM_ClearMap, cm_load.c:753
 *
 * Results are bit-for-bit what the reference writes into trace_t
 * (q_shared.h:976-986), including the stale plane.type/signbits bytes a patch
 * hit leaves behind (cm_patch.c:1309-1310).  entityNum is left 0, as CM does
 * (it is set by the server layer, server/sv_world.c:577).
 *
 * There is no CPU fallback: without a CUDA device every compute entry point
 * returns CMGPU_ERR_NO_DEVICE / CMGPU_ERR_CUDA.
 */
#ifndef CM_GPU_H
#define CM_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#pragma GCC visibility push(default)   /* the library is built with -fvisibility=hidden */

#define CMGPU_ABI_VERSION 2

#define CMGPU_BOX_MODEL_HANDLE 1023   /* BOX_MODEL_HANDLE, cm_local.h:28 */
#define CMGPU_MAX_SUBMODELS    1024   /* MAX_SUBMODELS,   cm_local.h:27 */

enum {
	CMGPU_OK = 0,
	CMGPU_ERR_INVALID = 1,      /* bad argument (NULL pointer, negative count) */
	CMGPU_ERR_NO_DEVICE = 2,    /* no CUDA device / wrong architecture */
	CMGPU_ERR_CUDA = 3,         /* CUDA runtime failure, see cmgpu_last_error */
	CMGPU_ERR_NO_MAP = 4,       /* compute call with no map uploaded (reference: zeroed trace, cm_trace.c:562) */
	CMGPU_ERR_BAD_HANDLE = 5,   /* CM_ClipHandleToModel: bad handle, cm_load.c:768-798 */
	CMGPU_ERR_BAD_MAP = 6,      /* malformed lump / index out of range (the reference's ERR_DROPs in cm_load.c) */
	CMGPU_ERR_NOMEM = 7,
	CMGPU_ERR_LIMIT = 8         /* cmgpu_upload: the BSP is more than 255 levels deep (the traversal stacks are sized by
	                               the depth of the uploaded tree; the reference recurses without a limit) */
};

/* trace_t, q_shared.h:976-986: 56 bytes, layout is ABI (QVM memory, botlib copies). */
typedef struct cmgpu_trace_s {
	int32_t allsolid;
	int32_t startsolid;
	float   fraction;
	float   endpos[3];
	float   plane_normal[3];
	float   plane_dist;
	uint8_t plane_type;
	uint8_t plane_signbits;
	uint8_t plane_pad[2];
	int32_t surfaceFlags;
	int32_t contents;
	int32_t entityNum;
} cmgpu_trace_t;

/*
 * The loaded clip map as index arrays.  Counts INCLUDE the box-hull extras the
 * reference appends in CM_InitBoxHull (cm_load.c:45-50,873-916): the last 12
 * planes, last 6 brush sides, last brush, and leafbrushes[boxLeafBrush].
 * Submodel brush/surface lists (cm_load.c:170-182) live in the same
 * leafbrushes[]/leafsurfaces[] arrays and are addressed by modelLeafs[].
 */
typedef struct cmgpu_flatmap_s {
	int32_t numPlanes;        const float   *planes;         /* [n][4] normal.xyz, dist */
	                          const uint8_t *planeType;      /* cplane_t.type */
	                          const uint8_t *planeSignbits;  /* cplane_t.signbits */
	int32_t numNodes;         const int32_t *nodes;          /* [n][3] plane, children[0], children[1] */
	int32_t numLeafs;         const int32_t *leafs;          /* [n][4] firstLeafBrush, numLeafBrushes, firstLeafSurface, numLeafSurfaces */
	int32_t numLeafBrushes;   const int32_t *leafbrushes;
	int32_t numLeafSurfaces;  const int32_t *leafsurfaces;
	int32_t numBrushes;       const int32_t *brushes;        /* [n][3] contents, firstSide, numSides */
	                          const float   *brushBounds;    /* [n][6] mins, maxs */
	int32_t numBrushSides;    const int32_t *brushsides;     /* [n][2] plane, surfaceFlags */
	int32_t numModels;        const int32_t *modelLeafs;     /* [n][4] as leafs[]; entry 0 (world) unused */
	                          const float   *modelBounds;    /* [n][6] cmodel_t mins, maxs (already spread by 1, cm_load.c:159-163) */
	int32_t numSurfaces;      const int32_t *surf2patch;     /* [n] patch ordinal, -1 where cm.surfaces[i] == NULL */
	int32_t numPatches;       const int32_t *patches;        /* [n][6] surfaceFlags, contents, firstPlane, numPlanes, firstFacet, numFacets */
	                          const float   *patchBounds;    /* [n][6] */
	int32_t numPatchPlanes;   const float   *patchPlanes;    /* [n][4] patchPlane_t.plane, cm_patch.h:26-29 */
	                          const int32_t *patchPlaneSignbits;
	int32_t numFacets;        const int32_t *facets;         /* [n][3] surfacePlane, numBorders, firstBorder */
	int32_t numBorders;       const int32_t *borders;        /* [n][2] plane (patch-local index), borderInward */
	const int32_t *leafClusterArea;                           /* [numLeafs][2] cluster, area (may be NULL) */
	int32_t boxBrush;        /* = numBrushes-1 */
	int32_t boxFirstPlane;   /* = numPlanes-12 */
	int32_t boxFirstSide;    /* = numBrushSides-6 */
	int32_t boxLeafBrush;    /* leafbrushes[] slot holding boxBrush */
	/* visibility and areas (cm_load.c:313-320,495-515) -- only the leaf / PVS queries below read them */
	int32_t numClusters;     /* max leaf cluster + 1, or the vis lump's own count when vised */
	int32_t numAreas;        /* max leaf area + 1 */
	int32_t vised;           /* 0: no vis lump, every cluster sees every cluster (visibility may be NULL) */
	int32_t clusterBytes;    /* bytes per row of visibility[] */
	const uint8_t *visibility;                                /* [numClusters][clusterBytes] bit c of row r: r sees c */
} cmgpu_flatmap;

typedef struct cmgpu_ctx_s cmgpu_ctx;

/* ---- lifetime ------------------------------------------------------------ */
int  cmgpu_abi_version(void);
int  cmgpu_device_count(void);
int  cmgpu_create(cmgpu_ctx **out, int device);
void cmgpu_destroy(cmgpu_ctx *ctx);
const char *cmgpu_last_error(const cmgpu_ctx *ctx);   /* ctx may be NULL: last error of a failed create */

/* ---- map ------------------------------------------------------------------ */
/* Parse an IBSP v46 image the way CM_LoadMap does (brushes, planes, nodes, leafs,
 * inline models, box hull; patches are tessellated on the host as
 * CM_GeneratePatchCollide does).  Free with cmgpu_flatmap_free. */
int  cmgpu_flatmap_from_bsp(const void *bsp, size_t len, cmgpu_flatmap **out, char *err, size_t errlen);
/* The same, with the Bezier patches tessellated ON THE GPU: one thread per MST_PATCH surface runs CM_GeneratePatchCollide
 * (cm_patch.c:1137-1203; the code the host loader runs, csrc/cm_patchgen_core.h) in a fixed scratch block; planes and
 * facets come out in the same order with the same float bits.  `device`: CUDA device ordinal. */
int  cmgpu_flatmap_from_bsp_device(int device, const void *bsp, size_t len, cmgpu_flatmap **out, char *err, size_t errlen);
void cmgpu_flatmap_free(cmgpu_flatmap *map);

/* Validate, pack into the device layout and copy to the context's GPU. Replaces any previous map. */
int  cmgpu_upload(cmgpu_ctx *ctx, const cmgpu_flatmap *map);
int  cmgpu_clear(cmgpu_ctx *ctx);
/* cvars that change results: cm_noCurves (cm_trace.c:159,405), cm_playerCurveClip (cm_patch.c:1234) */
int  cmgpu_set_cvars(cmgpu_ctx *ctx, int noCurves, int playerCurveClip);
int  cmgpu_num_inline_models(const cmgpu_ctx *ctx);
/* Tuning knobs; none of them changes results.  "bin" (0/1: trace large batches in the order of their
 * start cell), "bin_min_items", "bin_cell" (world units; takes effect at the next cmgpu_upload),
 * "bin_long_len", "bin_dir", "gang", "gang_fetch", "light_reps", "fast" (0: always the general kernel), "pool" (0: the
 * one-item-per-lane variant of the hot kernel), "pool_reps", "slab" (0: no float pre-test before the exact brush
 * clip), "two_phase" (0: binned hot launches store 56-byte records by item number from the walk itself instead of
 * 16-byte results by slot plus an expansion pass), "compact_by_item" (0: the pool kernel stores those results by slot and
 * the expansion gathers them), "node_thresholds" (0: the pool kernel's node step subtracts and compares instead of using
 * the per-hull threshold table), "gang_brush", "pipe_chunks" (pieces a host batch is cut into); for big host batches
 * "host_records", "host_threads", "stage_pageable" (described with cmgpu_last_batch_split below), "hybrid_min_items" (2 Mi:
 * smallest batch that shares its records), "hybrid_walk_items" (1 Mi: items per piece), and for measurements "host_share"
 * (>= 0 pins the threads' share), "hybrid_decouple", "hybrid_grid_eighths", "hybrid_in_flight".  Unknown names
 * are CMGPU_ERR_INVALID. */
int  cmgpu_set_option(cmgpu_ctx *ctx, const char *name, double value);

/* ---- batched queries, HOST pointers (copies in, kernel, copies out) -------
 * starts/ends [n][3]; mins/maxs [1][3] when hull_stride==0 or [n][3] when 1,
 * NULL means (0,0,0) as in cm_trace.c:569-574; models NULL = world (0);
 * masks [1] (mask_stride 0) or [n] (1).  boxmins/boxmaxs [n][3] are read only
 * for items whose model is CMGPU_BOX_MODEL_HANDLE (the CM_TempBoxModel box).
 * hit_ids (optional, may be NULL) receives the brush index that produced the
 * result, -2-patch for a patch, -1 for none.                                */
int cmgpu_box_trace_batch(cmgpu_ctx *ctx, int n,
                          const float *starts, const float *ends,
                          const float *mins, const float *maxs, int hull_stride,
                          const int32_t *models, const int32_t *masks, int mask_stride,
                          const float *boxmins, const float *boxmaxs,
                          cmgpu_trace_t *results, int32_t *hit_ids);

int cmgpu_transformed_box_trace_batch(cmgpu_ctx *ctx, int n,
                          const float *starts, const float *ends,
                          const float *mins, const float *maxs, int hull_stride,
                          const int32_t *models, const int32_t *masks, int mask_stride,
                          const float *origins, const float *angles,
                          const float *boxmins, const float *boxmaxs,
                          cmgpu_trace_t *results, int32_t *hit_ids);

/* origins/angles NULL: CM_PointContents; otherwise CM_TransformedPointContents */
int cmgpu_point_contents_batch(cmgpu_ctx *ctx, int n, const float *points,
                               const int32_t *models, const float *origins, const float *angles,
                               const float *boxmins, const float *boxmaxs, int32_t *contents);

/* ---- pinned host memory ----------------------------------------------------
 * The host-pointer calls cut a batch into pieces and overlap copy-in, tracing and copy-out.  That only works from
 * page-locked memory: copies from or to pageable memory are staged by the driver and block the caller, so the pieces
 * run one after the other (measured in bench.py: `e2e_pageable`).  The engine's callers own long-lived arenas -- the
 * game VM's data segment that trap_Trace results are written into (server/sv_game.c:198, VMA(1)), a bot library's
 * result arrays: register such a region ONCE, unregister it before it is freed.  Regions may overlap earlier ones (they
 * are merged); a region somebody else already pinned is accepted as it is.
 *
 * cmgpu_set_option("auto_register", 1) makes the host-pointer calls register any pageable starts / ends / results array
 * of 1 MiB or more on first sight.  Only safe when such arrays live until cmgpu_destroy or are released with
 * cmgpu_host_unregister first: memory that is freed while registered and handed out again by the allocator would be
 * copied through stale page mappings.  Off by default. */
int cmgpu_host_register(cmgpu_ctx *ctx, const void *ptr, size_t bytes);
int cmgpu_host_unregister(cmgpu_ctx *ctx, const void *ptr);

/* ---- who writes the records of a big host batch --------------------------------------------------------------
 * cmgpu_box_trace_batch with one hull and one mask for the whole batch, on a map without active patches, of
 * "hybrid_min_items" (2 Mi) items or more: every piece is traced into 16-byte results (fraction, solid flags, clip side,
 * hit brush); per range of items the 56-byte trace_t records are then EITHER written by the device and copied into the
 * caller's array, OR the 16-byte results are copied into pinned memory of the context and worker threads of the calling
 * process write the records into the caller's array (endpos, plane, surface flags, contents: cm_trace.c:352-356, 671-686,
 * same float operations in the same order -- formatting, not tracing).  The split follows the measured rates of the link
 * and of the threads.  A pageable result array is written by the threads only.
 *   cmgpu_set_option("host_records", v): -1 = as described (default), 0 = off: every record is written by the device
 *       (the round-1 pipeline), 1 = every record by the threads, 2 = the new pipeline with every record by the device,
 *       3 = the new pipeline with the measured share even where the plain one has measured faster.
 *   cmgpu_set_option("host_threads", t): worker threads; -1 (default) = host cores / visible GPUs - 1.
 *   cmgpu_set_option("stage_pageable", 0): pageable start / end arrays of such a batch are uploaded as they are (staged by
 *       the driver, blocking) instead of being copied into pinned memory of the context by the threads, a piece ahead.
 * cmgpu_last_batch_split: how many records of the last such call each side wrote. */
int cmgpu_last_batch_split(const cmgpu_ctx *ctx, unsigned long long *by_host, unsigned long long *by_device);

/* The host threads' record formatter on its own (no device, no context): n results of 16 bytes -- {fraction bits,
 * allsolid | startsolid << 1, clip side or -1, hit brush or -1} -- with the sweeps they belong to and the three tables a
 * record draws on (per brush side: plane normal + dist, plane type | signbits << 8, surfaceFlags; per brush: contents)
 * become n trace_t records (cm_trace.c:352-356, 671-686).  Single-threaded; CMGPU_ERR_INVALID for an index out of range. */
int cmgpu_format_records(int n, const uint32_t *results16, const float *starts, const float *ends,
                         int num_sides, const float *side_planes, const int32_t *side_type_signbits, const int32_t *side_surface_flags,
                         int num_brushes, const int32_t *brush_contents, cmgpu_trace_t *records, int32_t *hit_ids);
/* The same through the worker pool of a big host batch, driven the way cmgpu_box_trace_batch drives it (sweeps staged a piece
 * ahead, one held job per piece, released out of order): what the CPU tests use to exercise the pool without a device. */
int cmgpu_format_records_threads(int threads, int n, const uint32_t *results16, const float *starts, const float *ends,
                                 int num_sides, const float *side_planes, const int32_t *side_type_signbits, const int32_t *side_surface_flags,
                                 int num_brushes, const int32_t *brush_contents, cmgpu_trace_t *records, int32_t *hit_ids);

/* ---- batched queries, DEVICE pointers (inputs resident in HBM) -----------
 * Same meaning as above; per-item arrays are device pointers on the context's
 * GPU.  A SHARED hull (hull_stride == 0: mins/maxs, 3 floats each, may be NULL)
 * and a SHARED mask (mask_stride == 0: 1 int) are HOST pointers read at call
 * time, like the by-value arguments of CM_BoxTrace.  `stream` is a cudaStream_t
 * (NULL = default stream).  The call only enqueues work.  rotations: [n][9] row-major rotation matrices
 * (CreateRotationMatrix, cm_trace.c:65-68) with rot_index[i] == -1 for an
 * unrotated item; build them with cmgpu_rotation_matrices on the host.
 *
 * Model handles in a device array cannot be checked before the launch.  The kernels check them: an item whose handle
 * is neither a loaded inline model nor 1023 touches no map data, is traced against nothing (fraction 1, contents 0),
 * and the next cmgpu_check_status returns CMGPU_ERR_BAD_HANDLE (the reference drops the frame, cm_load.c:768-798).
 *
 * CUDA graphs.  The _device calls may be issued on a stream that is being captured: they then allocate nothing (a
 * workspace that would have to grow is CMGPU_ERR_INVALID: run the same call once before capturing it) and record no
 * events.  The captured launches name the context's workspaces by address; from the first capture on, a workspace
 * that later has to grow (or a hull table that has to make room) is retired, not freed, so replays stay valid until
 * the map is replaced (cmgpu_upload / cmgpu_clear) or the context destroyed -- a graph must not be replayed after
 * that.  Replays and other calls on the same context share those workspaces: keep them in stream order.
 * (Not capturable: cmgpu_sv_trace_batch_device / cmgpu_sv_point_contents_batch_device, which read a count back
 * between their passes.)
 */
int cmgpu_box_trace_batch_device(cmgpu_ctx *ctx, int n,
                          const float *starts, const float *ends,
                          const float *mins, const float *maxs, int hull_stride,
                          const int32_t *models, const int32_t *masks, int mask_stride,
                          const float *origins, const float *rotations, const int32_t *rot_index,
                          const float *boxmins, const float *boxmaxs,
                          cmgpu_trace_t *results, int32_t *hit_ids, void *stream);

int cmgpu_point_contents_batch_device(cmgpu_ctx *ctx, int n, const float *points,
                               const int32_t *models, const float *origins,
                               const float *rotations, const int32_t *rot_index,
                               const float *boxmins, const float *boxmaxs,
                               int32_t *contents, void *stream);

/* AngleVectors (q_math.c:999-1032) + CreateRotationMatrix on the host, same libm.
 * matrices [n][9]; rot_index[i] = i if any angle is non-zero, else -1. */
void cmgpu_rotation_matrices(int n, const float *angles, float *matrices, int32_t *rot_index);

/* Device status word of the launches since the last check (cleared by the call): CMGPU_OK, CMGPU_ERR_BAD_HANDLE if a
 * device-side model handle was out of range, CMGPU_ERR_LIMIT if a traversal stack overflowed (cannot happen on a map
 * that passed cmgpu_upload; kept as a tripwire).  Synchronises `stream`. */
int cmgpu_check_status(cmgpu_ctx *ctx, void *stream);

/* ---- whole SV_Trace / SV_PointContents on the device (server/sv_world.c:330-647) -----
 * The world pass, the broad phase over the linked entities (SV_AreaEntities, :330-394), the pass / owner /
 * contents filters (:496-516), one CM_TransformedBoxTrace per surviving entity (:518-528) and the merge
 * (:532-549) all run on the GPU; results carry entityNum as SV_Trace sets it (:577, :534-546).
 *
 * cmgpu_sv_set_entities snapshots the LINKED entities (call it after the frame's SV_LinkEntity calls).  `ents`
 * should be in the order SV_AreaEntities returns for a box that holds the whole world: ties between entities a
 * move hits at the same fraction go to the first one in that order.  (Entities are re-bucketed by sector node
 * here, so only the relative order of entities of the same node is taken from the caller.)  ownerNums[g] is
 * r.ownerNum of gentity g for g < numGentities (the pass entity's owner, :484-488, need not be linked);
 * it may be NULL when no move passes an entity.  A bmodel entity whose modelindex is not an inline model of
 * the uploaded map is CMGPU_ERR_BAD_HANDLE (CM_InlineModel's ERR_DROP, cm_load.c:792-795).                   */
typedef struct cmgpu_sv_entity_s {
	int32_t number;                  /* touch->s.number */
	int32_t ownerNum;                /* touch->r.ownerNum */
	int32_t contents;                /* touch->r.contents */
	int32_t bmodel;                  /* touch->r.bmodel: inline model, else the box r.mins/r.maxs (sv_world.c:34-42) */
	int32_t modelindex;              /* touch->s.modelindex (bmodel only) */
	float   mins[3], maxs[3];        /* touch->r.mins / r.maxs */
	float   absmin[3], absmax[3];    /* touch->r.absmin / r.absmax as SV_LinkEntity left them (:243-264) */
	float   origin[3], angles[3];    /* touch->r.currentOrigin / currentAngles */
} cmgpu_sv_entity;

int cmgpu_sv_set_entities(cmgpu_ctx *ctx, int n, const cmgpu_sv_entity *ents, const int32_t *ownerNums, int numGentities);

/* n calls of SV_Trace.  HOST pointers; hull_stride / pass_stride / mask_stride 0 = one shared value. */
int cmgpu_sv_trace_batch(cmgpu_ctx *ctx, int n, const float *starts, const float *ends,
                         const float *mins, const float *maxs, int hull_stride,
                         const int32_t *passEntityNums, int pass_stride,
                         const int32_t *contentmasks, int mask_stride, cmgpu_trace_t *results);
/* DEVICE pointers for the per-move arrays; shared values (stride 0) are host pointers read at call time.
 * Enqueues the world pass and the broad phase on `stream`, waits for the candidate count (one 4-byte read),
 * then enqueues the entity pass and the merge. */
int cmgpu_sv_trace_batch_device(cmgpu_ctx *ctx, int n, const float *starts, const float *ends,
                         const float *mins, const float *maxs, int hull_stride,
                         const int32_t *passEntityNums, int pass_stride,
                         const int32_t *contentmasks, int mask_stride, cmgpu_trace_t *results, void *stream);
/* n calls of SV_PointContents (:616-647): world contents | CM_TransformedPointContents of every linked entity
 * whose box holds the point, except passEntityNum. */
int cmgpu_sv_point_contents_batch(cmgpu_ctx *ctx, int n, const float *points,
                         const int32_t *passEntityNums, int pass_stride, int32_t *contents);
int cmgpu_sv_point_contents_batch_device(cmgpu_ctx *ctx, int n, const float *points,
                         const int32_t *passEntityNums, int pass_stride, int32_t *contents, void *stream);
/* candidates (entity clips) the last cmgpu_sv_* call generated */
unsigned long long cmgpu_sv_last_candidates(const cmgpu_ctx *ctx);

/* ---- leaf queries (cm_test.c:31-61, 131-181): the tree lookups behind entity linking and PVS culling --------
 * cmgpu_point_leafnum_batch: CM_PointLeafnum per point.  cmgpu_box_leafnums_batch: CM_BoxLeafnums per box --
 * lists[i*listsize ..] receives the leafs the box touches in the reference's order (children[0] first), counts[i]
 * how many were stored (<= listsize), lastLeafs[i] the last touched leaf that has a cluster, counted even when
 * the list is full (cm_test.c:78-81).  The _device variants take device pointers and only enqueue work.     */
int cmgpu_point_leafnum_batch(cmgpu_ctx *ctx, int n, const float *points, int32_t *leafnums);
int cmgpu_box_leafnums_batch(cmgpu_ctx *ctx, int n, const float *mins, const float *maxs, int listsize,
                             int32_t *lists, int32_t *counts, int32_t *lastLeafs);
int cmgpu_point_leafnum_batch_device(cmgpu_ctx *ctx, int n, const float *points, int32_t *leafnums, void *stream);
int cmgpu_box_leafnums_batch_device(cmgpu_ctx *ctx, int n, const float *mins, const float *maxs, int listsize,
                             int32_t *lists, int32_t *counts, int32_t *lastLeafs, void *stream);

/* ---- entity linking, PVS part, and snapshot culling (server/sv_world.c:255-303, sv_snapshot.c:316-355) ----
 * What SV_LinkEntity stores in the svEntity_t (server.h:42-45) for the box absmin..absmax, computed in ONE tree
 * walk per entity (the 128-leaf list of sv_world.c:178-181 never leaves the thread). */
#define CMGPU_MAX_AREAS          256     /* limit of this library (the q3map tools stop at 0x100 areas) */
#define CMGPU_MAX_ENT_CLUSTERS   16      /* server.h:35 */
#define CMGPU_MAX_ENT_LEAFS      128     /* MAX_TOTAL_ENT_LEAFS, sv_world.c:178 */
typedef struct cmgpu_entity_clusters_s {
	int32_t numLeafs;                    /* leafs CM_BoxLeafnums stored (<= 128); 0 = outside the world, not linked */
	int32_t areanum, areanum2;
	int32_t numClusters;
	int32_t lastCluster;                 /* set when the clusters did not fit (sv_world.c:300-303) */
	int32_t clusternums[CMGPU_MAX_ENT_CLUSTERS];   /* entries >= numClusters are 0 */
} cmgpu_entity_clusters_t;               /* 84 bytes */
int cmgpu_entity_clusters_batch(cmgpu_ctx *ctx, int n, const float *absmins, const float *absmaxs,
                                cmgpu_entity_clusters_t *out);
int cmgpu_entity_clusters_batch_device(cmgpu_ctx *ctx, int n, const float *absmins, const float *absmaxs,
                                cmgpu_entity_clusters_t *out, void *stream);

/* Area portals (cm_test.c:322-424).  The context keeps the reference counts and the flood numbers;
 * cmgpu_upload resets them like CM_LoadMap does (all closed, cm_load.c:734). */
int cmgpu_adjust_area_portal_state(cmgpu_ctx *ctx, int area1, int area2, int open);
int cmgpu_areas_connected(cmgpu_ctx *ctx, int area1, int area2);          /* 1 / 0, < 0: error code */
/* CM_WriteAreaBits (cm_test.c:441-470): ORs into buffer, returns the byte count (< 0: error code) */
int cmgpu_write_area_bits(cmgpu_ctx *ctx, uint8_t *buffer, int area);
/* cm_noAreas (cm_load.c:640): when set every pair of areas counts as connected */
int cmgpu_set_no_areas(cmgpu_ctx *ctx, int noAreas);

/* The PVS stage of SV_AddEntitiesVisibleFromPoint (sv_snapshot.c:316-355) for every (viewpoint, entity) pair:
 * bit e%32 of visible[v * ((nEnts+31)/32) + e/32] is set when entity e passes the area and cluster tests seen
 * from origins[v].  viewLeafs (may be NULL) receives CM_PointLeafnum(origins[v]).  ents are records made by
 * cmgpu_entity_clusters_batch. */
int cmgpu_visible_from_points_batch(cmgpu_ctx *ctx, int nViews, const float *origins, int nEnts,
                                    const cmgpu_entity_clusters_t *ents, uint32_t *visible, int32_t *viewLeafs);
int cmgpu_visible_from_points_batch_device(cmgpu_ctx *ctx, int nViews, const float *origins, int nEnts,
                                    const cmgpu_entity_clusters_t *ents, uint32_t *visible, int32_t *viewLeafs, void *stream);

/* ---- instrumentation (successor of c_traces/c_brush_traces, cm_load.c:60-61) */
typedef struct cmgpu_counters_s {
	unsigned long long traces, nodes, leafs, brushRefs, brushBounds, brushSides, crossings, splits,
	                   patches, patchPlanes, facets;
} cmgpu_counters;
/* enable != 0 switches the trace kernels to their counting variant (slower). */
int cmgpu_enable_counters(cmgpu_ctx *ctx, int enable);
int cmgpu_read_counters(cmgpu_ctx *ctx, cmgpu_counters *out, int reset);

/* number of kernels this context has launched (for bench.py's gpu_launches) */
unsigned long long cmgpu_launch_count(const cmgpu_ctx *ctx);
/* CUDA-event timing of the trace kernel alone, on the stream it is launched on (the binning kernels that
 * precede it are outside the bracket).  cmgpu_kernel_time_ms waits for the bracketed launches, returns the
 * sum of their durations and how many there were, and starts a new measurement. */
int cmgpu_kernel_timing(cmgpu_ctx *ctx, int enable);
int cmgpu_kernel_time_ms(cmgpu_ctx *ctx, double *total_ms, int *launches);
/* The denominator of the L2 working-set roofline: GB/s at which this GPU's SMs can READ a buffer of `bytes` that stays in
 * L2 (16-byte loads, `passes` sweeps inside one kernel, CUDA-event timed).  bytes <= 64 MiB. */
int cmgpu_measure_l2_read(cmgpu_ctx *ctx, size_t bytes, int passes, double *gbs);

/* ---- multi-GPU: a result buffer shared by the processes of one node ------------
 * One process per GPU.  The rank that wants all records allocates the buffer (cmgpu_ipc_alloc) and
 * passes the 64-byte handle to the others by whatever control plane it has; they open it (cmgpu_ipc_open) and give
 * `pointer + 56 * first_item_of_my_range` as `results` to the device-pointer entry points: their kernels store the
 * records into the owner's HBM over NVLink as they finish items, so there is no gather pass.  A barrier after each
 * rank's stream synchronisation makes the records visible to the owner.                                           */
int cmgpu_ipc_alloc(cmgpu_ctx *ctx, size_t bytes, void **devptr, unsigned char handle[64]);
int cmgpu_ipc_open(cmgpu_ctx *ctx, const unsigned char handle[64], void **devptr);
int cmgpu_ipc_close(cmgpu_ctx *ctx, void *devptr);
int cmgpu_ipc_free(cmgpu_ctx *ctx, void *devptr);

/* ---- botlib's AAS tracer -----------------------------------
 * AAS_TraceClientBBox (code/botlib/be_aas_sample.c:399-665) and AAS_PointAreaNum (:212-258), batched: the line traces a
 * bot's movement prediction issues (be_aas_move.c:644,772,923) against the AAS file's own tree.  The world is handed
 * over as the three arrays the tracer reads of aas_t (be_aas_def.h): planes [numplanes][5] floats (normal, dist, type;
 * stored in opposite pairs, plane n ^ 1 is plane n flipped), nodes [numnodes][3] ints (planenum, children[0],
 * children[1]; node 0 is the dummy, the root is node 1; child > 0 a node, 0 a solid leaf, < 0 minus an area number) and
 * areasettings[].presencetype per area (area 0 is the dummy).  Entity collisions (passent >= 0, :472-487) are server
 * traces -- cmgpu_sv_trace_batch -- and not part of this call: it is AAS_TraceClientBBox(start, end, presencetype, -1).
 * Results are aas_trace_t records (be_aas.h:84-93), byte for byte.  Independent of the clip map: cmgpu_upload /
 * cmgpu_clear do not touch it. */
typedef struct {
	int32_t startsolid;
	float   fraction;
	float   endpos[3];
	int32_t ent;
	int32_t lastarea;
	int32_t area;
	int32_t planenum;
} cmgpu_aas_trace_t;                 /* 36 bytes */

/* areacontents: areasettings[].contents per area (AREACONTENTS_*, aasfile.h:73-89), or NULL for none */
int cmgpu_aas_upload(cmgpu_ctx *ctx, int numplanes, const float *planes, int numnodes, const int32_t *nodes, int numareas,
                     const int32_t *presencetypes, const int32_t *areacontents);
/* presencetypes: one per item (presence_stride 1) or one for the batch (stride 0: a HOST pointer, also for the _device call) */
int cmgpu_aas_trace_client_bbox_batch(cmgpu_ctx *ctx, int n, const float *starts, const float *ends,
                                      const int32_t *presencetypes, int presence_stride, cmgpu_aas_trace_t *results);
int cmgpu_aas_trace_client_bbox_batch_device(cmgpu_ctx *ctx, int n, const float *starts, const float *ends,
                                             const int32_t *presencetypes, int presence_stride, cmgpu_aas_trace_t *results, void *stream);
int cmgpu_aas_point_area_num_batch(cmgpu_ctx *ctx, int n, const float *points, int32_t *areas);
int cmgpu_aas_point_area_num_batch_device(cmgpu_ctx *ctx, int n, const float *points, int32_t *areas, void *stream);

/* AAS_PredictClientMovement (code/botlib/be_aas_move.c:495-979), batched: one bot's movement predicted up to maxframes frames
 * ahead -- gravity, friction, acceleration by the command, a slide along the AAS with step-ups -- until one of the events in
 * stopevent happens (SE_*, be_aas.h:163-176; SE_HITBOUNDINGBOX tests the fixed 8-unit box of :973-974).  It is the call
 * AAS_PredictClientMovement(&move, -1, origin, presencetype, onground, velocity, cmdmove, cmdframes, maxframes, frametime,
 * stopevent, stopareanum, qfalse): entnum -1, i.e. no entity collisions (those are server traces).  The contents of the world
 * (AAS_PointContents -> botimport.PointContents, be_aas_bspq3.c:138) are CM_PointContents of the clip map uploaded to the same
 * context (0 everywhere when there is none).  moves: aas_clientmove_t records (be_aas.h:178-189), byte for byte; results: the
 * function's return value per item (0 only when the slide loop gives up, :875), may be NULL. */
typedef struct {
	float phys_friction, phys_stopspeed, phys_gravity, phys_waterfriction, phys_watergravity, phys_maxwalkvelocity, phys_maxcrouchvelocity,
	      phys_maxswimvelocity, phys_walkaccelerate, phys_airaccelerate, phys_swimaccelerate, phys_maxstep, phys_maxsteepness, phys_maxbarrier,
	      phys_jumpvel;
} cmgpu_aas_settings;                /* the fields of aas_settings_t (be_aas_def.h:86-104) the predictor reads */
void cmgpu_aas_default_settings(cmgpu_aas_settings *settings);            /* AAS_InitSettings' defaults, be_aas_move.c:74-92 */
int  cmgpu_aas_set_settings(cmgpu_ctx *ctx, const cmgpu_aas_settings *settings);

typedef struct {
	const float   *origins, *velocities, *cmdmoves;      /* [n][3] */
	const float   *frametimes;                           /* [n] */
	const int32_t *presencetypes, *ongrounds, *cmdframes, *maxframes, *stopevents, *stopareanums;   /* [n] */
} cmgpu_aas_moves;

typedef struct {
	float   endpos[3];
	int32_t endarea;
	float   velocity[3];
	cmgpu_aas_trace_t trace;
	int32_t presencetype, stopevent, endcontents;
	float   time;
	int32_t frames;
} cmgpu_aas_clientmove_t;            /* 84 bytes */

int cmgpu_aas_predict_client_movement_batch(cmgpu_ctx *ctx, int n, const cmgpu_aas_moves *in, cmgpu_aas_clientmove_t *moves, int32_t *results);
int cmgpu_aas_predict_client_movement_batch_device(cmgpu_ctx *ctx, int n, const cmgpu_aas_moves *in, cmgpu_aas_clientmove_t *moves,
                                                   int32_t *results, void *stream);

#pragma GCC visibility pop
#ifdef __cplusplus
}
#endif
#endif /* CM_GPU_H */
