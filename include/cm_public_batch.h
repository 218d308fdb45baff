/*
 * cm_public_batch.h -- the reference's clip-map API (code/qcommon/cm_public.h:26-69)
 * served by the GPU engine, plus the batched entry points the engine adds.
 *
 * The functions below keep the reference's names, argument order and meaning
 * (note: no `capsule` argument, as in ThreeCore's cm_public.h:45-52), so
 * SV_Trace (server/sv_world.c:562), trap_Trace (server/sv_game.c:198), the
 * cgame syscalls (client/cl_cgame.c:248-252) and botlib's import callbacks
 * (server/sv_bot.c:140-195) can call them unchanged.  Each single-item call is
 * a batch of one; the batch calls are where the GPU pays off.
 *
 * State: like the reference (one global clipMap_t, cm_load.c:54-59) this layer
 * owns one process-wide engine context on one GPU; it is single-threaded and
 * non-reentrant like the code it replaces.  Errors the reference raises with
 * Com_Error(ERR_DROP, ...) are delivered to the handler installed with
 * CM_SetErrorHandler (the engine passes a function that calls Com_Error); with
 * no handler installed the message is kept in CM_LastError() and the call
 * returns a zeroed result.
 */
#ifndef CM_PUBLIC_BATCH_H
#define CM_PUBLIC_BATCH_H

#include "cm_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif
#pragma GCC visibility push(default)   /* the library is built with -fvisibility=hidden */

typedef float vec_t;
typedef vec_t vec3_t[3];           /* q_shared.h:369 */
typedef int clipHandle_t;          /* q_shared.h:228 */
typedef int qboolean;
typedef unsigned char byte;       /* q_shared.h */
typedef cmgpu_trace_t trace_t;     /* q_shared.h:976-986, same 56 bytes */

#define ENTITYNUM_NONE   4095      /* q_shared.h: MAX_GENTITIES-1 */
#define ENTITYNUM_WORLD  4094

typedef void (*cm_error_fn)(const char *message);
void        CM_SetErrorHandler(cm_error_fn fn);
const char *CM_LastError(void);
int         CM_SetDevice(int device);                 /* before the first load; default 0 */
cmgpu_ctx  *CM_Context(void);                         /* the context behind this layer (NULL before a load) */

/* ---- map (cm_public.h:26-41) ---------------------------------------------- */
/* CM_LoadMap with FS_ReadFile replaced by caller-supplied bytes (the dedicated
 * server hands over what FS_ReadFile returned). Returns 0 like the reference's
 * first-map case and uploads the map to the GPU. */
int          CM_LoadMapFromMemory(const char *name, const void *bsp, int length, int *checksum);
void         CM_ClearMap(void);
clipHandle_t CM_InlineModel(int index);
clipHandle_t CM_TempBoxModel(const vec3_t mins, const vec3_t maxs);
void         CM_ModelBounds(clipHandle_t model, vec3_t mins, vec3_t maxs);
int          CM_NumInlineModels(void);
void         CM_SetCvars(int cm_noCurves, int cm_playerCurveClip);   /* cm_load.c:640-642 */

/* ---- single queries (cm_public.h:42-52) ----------------------------------- */
int  CM_PointContents(const vec3_t p, clipHandle_t model);
int  CM_TransformedPointContents(const vec3_t p, clipHandle_t model, const vec3_t origin, const vec3_t angles);
void CM_Trace(trace_t *results, const vec3_t start, const vec3_t end, const vec3_t mins, const vec3_t maxs,
              clipHandle_t model, const vec3_t origin, int brushmask);
void CM_BoxTrace(trace_t *results, const vec3_t start, const vec3_t end, const vec3_t mins, const vec3_t maxs,
                 clipHandle_t model, int brushmask);
void CM_TransformedBoxTrace(trace_t *results, const vec3_t start, const vec3_t end, const vec3_t mins, const vec3_t maxs,
                            clipHandle_t model, int brushmask, const vec3_t origin, const vec3_t angles);

/* ---- batched queries (new) -------------------------------------------------- */
/* n sweeps of one hull through one model with one mask; starts/ends are [n][3]. Returns a CMGPU_* status. */
int CM_BoxTraceBatch(trace_t *results, int n, const float *starts, const float *ends,
                     const vec3_t mins, const vec3_t maxs, clipHandle_t model, int brushmask);
/* per-item hull (hull_stride 1), model and mask (mask_stride 1); boxmins/boxmaxs give the
 * CM_TempBoxModel box of every item whose model is BOX_MODEL_HANDLE (1023). */
int CM_BoxTraceBatchEx(trace_t *results, int n, const float *starts, const float *ends,
                       const float *mins, const float *maxs, int hull_stride,
                       const clipHandle_t *models, const int *brushmasks, int mask_stride,
                       const float *boxmins, const float *boxmaxs);
int CM_TransformedBoxTraceBatch(trace_t *results, int n, const float *starts, const float *ends,
                                const float *mins, const float *maxs, int hull_stride,
                                const clipHandle_t *models, const int *brushmasks, int mask_stride,
                                const float *origins, const float *angles,
                                const float *boxmins, const float *boxmaxs);
/* origins/angles NULL: CM_PointContents per item, else CM_TransformedPointContents */
int CM_PointContentsBatch(int *contents, int n, const float *points, const clipHandle_t *models,
                          const float *origins, const float *angles,
                          const float *boxmins, const float *boxmaxs);

/* ---- leaf queries (cm_public.h:54-62), single and batched ------------------------------------
 * What SV_LinkEntity asks per entity (sv_world.c:266-303) and the snapshot code per client. */
int CM_PointLeafnum(const vec3_t p);
int CM_BoxLeafnums(const vec3_t mins, const vec3_t maxs, int *list, int listsize, int *lastLeaf);
int CM_LeafCluster(int leafnum);
int CM_LeafArea(int leafnum);
int CM_PointLeafnumBatch(int *leafnums, int n, const float *points);
/* lists is [n][listsize]; counts[i] leafs were stored for box i, lastLeafs[i] as CM_BoxLeafnums' *lastLeaf */
int CM_BoxLeafnumsBatch(int n, const float *mins, const float *maxs, int listsize, int *lists, int *counts, int *lastLeafs);

/* ---- visibility and area portals (cm_public.h:56-69) -----------------------------------------
 * Table lookups on the loaded map (no tree work): served from the host copy of the map.  The batched,
 * device-side consumers of the same tables are SV_EntityClustersBatch / SV_VisibleFromPointsBatch below. */
int   CM_NumClusters(void);
byte *CM_ClusterPVS(int cluster);                        /* row of the visibility matrix; all ones when not vised */
void  CM_AdjustAreaPortalState(int area1, int area2, qboolean open);
qboolean CM_AreasConnected(int area1, int area2);
int   CM_WriteAreaBits(byte *buffer, int area);

/* ---- server: entity linking (PVS part) and snapshot culling ----------------------------------
 * SV_EntityClustersBatch: what SV_LinkEntity (server/sv_world.c:255-303) stores in svEntity_t.{numClusters,
 * clusternums, lastCluster, areanum, areanum2} for n boxes absmin..absmax, one device thread per entity.
 * SV_VisibleFromPointsBatch: the PVS stage of SV_AddEntitiesVisibleFromPoint (server/sv_snapshot.c:316-355) for
 * every (viewpoint, entity) pair; bit e%32 of visible[v*((nEnts+31)/32) + e/32].  SVF_BROADCAST, singleClient,
 * draw distance and the anti-cheat traces stay with the caller (they need game state; the traces are SV_TraceBatch). */
typedef cmgpu_entity_clusters_t sv_entity_clusters_t;
int SV_EntityClustersBatch(int n, const float *absmins, const float *absmaxs, sv_entity_clusters_t *out);
int SV_VisibleFromPointsBatch(int nViews, const float *origins, int nEnts, const sv_entity_clusters_t *ents,
                              unsigned int *visible, int *viewLeafs);

/* ---- server hook: SV_ClipMoveToEntities (server/sv_world.c:476-551) ----------
 * The caller keeps the broad phase and the pass/owner/contents filters
 * (sv_world.c:486-520, they need game state) and hands over the surviving
 * candidates in touch-list order.  All candidates are clipped in ONE batch,
 * then merged with the reference's rules (sv_world.c:532-549), including the
 * early stop once the accumulated trace is allsolid. */
typedef struct cm_touch_s {
	int          entityNum;        /* touch->s.number */
	qboolean     bmodel;           /* touch->r.bmodel: inline model, else bounding box */
	int          modelindex;       /* touch->s.modelindex (bmodel) */
	vec3_t       mins, maxs;       /* touch->r.mins/maxs (box entities) */
	vec3_t       origin, angles;   /* touch->r.currentOrigin / currentAngles */
} cm_touch_t;

int SV_ClipMoveToEntitiesBatch(trace_t *clipTrace, const vec3_t start, const vec3_t end,
                               const vec3_t mins, const vec3_t maxs, int contentmask,
                               int numTouch, const cm_touch_t *touch);

/* many moves at once: move m owns touch[touchFirst[m] .. touchFirst[m+1]).  clipTraces[m] must
 * hold the world trace on entry (SV_Trace, sv_world.c:576-577) and holds the merged result on exit. */
int SV_ClipMovesToEntitiesBatch(trace_t *clipTraces, int numMoves, const float *starts, const float *ends,
                                const float *mins, const float *maxs, int hull_stride,
                                const int *contentmasks, int mask_stride,
                                const int *touchFirst, const cm_touch_t *touch);

/* SV_PointContents (server/sv_world.c:616-647) for many points: point m owns touch[touchFirst[m] .. touchFirst[m+1])
 * (the caller's SV_AreaEntities result minus passEntityNum); contents[m] = world contents | every candidate's. */
int SV_PointContentsBatch(int *contents, int numPoints, const float *points, const int *touchFirst, const cm_touch_t *touch);

#pragma GCC visibility pop
#ifdef __cplusplus
}
#endif
#endif /* CM_PUBLIC_BATCH_H */
