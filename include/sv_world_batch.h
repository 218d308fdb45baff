/*
 * sv_world_batch.h -- the reference's server world interface (code/server/sv_world.c, declared in
 * code/server/server.h) with SV_Trace and SV_PointContents turned into batch calls that run WHOLE on the GPU:
 * world clip, entity broad phase, per-entity clips and the merge (cm_gpu.h: cmgpu_sv_*).
 *
 * Entity linking keeps the reference's names and behaviour -- SV_ClearWorld (sv_world.c:121-132), SV_LinkEntity
 * (:203-328), SV_UnlinkEntity (:141-172), SV_AreaEntities (:383-396) -- on an index-based sector tree, so that
 * a host which has no sv_world.c of its own (a rollout driver) can use them, and so that the parity tests read
 * like the reference: link entities, trace, compare.  SV_LinkEntity's PVS part (cluster and area numbers,
 * :266-303) is snapshot culling, not collision, and is not mirrored.
 *
 * An engine that keeps its own sv_world.c needs none of the linking functions: it hands its linked entities to
 * cmgpu_sv_set_entities once per frame.
 */
#ifndef SV_WORLD_BATCH_H
#define SV_WORLD_BATCH_H

#include "cm_public_batch.h"

#ifdef __cplusplus
extern "C" {
#endif
#pragma GCC visibility push(default)

#define MAX_GENTITIES 4096          /* 1 << GENTITYNUM_BITS, q_shared.h:1054-1055 */

/* the fields of sharedEntity_t (game/g_public.h:13-50) that sv_world.c reads or writes */
typedef struct sv_entity_s {
	int       number;                /* s.number: index in the game's entity array */
	int       modelindex;            /* s.modelindex: inline model of a bmodel entity */
	qboolean  linked;                /* r.linked      (written by SV_LinkEntity / SV_UnlinkEntity) */
	int       linkcount;             /* r.linkcount   (written) */
	qboolean  bmodel;                /* r.bmodel */
	vec3_t    mins, maxs;            /* r.mins, r.maxs */
	int       contents;              /* r.contents */
	vec3_t    absmin, absmax;        /* r.absmin, r.absmax (written) */
	vec3_t    currentOrigin;         /* r.currentOrigin */
	vec3_t    currentAngles;         /* r.currentAngles */
	int       ownerNum;              /* r.ownerNum */
} sv_entity_t;

/* SV_LocateGameData (sv_game.c): where the game keeps its entities; entity i is at (char *)gEnts + i * sizeofGEntity
 * and starts with an sv_entity_t. */
void SV_LocateGameData(sv_entity_t *gEnts, int numGEntities, int sizeofGEntity);
sv_entity_t *SV_GentityNum(int num);

void SV_ClearWorld(void);
/* the same with the world bounds given (what CM_ModelBounds(0) returns): linking without a loaded map */
void SV_ClearWorldForBounds(const vec3_t mins, const vec3_t maxs);
void SV_LinkEntity(sv_entity_t *gEnt);
void SV_UnlinkEntity(sv_entity_t *gEnt);
int  SV_AreaEntities(const vec3_t mins, const vec3_t maxs, int *entityList, int maxcount);

/* n calls of SV_Trace(results, start, mins, maxs, end, passEntityNum, contentmask) (sv_world.c:562-609);
 * hull_stride / pass_stride / mask_stride 0 = one shared value.  mins/maxs may be NULL (:565-570). */
int SV_TraceBatch(trace_t *results, int n, const float *starts, const float *mins, const float *maxs, int hull_stride,
                  const float *ends, const int *passEntityNums, int pass_stride, const int *contentmasks, int mask_stride);
/* n calls of SV_PointContents(p, passEntityNum) (sv_world.c:616-647) */
int SV_PointContentsWorldBatch(int *contents, int n, const float *points, const int *passEntityNums, int pass_stride);
/* entity clips the last of the two calls above generated */
unsigned long long SV_LastBatchCandidates(void);

#pragma GCC visibility pop
#ifdef __cplusplus
}
#endif
#endif /* SV_WORLD_BATCH_H */
