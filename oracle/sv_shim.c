/*
 * oracle/sv_shim.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Drives the UNMODIFIED reference server world code (/root/reference/code/server/sv_world.c:
 * SV_ClearWorld, SV_LinkEntity, SV_UnlinkEntity, SV_AreaEntities, SV_Trace, SV_ClipToEntity,
 * SV_PointContents) inside oracle/_ref/libcmref.so, on top of the reference clip-map objects that
 * library already holds.  sv_world.c is compiled where it lies (oracle/Makefile); this file supplies
 * what it leaves undefined --
 *     sv                      the server_t it reads (svEntities[], gentities, gentitySize, state),
 *     SV_GentityNum, SV_SvEntityForGentity, SV_GEntityForSvEntity
 *                             the three accessors of server/sv_game.c:17-42 (index arithmetic on sv),
 * -- and a batch API (svref_*) that the parity tests of the device SV_Trace call.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's reference / cpu_baseline legs may load the library.
 */
#define _GNU_SOURCE
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "server.h"

#define API __attribute__((visibility("default")))

server_t sv;                       /* server/sv_main.c:27 in the engine; here: only the fields sv_world.c touches */
static sharedEntity_t *s_gents;    /* the "game module's" entity array */

sharedEntity_t *SV_GentityNum(int num) {
	return (sharedEntity_t *)((byte *)sv.gentities + sv.gentitySize * num);
}

svEntity_t *SV_SvEntityForGentity(sharedEntity_t *gEnt) {
	if (!gEnt || gEnt->s.number < 0 || gEnt->s.number >= MAX_GENTITIES) {
		Com_Error(ERR_DROP, "SV_SvEntityForGentity: bad gEnt");
	}
	return &sv.svEntities[gEnt->s.number];
}

sharedEntity_t *SV_GEntityForSvEntity(svEntity_t *svEnt) {
	return SV_GentityNum((int)(svEnt - sv.svEntities));
}

/* ---- batch API ------------------------------------------------------------------ */

/* A map must be loaded (cmref_load_map).  Unlinks everything and rebuilds the sector tree. */
API void svref_clear_world(void) {
	int i;
	if (!s_gents) {
		s_gents = calloc(MAX_GENTITIES, sizeof(sharedEntity_t));
	}
	memset(&sv, 0, sizeof(sv));
	memset(s_gents, 0, MAX_GENTITIES * sizeof(sharedEntity_t));
	sv.gentities = s_gents;
	sv.gentitySize = (int)sizeof(sharedEntity_t);
	sv.num_entities = MAX_GENTITIES;
	sv.state = SS_GAME;
	for (i = 0; i < MAX_GENTITIES; i++) {
		s_gents[i].s.number = i;
		s_gents[i].r.ownerNum = ENTITYNUM_NONE;
	}
	SV_ClearWorld();
}

/* fill in the fields the game sets before trap_LinkEntity, then link (returns r.linked) */
API int svref_link_entity(int num, int bmodel, int modelindex, const float *mins, const float *maxs,
                          const float *origin, const float *angles, int contents, int ownerNum) {
	sharedEntity_t *e;
	if (num < 0 || num >= MAX_GENTITIES || !s_gents) return -1;
	e = &s_gents[num];
	e->r.bmodel = bmodel ? qtrue : qfalse;
	e->s.modelindex = modelindex;
	VectorCopy(mins, e->r.mins);
	VectorCopy(maxs, e->r.maxs);
	VectorCopy(origin, e->r.currentOrigin);
	VectorCopy(angles, e->r.currentAngles);
	e->r.contents = contents;
	e->r.ownerNum = ownerNum;
	SV_LinkEntity(e);
	return e->r.linked;
}

API void svref_set_owner(int num, int ownerNum) {
	if (num >= 0 && num < MAX_GENTITIES && s_gents) s_gents[num].r.ownerNum = ownerNum;
}

API void svref_unlink_entity(int num) {
	if (num >= 0 && num < MAX_GENTITIES && s_gents) SV_UnlinkEntity(&s_gents[num]);
}

/* what linking derived: absmin[3], absmax[3]; returns r.linked */
API int svref_entity_bounds(int num, float *absmin, float *absmax) {
	const sharedEntity_t *e;
	if (num < 0 || num >= MAX_GENTITIES || !s_gents) return -1;
	e = &s_gents[num];
	VectorCopy(e->r.absmin, absmin);
	VectorCopy(e->r.absmax, absmax);
	return e->r.linked;
}

/* what the PVS part of SV_LinkEntity (sv_world.c:255-303) left in the svEntity_t:
 * out[0..3] = areanum, areanum2, numClusters, lastCluster; out[4..19] = clusternums[16] */
API int svref_entity_clusters(int num, int *out) {
	const svEntity_t *se;
	int i;
	if (num < 0 || num >= MAX_GENTITIES || !s_gents) return -1;
	se = &sv.svEntities[num];
	out[0] = se->areanum; out[1] = se->areanum2; out[2] = se->numClusters; out[3] = se->lastCluster;
	for (i = 0; i < MAX_ENT_CLUSTERS; i++) out[4 + i] = se->clusternums[i];
	return s_gents[num].r.linked;
}

/* The PVS stage of SV_AddEntitiesVisibleFromPoint (sv_snapshot.c:316-355; that function is static and walks
 * svs.currFrame, so its culling lines are REPEATED here over the svEntity_t the reference's SV_LinkEntity filled,
 * calling the reference's own CM_PointLeafnum / CM_LeafArea / CM_LeafCluster / CM_ClusterPVS / CM_AreasConnected).
 * visible[v * nEnts + e] = 1 when entity nums[e] passes for viewpoint v. */
API void svref_visible_from_points(int nViews, const float *origins, int nEnts, const int *nums, unsigned char *visible) {
	int v, e, i, l;
	for (v = 0; v < nViews; v++) {
		int leafnum = CM_PointLeafnum(origins + 3 * (size_t)v);
		int clientarea = CM_LeafArea(leafnum);
		int clientcluster = CM_LeafCluster(leafnum);
		byte *clientpvs = CM_ClusterPVS(clientcluster);
		for (e = 0; e < nEnts; e++) {
			const svEntity_t *svEnt = &sv.svEntities[nums[e]];
			byte *bitvector = clientpvs;
			unsigned char *out = &visible[(size_t)v * nEnts + e];
			*out = 0;
			if (!CM_AreasConnected(clientarea, svEnt->areanum)) {
				if (!CM_AreasConnected(clientarea, svEnt->areanum2)) continue;
			}
			if (!svEnt->numClusters) continue;
			l = 0;
			for (i = 0; i < svEnt->numClusters; i++) {
				l = svEnt->clusternums[i];
				if (bitvector[l >> 3] & (1 << (l & 7))) break;
			}
			if (i == svEnt->numClusters) {
				if (svEnt->lastCluster) {
					for (; l <= svEnt->lastCluster; l++) {
						if (bitvector[l >> 3] & (1 << (l & 7))) break;
					}
					if (l == svEnt->lastCluster) continue;
				} else {
					continue;
				}
			}
			*out = 1;
		}
	}
}

API int svref_area_entities(const float *mins, const float *maxs, int *list, int maxcount) {
	return SV_AreaEntities(mins, maxs, list, maxcount);
}

/* n calls of SV_Trace; hull_stride/pass_stride/mask_stride 0 = one shared value */
API void svref_trace_batch(trace_t *results, int n, const float *starts, const float *ends,
                           const float *mins, const float *maxs, int hull_stride,
                           const int *passEntityNums, int pass_stride, const int *contentmasks, int mask_stride) {
	int i;
	for (i = 0; i < n; i++) {
		const float *mn = mins ? mins + 3 * (size_t)i * hull_stride : NULL;
		const float *mx = maxs ? maxs + 3 * (size_t)i * hull_stride : NULL;
		SV_Trace(&results[i], starts + 3 * (size_t)i, mn, mx, ends + 3 * (size_t)i,
		         passEntityNums[(size_t)i * pass_stride], contentmasks[(size_t)i * mask_stride]);
	}
}

API void svref_point_contents_batch(int *contents, int n, const float *points, const int *passEntityNums, int pass_stride) {
	int i;
	for (i = 0; i < n; i++) {
		contents[i] = SV_PointContents(points + 3 * (size_t)i, passEntityNums[(size_t)i * pass_stride]);
	}
}

API void svref_clip_to_entity_batch(trace_t *results, int n, const float *starts, const float *ends,
                                    const float *mins, const float *maxs, const int *entityNums, int contentmask) {
	int i;
	for (i = 0; i < n; i++) {
		SV_ClipToEntity(&results[i], starts + 3 * (size_t)i, mins, maxs, ends + 3 * (size_t)i, entityNums[i], contentmask);
	}
}

/* wall-clock of n SV_Trace calls on this thread (for scripts/sv_bench.py) */
API double svref_time_traces(int n, const float *starts, const float *ends, const float *mins, const float *maxs,
                             const int *passEntityNums, int contentmask) {
	struct timespec a, b;
	trace_t tr;
	int i;
	clock_gettime(CLOCK_MONOTONIC, &a);
	for (i = 0; i < n; i++) {
		SV_Trace(&tr, starts + 3 * (size_t)i, mins, maxs, ends + 3 * (size_t)i, passEntityNums[i], contentmask);
	}
	clock_gettime(CLOCK_MONOTONIC, &b);
	return (double)(b.tv_sec - a.tv_sec) + 1e-9 * (double)(b.tv_nsec - a.tv_nsec);
}
