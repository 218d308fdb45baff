// cm_api.cpp -- the reference's CM_* interface (code/qcommon/cm_public.h:26-69) on
// top of the cmgpu C ABI, plus the batched entry points and the server-side
// batching hook for SV_ClipMoveToEntities (code/server/sv_world.c:476-551).
//
// Everything here is host-side glue: it owns the process-wide context (the
// counterpart of the reference's global clipMap_t, cm_load.c:54-59), keeps the
// CM_TempBoxModel box (cm_load.c:927-949) and forwards work to the GPU.  There is
// no CPU tracing code in this file.
#include "../../include/cm_public_batch.h"

#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

namespace {

cmgpu_ctx *g_ctx = nullptr;
cmgpu_flatmap *g_map = nullptr;
int g_device = 0;
cm_error_fn g_onError = nullptr;
std::string g_lastError;
float g_boxMins[3] = {0, 0, 0}, g_boxMaxs[3] = {0, 0, 0};
int g_noCurves = 0, g_playerCurveClip = 1;

void raise(const char *fmt, ...) {
	char buf[600];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	g_lastError = buf;
	if (g_onError) g_onError(buf);     // the engine's handler calls Com_Error(ERR_DROP, ...) and does not return
}

// a compute status other than OK / NO_MAP is an ERR_DROP in the reference's terms
bool dropOn(int rc, const char *where) {
	if (rc == CMGPU_OK) return false;
	raise("%s: %s", where, g_ctx ? cmgpu_last_error(g_ctx) : cmgpu_last_error(nullptr));
	return true;
}

void defaultTrace(trace_t *t) {          // cm_trace.c:558-566: zeroed, fraction 1
	memset(t, 0, sizeof(*t));
	t->fraction = 1.0f;
}

}  // namespace

extern "C" {

void CM_SetErrorHandler(cm_error_fn fn) { g_onError = fn; }
const char *CM_LastError(void) { return g_lastError.c_str(); }
cmgpu_ctx *CM_Context(void) { return g_ctx; }

int CM_SetDevice(int device) {
	if (g_ctx) return CMGPU_ERR_INVALID;
	g_device = device;
	return CMGPU_OK;
}

void CM_SetCvars(int noCurves, int playerCurveClip) {
	g_noCurves = noCurves;
	g_playerCurveClip = playerCurveClip;
	if (g_ctx) cmgpu_set_cvars(g_ctx, noCurves, playerCurveClip);
}

int CM_LoadMapFromMemory(const char *name, const void *bsp, int length, int *checksum) {
	if (!name || !name[0]) { raise("CM_LoadMap: NULL name"); return 0; }
	if (checksum) *checksum = 0;                 // md4 is gone in the reference too (cm_load.c:692)
	if (!g_ctx) {
		if (dropOn(cmgpu_create(&g_ctx, g_device), "CM_LoadMap")) { g_ctx = nullptr; return 0; }
		cmgpu_set_cvars(g_ctx, g_noCurves, g_playerCurveClip);
	}
	char err[256];
	cmgpu_flatmap *map = nullptr;
	int rc = cmgpu_flatmap_from_bsp(bsp, length < 0 ? 0 : (size_t)length, &map, err, sizeof(err));
	if (rc) { raise("%s", err); return 0; }
	if (dropOn(cmgpu_upload(g_ctx, map), "CM_LoadMap")) { cmgpu_flatmap_free(map); return 0; }
	if (g_map) cmgpu_flatmap_free(g_map);
	g_map = map;
	return 0;
}

void CM_ClearMap(void) {
	if (g_ctx) cmgpu_clear(g_ctx);
	if (g_map) cmgpu_flatmap_free(g_map);
	g_map = nullptr;
}

int CM_NumInlineModels(void) { return g_map ? g_map->numModels : 0; }

clipHandle_t CM_InlineModel(int index) {
	if (index < 0 || index >= CM_NumInlineModels()) raise("CM_InlineModel: bad number %i >= %i", index, CM_NumInlineModels());
	return index;
}

clipHandle_t CM_TempBoxModel(const vec3_t mins, const vec3_t maxs) {
	memcpy(g_boxMins, mins, sizeof(g_boxMins));
	memcpy(g_boxMaxs, maxs, sizeof(g_boxMaxs));
	return CMGPU_BOX_MODEL_HANDLE;
}

void CM_ModelBounds(clipHandle_t model, vec3_t mins, vec3_t maxs) {
	if (model == CMGPU_BOX_MODEL_HANDLE) {
		memcpy(mins, g_boxMins, sizeof(g_boxMins));
		memcpy(maxs, g_boxMaxs, sizeof(g_boxMaxs));
		return;
	}
	if (!g_map || model < 0 || model >= g_map->numModels) {
		raise("CM_ClipHandleToModel: bad handle %i", model);
		mins[0] = mins[1] = mins[2] = maxs[0] = maxs[1] = maxs[2] = 0;
		return;
	}
	memcpy(mins, g_map->modelBounds + 6 * (size_t)model, 12);
	memcpy(maxs, g_map->modelBounds + 6 * (size_t)model + 3, 12);
}

// ---- batched -------------------------------------------------------------------------
int CM_BoxTraceBatchEx(trace_t *results, int n, const float *starts, const float *ends,
                       const float *mins, const float *maxs, int hull_stride,
                       const clipHandle_t *models, const int *brushmasks, int mask_stride,
                       const float *boxmins, const float *boxmaxs) {
	if (!g_map) {                                  // map not loaded: zeroed traces (cm_trace.c:562-566)
		for (int i = 0; i < n; i++) defaultTrace(&results[i]);
		return CMGPU_OK;
	}
	int rc = cmgpu_box_trace_batch(g_ctx, n, starts, ends, mins, maxs, hull_stride, models, brushmasks, mask_stride,
	                               boxmins, boxmaxs, results, nullptr);
	dropOn(rc, "CM_BoxTraceBatch");
	return rc;
}

int CM_BoxTraceBatch(trace_t *results, int n, const float *starts, const float *ends,
                     const vec3_t mins, const vec3_t maxs, clipHandle_t model, int brushmask) {
	std::vector<clipHandle_t> models;
	std::vector<float> bmn, bmx;
	if (model) {
		models.assign((size_t)n, model);
		if (model == CMGPU_BOX_MODEL_HANDLE) {
			bmn.resize(3 * (size_t)n); bmx.resize(3 * (size_t)n);
			for (int i = 0; i < n; i++) { memcpy(&bmn[3 * (size_t)i], g_boxMins, 12); memcpy(&bmx[3 * (size_t)i], g_boxMaxs, 12); }
		}
	}
	return CM_BoxTraceBatchEx(results, n, starts, ends, mins, maxs, 0, model ? models.data() : nullptr, &brushmask, 0,
	                          bmn.empty() ? nullptr : bmn.data(), bmx.empty() ? nullptr : bmx.data());
}

int CM_TransformedBoxTraceBatch(trace_t *results, int n, const float *starts, const float *ends,
                                const float *mins, const float *maxs, int hull_stride,
                                const clipHandle_t *models, const int *brushmasks, int mask_stride,
                                const float *origins, const float *angles,
                                const float *boxmins, const float *boxmaxs) {
	if (!g_map) {
		// cm_trace.c:562 leaves a zeroed trace; CM_TransformedBoxTrace then writes endpos = start + 1*(end-start)
		for (int i = 0; i < n; i++) {
			defaultTrace(&results[i]);
			for (int k = 0; k < 3; k++)
				results[i].endpos[k] = starts[3 * (size_t)i + k] + results[i].fraction * (ends[3 * (size_t)i + k] - starts[3 * (size_t)i + k]);
		}
		return CMGPU_OK;
	}
	int rc = cmgpu_transformed_box_trace_batch(g_ctx, n, starts, ends, mins, maxs, hull_stride, models, brushmasks, mask_stride,
	                                           origins, angles, boxmins, boxmaxs, results, nullptr);
	dropOn(rc, "CM_TransformedBoxTraceBatch");
	return rc;
}

int CM_PointContentsBatch(int *contents, int n, const float *points, const clipHandle_t *models,
                          const float *origins, const float *angles, const float *boxmins, const float *boxmaxs) {
	if (!g_map) {                                  // cm_test.c:227
		for (int i = 0; i < n; i++) contents[i] = 0;
		return CMGPU_OK;
	}
	int rc = cmgpu_point_contents_batch(g_ctx, n, points, models, origins, angles, boxmins, boxmaxs, contents);
	dropOn(rc, "CM_PointContentsBatch");
	return rc;
}

// ---- single-item calls: batches of one ---------------------------------------------------
void CM_BoxTrace(trace_t *results, const vec3_t start, const vec3_t end, const vec3_t mins, const vec3_t maxs,
                 clipHandle_t model, int brushmask) {
	defaultTrace(results);
	CM_BoxTraceBatchEx(results, 1, start, end, mins, maxs, 0, model ? &model : nullptr, &brushmask, 0, g_boxMins, g_boxMaxs);
}

void CM_Trace(trace_t *results, const vec3_t start, const vec3_t end, const vec3_t mins, const vec3_t maxs,
              clipHandle_t model, const vec3_t origin, int brushmask) {
	(void)origin;       // only stored in tw.modelOrigin and never read by the trace (cm_trace.c:560)
	CM_BoxTrace(results, start, end, mins, maxs, model, brushmask);
}

void CM_TransformedBoxTrace(trace_t *results, const vec3_t start, const vec3_t end, const vec3_t mins, const vec3_t maxs,
                            clipHandle_t model, int brushmask, const vec3_t origin, const vec3_t angles) {
	defaultTrace(results);
	CM_TransformedBoxTraceBatch(results, 1, start, end, mins, maxs, 0, &model, &brushmask, 0, origin, angles, g_boxMins, g_boxMaxs);
}

int CM_PointContents(const vec3_t p, clipHandle_t model) {
	int c = 0;
	CM_PointContentsBatch(&c, 1, p, model ? &model : nullptr, nullptr, nullptr, g_boxMins, g_boxMaxs);
	return c;
}

int CM_TransformedPointContents(const vec3_t p, clipHandle_t model, const vec3_t origin, const vec3_t angles) {
	int c = 0;
	CM_PointContentsBatch(&c, 1, p, &model, origin, angles, g_boxMins, g_boxMaxs);
	return c;
}

// ---- leaf queries (cm_public.h:54-62) --------------------------------------------------------
int CM_PointLeafnumBatch(int *leafnums, int n, const float *points) {
	if (!g_map) {                                  // map not loaded: leaf 0 (cm_test.c:57-59)
		for (int i = 0; i < n; i++) leafnums[i] = 0;
		return CMGPU_OK;
	}
	int rc = cmgpu_point_leafnum_batch(g_ctx, n, points, leafnums);
	dropOn(rc, "CM_PointLeafnum");
	return rc;
}

int CM_BoxLeafnumsBatch(int n, const float *mins, const float *maxs, int listsize, int *lists, int *counts, int *lastLeafs) {
	if (!g_map) { raise("CM_BoxLeafnums: no map loaded"); return CMGPU_ERR_NO_MAP; }
	int rc = cmgpu_box_leafnums_batch(g_ctx, n, mins, maxs, listsize, lists, counts, lastLeafs);
	dropOn(rc, "CM_BoxLeafnums");
	return rc;
}

int CM_PointLeafnum(const vec3_t p) {
	int leaf = 0;
	CM_PointLeafnumBatch(&leaf, 1, p);
	return leaf;
}

int CM_BoxLeafnums(const vec3_t mins, const vec3_t maxs, int *list, int listsize, int *lastLeaf) {
	int count = 0, last = 0;
	CM_BoxLeafnumsBatch(1, mins, maxs, listsize, list, &count, &last);
	if (lastLeaf) *lastLeaf = last;
	return count;
}

int CM_LeafCluster(int leafnum) {                  // cm_load.c: ERR_DROP on a bad number
	if (!g_map || leafnum < 0 || leafnum >= g_map->numLeafs) { raise("CM_LeafCluster: bad number"); return -1; }
	return g_map->leafClusterArea ? g_map->leafClusterArea[2 * (size_t)leafnum] : -1;
}

int CM_LeafArea(int leafnum) {
	if (!g_map || leafnum < 0 || leafnum >= g_map->numLeafs) { raise("CM_LeafArea: bad number"); return -1; }
	return g_map->leafClusterArea ? g_map->leafClusterArea[2 * (size_t)leafnum + 1] : -1;
}

// ---- visibility and area portals (cm_test.c:306-470): table lookups on the host copy of the map ----
int CM_NumClusters(void) { return g_map ? g_map->numClusters : 0; }

byte *CM_ClusterPVS(int cluster) {
	static std::vector<byte> allOnes;
	if (!g_map) return nullptr;
	if (!g_map->vised) {                                   // cm_load.c:503-505: a single row of 0xff
		if (allOnes.size() < (size_t)g_map->clusterBytes + 1) allOnes.assign((size_t)g_map->clusterBytes + 1, 255);
		return allOnes.data();
	}
	byte *vis = const_cast<byte *>(g_map->visibility);
	if (cluster < 0 || cluster >= g_map->numClusters) return vis;
	return vis + (size_t)cluster * (size_t)g_map->clusterBytes;
}

void CM_AdjustAreaPortalState(int area1, int area2, qboolean open) {
	if (!g_ctx || !g_map) return;
	if (cmgpu_adjust_area_portal_state(g_ctx, area1, area2, open ? 1 : 0)) raise("%s", cmgpu_last_error(g_ctx));
}

qboolean CM_AreasConnected(int area1, int area2) {
	if (!g_ctx || !g_map) return 0;
	int r = cmgpu_areas_connected(g_ctx, area1, area2);
	if (r < 0) { raise("%s", cmgpu_last_error(g_ctx)); return 0; }
	return r ? 1 : 0;
}

int CM_WriteAreaBits(byte *buffer, int area) {
	if (!g_ctx || !g_map) return 0;
	int r = cmgpu_write_area_bits(g_ctx, buffer, area);
	if (r < 0) { raise("%s", cmgpu_last_error(g_ctx)); return 0; }
	return r;
}

int SV_EntityClustersBatch(int n, const float *absmins, const float *absmaxs, sv_entity_clusters_t *out) {
	if (!g_map) { raise("SV_LinkEntity: no map loaded"); return CMGPU_ERR_NO_MAP; }
	int rc = cmgpu_entity_clusters_batch(g_ctx, n, absmins, absmaxs, out);
	dropOn(rc, "SV_LinkEntity");
	return rc;
}

int SV_VisibleFromPointsBatch(int nViews, const float *origins, int nEnts, const sv_entity_clusters_t *ents,
                              unsigned int *visible, int *viewLeafs) {
	if (!g_map) { raise("SV_AddEntitiesVisibleFromPoint: no map loaded"); return CMGPU_ERR_NO_MAP; }
	int rc = cmgpu_visible_from_points_batch(g_ctx, nViews, origins, nEnts, ents, visible, viewLeafs);
	dropOn(rc, "SV_AddEntitiesVisibleFromPoint");
	return rc;
}

// ---- SV_ClipMoveToEntities, batched ---------------------------------------------------------
// merge rules of sv_world.c:497-549 applied to already computed per-entity traces
static void mergeEntityTraces(trace_t *clip, int numTouch, const cm_touch_t *touch, const trace_t *traces) {
	for (int i = 0; i < numTouch; i++) {
		if (clip->allsolid) return;
		trace_t tr = traces[i];
		if (tr.allsolid) {
			clip->allsolid = 1;
			tr.entityNum = touch[i].entityNum;
		} else if (tr.startsolid) {
			clip->startsolid = 1;
			tr.entityNum = touch[i].entityNum;
		}
		if (tr.fraction < clip->fraction) {
			const int oldStart = clip->startsolid;
			tr.entityNum = touch[i].entityNum;
			*clip = tr;
			clip->startsolid |= oldStart;
		}
	}
}

int SV_ClipMovesToEntitiesBatch(trace_t *clipTraces, int numMoves, const float *starts, const float *ends,
                                const float *mins, const float *maxs, int hull_stride,
                                const int *contentmasks, int mask_stride,
                                const int *touchFirst, const cm_touch_t *touch) {
	if (numMoves <= 0) return CMGPU_OK;
	const int total = touchFirst[numMoves];
	if (total <= 0) return CMGPU_OK;
	const size_t T = (size_t)total;
	std::vector<float> s(3 * T), e(3 * T), mn(3 * T), mx(3 * T), org(3 * T), ang(3 * T), bmn(3 * T), bmx(3 * T);
	std::vector<int> models(T), masks(T);
	static const float zero3[3] = {0, 0, 0};
	for (int m = 0; m < numMoves; m++) {
		const float *hmn = mins ? mins + 3 * (size_t)m * (hull_stride ? 1 : 0) : zero3;
		const float *hmx = maxs ? maxs + 3 * (size_t)m * (hull_stride ? 1 : 0) : zero3;
		for (int t = touchFirst[m]; t < touchFirst[m + 1]; t++) {
			const cm_touch_t &tc = touch[t];
			memcpy(&s[3 * (size_t)t], starts + 3 * (size_t)m, 12);
			memcpy(&e[3 * (size_t)t], ends + 3 * (size_t)m, 12);
			memcpy(&mn[3 * (size_t)t], hmn, 12);
			memcpy(&mx[3 * (size_t)t], hmx, 12);
			memcpy(&org[3 * (size_t)t], tc.origin, 12);
			memcpy(&ang[3 * (size_t)t], tc.angles, 12);
			memcpy(&bmn[3 * (size_t)t], tc.mins, 12);
			memcpy(&bmx[3 * (size_t)t], tc.maxs, 12);
			// SV_ClipHandleForEntity, sv_world.c:35-43
			models[t] = tc.bmodel ? CM_InlineModel(tc.modelindex) : CMGPU_BOX_MODEL_HANDLE;
			masks[t] = contentmasks[(size_t)m * (mask_stride ? 1 : 0)];
		}
	}
	std::vector<trace_t> traces(T);
	int rc = CM_TransformedBoxTraceBatch(traces.data(), total, s.data(), e.data(), mn.data(), mx.data(), 1,
	                                     models.data(), masks.data(), 1, org.data(), ang.data(), bmn.data(), bmx.data());
	if (rc) return rc;
	for (int m = 0; m < numMoves; m++) {
		mergeEntityTraces(&clipTraces[m], touchFirst[m + 1] - touchFirst[m], touch + touchFirst[m], traces.data() + touchFirst[m]);
	}
	return CMGPU_OK;
}

// SV_PointContents (sv_world.c:616-647) for many points: the world's contents of every point and the contents of
// every (point, candidate entity) pair in two batches, OR-ed per point.
int SV_PointContentsBatch(int *contents, int numPoints, const float *points, const int *touchFirst, const cm_touch_t *touch) {
	if (numPoints <= 0) return CMGPU_OK;
	int rc = CM_PointContentsBatch(contents, numPoints, points, nullptr, nullptr, nullptr, nullptr, nullptr);
	if (rc) return rc;
	const int total = touchFirst[numPoints];
	if (total <= 0) return CMGPU_OK;
	const size_t T = (size_t)total;
	std::vector<float> p(3 * T), org(3 * T), ang(3 * T), bmn(3 * T), bmx(3 * T);
	std::vector<int> models(T), c2(T);
	for (int m = 0; m < numPoints; m++) {
		for (int t = touchFirst[m]; t < touchFirst[m + 1]; t++) {
			const cm_touch_t &tc = touch[t];
			memcpy(&p[3 * (size_t)t], points + 3 * (size_t)m, 12);
			memcpy(&org[3 * (size_t)t], tc.origin, 12);
			memcpy(&ang[3 * (size_t)t], tc.angles, 12);
			memcpy(&bmn[3 * (size_t)t], tc.mins, 12);
			memcpy(&bmx[3 * (size_t)t], tc.maxs, 12);
			models[t] = tc.bmodel ? CM_InlineModel(tc.modelindex) : CMGPU_BOX_MODEL_HANDLE;   // SV_ClipHandleForEntity
		}
	}
	rc = CM_PointContentsBatch(c2.data(), total, p.data(), models.data(), org.data(), ang.data(), bmn.data(), bmx.data());
	if (rc) return rc;
	for (int m = 0; m < numPoints; m++)
		for (int t = touchFirst[m]; t < touchFirst[m + 1]; t++) contents[m] |= c2[t];
	return CMGPU_OK;
}

int SV_ClipMoveToEntitiesBatch(trace_t *clipTrace, const vec3_t start, const vec3_t end,
                               const vec3_t mins, const vec3_t maxs, int contentmask,
                               int numTouch, const cm_touch_t *touch) {
	const int first[2] = {0, numTouch};
	return SV_ClipMovesToEntitiesBatch(clipTrace, 1, start, end, mins, maxs, 0, &contentmask, 0, first, touch);
}

}  // extern "C"
