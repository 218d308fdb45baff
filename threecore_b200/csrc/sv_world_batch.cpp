// sv_world_batch.cpp -- host side of include/sv_world_batch.h: the reference's entity linking
// (code/server/sv_world.c:60-396) on an index-based sector tree, and the batch calls that hand the
// linked entities and the moves to the GPU (cmgpu_sv_*, cm_gpu.cu).  No tracing happens here.
#include "../../include/sv_world_batch.h"

#include <math.h>
#include <string.h>

#include <vector>

namespace {

constexpr int kAreaDepth = 4;            // AREA_DEPTH, sv_world.c:68
constexpr int kAreaNodes = 64;           // AREA_NODES

struct Sector {                          // worldSector_t, sv_world.c:60-66, with indices for pointers
	int axis;                            // -1 = leaf
	float dist;
	int children[2];
	int entities;                        // head of the chain (entity number), -1 = empty
};

Sector g_sectors[kAreaNodes];
int g_numSectors = 0;
int g_sectorOf[MAX_GENTITIES];           // svEntity_t.worldSector, -1 = not linked
int g_nextInSector[MAX_GENTITIES];       // svEntity_t.nextEntityInWorldSector

char *g_gents = nullptr;
int g_numGents = 0, g_gentSize = 0;
unsigned long long g_lastCandidates = 0;

sv_entity_t *gent(int num) { return reinterpret_cast<sv_entity_t *>(g_gents + (size_t)g_gentSize * (size_t)num); }

// SV_CreateworldSector, sv_world.c:81-113
int createSector(int depth, const float *mins, const float *maxs) {
	const int me = g_numSectors++;
	Sector &a = g_sectors[me];
	a.entities = -1;
	if (depth == kAreaDepth) {
		a.axis = -1;
		a.children[0] = a.children[1] = -1;
		return me;
	}
	const float sx = maxs[0] - mins[0], sy = maxs[1] - mins[1];
	a.axis = sx > sy ? 0 : 1;
	a.dist = (float)(0.5 * (double)(maxs[a.axis] + mins[a.axis]));
	float mins1[3], maxs1[3], mins2[3], maxs2[3];
	memcpy(mins1, mins, 12); memcpy(mins2, mins, 12); memcpy(maxs1, maxs, 12); memcpy(maxs2, maxs, 12);
	maxs1[a.axis] = mins2[a.axis] = a.dist;
	const int c0 = createSector(depth + 1, mins2, maxs2);
	const int c1 = createSector(depth + 1, mins1, maxs1);
	g_sectors[me].children[0] = c0;
	g_sectors[me].children[1] = c1;
	return me;
}

// RadiusFromBounds, q_math.c:766-778: fabs in double rounded back to float, float length
float radiusFromBounds(const float *mins, const float *maxs) {
	float corner[3];
	for (int i = 0; i < 3; i++) {
		const float a = (float)fabs((double)mins[i]), b = (float)fabs((double)maxs[i]);
		corner[i] = a > b ? a : b;
	}
	return sqrtf(corner[0] * corner[0] + corner[1] * corner[1] + corner[2] * corner[2]);
}

struct AreaParms { const float *mins, *maxs; int *list; int count, maxcount; };

// SV_AreaEntities_r, sv_world.c:344-377
void areaEntities_r(int node, AreaParms &ap) {
	const Sector &s = g_sectors[node];
	for (int e = s.entities; e >= 0; e = g_nextInSector[e]) {
		const sv_entity_t *g = gent(e);
		if (g->absmin[0] > ap.maxs[0] || g->absmin[1] > ap.maxs[1] || g->absmin[2] > ap.maxs[2] ||
		    g->absmax[0] < ap.mins[0] || g->absmax[1] < ap.mins[1] || g->absmax[2] < ap.mins[2]) continue;
		if (ap.count == ap.maxcount) return;             // "SV_AreaEntities: MAXCOUNT"
		ap.list[ap.count++] = e;
	}
	if (s.axis == -1) return;
	if (ap.maxs[s.axis] > s.dist) areaEntities_r(s.children[0], ap);
	if (ap.mins[s.axis] < s.dist) areaEntities_r(s.children[1], ap);
}

void allLinked_r(int node, std::vector<int> &out) {
	const Sector &s = g_sectors[node];
	for (int e = s.entities; e >= 0; e = g_nextInSector[e]) out.push_back(e);
	if (s.axis == -1) return;
	allLinked_r(s.children[0], out);
	allLinked_r(s.children[1], out);
}

// the linked entities in SV_AreaEntities order, and every gentity's owner, to the GPU
int snapshot() {
	cmgpu_ctx *ctx = CM_Context();
	if (!ctx) return CMGPU_ERR_NO_MAP;
	std::vector<int> order;
	if (g_numSectors) allLinked_r(0, order);
	std::vector<cmgpu_sv_entity> ents(order.size());
	for (size_t j = 0; j < order.size(); j++) {
		const sv_entity_t *g = gent(order[j]);
		cmgpu_sv_entity &e = ents[j];
		e.number = g->number; e.ownerNum = g->ownerNum; e.contents = g->contents;
		e.bmodel = g->bmodel; e.modelindex = g->modelindex;
		memcpy(e.mins, g->mins, 12); memcpy(e.maxs, g->maxs, 12);
		memcpy(e.absmin, g->absmin, 12); memcpy(e.absmax, g->absmax, 12);
		memcpy(e.origin, g->currentOrigin, 12); memcpy(e.angles, g->currentAngles, 12);
	}
	std::vector<int32_t> owners((size_t)g_numGents);
	for (int i = 0; i < g_numGents; i++) owners[i] = gent(i)->ownerNum;
	return cmgpu_sv_set_entities(ctx, (int)ents.size(), ents.data(), owners.data(), g_numGents);
}

}  // namespace

extern "C" {

void SV_LocateGameData(sv_entity_t *gEnts, int numGEntities, int sizeofGEntity) {
	g_gents = reinterpret_cast<char *>(gEnts);
	g_numGents = numGEntities < 0 ? 0 : (numGEntities > MAX_GENTITIES ? MAX_GENTITIES : numGEntities);
	g_gentSize = sizeofGEntity;
}

sv_entity_t *SV_GentityNum(int num) { return gent(num); }

void SV_ClearWorldForBounds(const vec3_t mins, const vec3_t maxs) {
	memset(g_sectors, 0, sizeof(g_sectors));
	g_numSectors = 0;
	for (int i = 0; i < MAX_GENTITIES; i++) g_sectorOf[i] = g_nextInSector[i] = -1;
	createSector(0, mins, maxs);
}

void SV_ClearWorld(void) {
	float mins[3], maxs[3];
	CM_ModelBounds(CM_InlineModel(0), mins, maxs);       // world map bounds, sv_world.c:128-130
	SV_ClearWorldForBounds(mins, maxs);
}

void SV_UnlinkEntity(sv_entity_t *gEnt) {
	const int num = gEnt->number;
	gEnt->linked = 0;
	if (num < 0 || num >= MAX_GENTITIES) return;
	const int ws = g_sectorOf[num];
	if (ws < 0) return;                                  // not linked in anywhere
	g_sectorOf[num] = -1;
	if (g_sectors[ws].entities == num) {
		g_sectors[ws].entities = g_nextInSector[num];
		return;
	}
	for (int scan = g_sectors[ws].entities; scan >= 0; scan = g_nextInSector[scan]) {
		if (g_nextInSector[scan] == num) {
			g_nextInSector[scan] = g_nextInSector[num];
			return;
		}
	}
}

void SV_LinkEntity(sv_entity_t *gEnt) {
	const int num = gEnt->number;
	if (num < 0 || num >= MAX_GENTITIES || !g_numSectors) return;
	if (g_sectorOf[num] >= 0) SV_UnlinkEntity(gEnt);     // unlink from old position
	const float *origin = gEnt->currentOrigin, *angles = gEnt->currentAngles;
	// set the abs box (sv_world.c:243-264)
	if (angles[0] || angles[1] || angles[2]) {
		const float max = radiusFromBounds(gEnt->mins, gEnt->maxs);      // expand for rotation
		for (int i = 0; i < 3; i++) {
			gEnt->absmin[i] = origin[i] - max;
			gEnt->absmax[i] = origin[i] + max;
		}
	} else {
		for (int i = 0; i < 3; i++) {
			gEnt->absmin[i] = origin[i] + gEnt->mins[i];
			gEnt->absmax[i] = origin[i] + gEnt->maxs[i];
		}
	}
	// movement is clipped an epsilon away from an actual edge: check even when the boxes don't quite touch
	for (int i = 0; i < 3; i++) {
		gEnt->absmin[i] -= 1;
		gEnt->absmax[i] += 1;
	}
	gEnt->linkcount++;
	// the first sector node that the box crosses (sv_world.c:308-320)
	int node = 0;
	while (g_sectors[node].axis != -1) {
		const Sector &s = g_sectors[node];
		if (gEnt->absmin[s.axis] > s.dist) node = s.children[0];
		else if (gEnt->absmax[s.axis] < s.dist) node = s.children[1];
		else break;
	}
	g_sectorOf[num] = node;
	g_nextInSector[num] = g_sectors[node].entities;      // link it in at the head
	g_sectors[node].entities = num;
	gEnt->linked = 1;
}

int SV_AreaEntities(const vec3_t mins, const vec3_t maxs, int *entityList, int maxcount) {
	AreaParms ap{mins, maxs, entityList, 0, maxcount};
	if (g_numSectors) areaEntities_r(0, ap);
	return ap.count;
}

int SV_TraceBatch(trace_t *results, int n, const float *starts, const float *mins, const float *maxs, int hull_stride,
                  const float *ends, const int *passEntityNums, int pass_stride, const int *contentmasks, int mask_stride) {
	g_lastCandidates = 0;
	if (n <= 0) return CMGPU_OK;
	cmgpu_ctx *ctx = CM_Context();
	if (!ctx || !CM_NumInlineModels()) {                 // no map: SV_Trace would return CM_BoxTrace's zeroed trace
		for (int i = 0; i < n; i++) {
			memset(&results[i], 0, sizeof(trace_t));
			results[i].fraction = 1.0f;
			results[i].entityNum = ENTITYNUM_NONE;
		}
		return CMGPU_OK;
	}
	int rc = snapshot();
	if (rc) return rc;
	rc = cmgpu_sv_trace_batch(ctx, n, starts, ends, mins, maxs, hull_stride, passEntityNums, pass_stride, contentmasks, mask_stride, results);
	g_lastCandidates = cmgpu_sv_last_candidates(ctx);
	return rc;
}

int SV_PointContentsWorldBatch(int *contents, int n, const float *points, const int *passEntityNums, int pass_stride) {
	g_lastCandidates = 0;
	if (n <= 0) return CMGPU_OK;
	cmgpu_ctx *ctx = CM_Context();
	if (!ctx || !CM_NumInlineModels()) {
		memset(contents, 0, sizeof(int) * (size_t)n);
		return CMGPU_OK;
	}
	int rc = snapshot();
	if (rc) return rc;
	rc = cmgpu_sv_point_contents_batch(ctx, n, points, passEntityNums, pass_stride, contents);
	g_lastCandidates = cmgpu_sv_last_candidates(ctx);
	return rc;
}

unsigned long long SV_LastBatchCandidates(void) { return g_lastCandidates; }

}  // extern "C"
