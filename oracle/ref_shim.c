/*
 * oracle/ref_shim.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Harness that lets the UNMODIFIED reference clip-map sources
 * (/root/reference/code/qcommon/cm_trace.c cm_test.c cm_patch.c cm_polylib.c
 * cm_load.c q_math.c q_shared.c) be driven as a shared library,
 * oracle/_ref/libcmref.so.  The reference sources are compiled where they lie
 * (see oracle/Makefile); nothing of them is copied into this repository.
 *
 * This file supplies
 *   (1) the ten engine symbols the reference objects leave undefined
 *       (Com_Error, Com_Printf, Com_DPrintf, Cvar_Get, Cmd_AddCommand,
 *       FS_ReadFile, FS_FreeFile, Hunk_Alloc, Z_Malloc, Z_Free) with the
 *       behaviour the CM layer relies on (zero-filled, 64-byte aligned,
 *       contiguous hunk: qcommon/common.c:1335-1348), and
 *   (2) a batch API (cmref_*) around the reference's own public entry points
 *       (qcommon/cm_public.h:26-69) plus dump helpers that expose the LOADED
 *       clipMap_t as index-based arrays, so the product's independent loader
 *       can be checked against it.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's reference / cpu_baseline
 * legs may load the resulting library.
 */
#define _GNU_SOURCE
#include <setjmp.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sched.h>
#include <time.h>
#include <unistd.h>
#include <sys/mman.h>
#include <sys/wait.h>

#include "cm_local.h"
#include "cm_patch.h"

#define API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ */
/* engine stubs                                                        */
/* ------------------------------------------------------------------ */

static jmp_buf  s_abort;
static int      s_abortArmed;
static char     s_errmsg[1024];
static int      s_verbose;

void QDECL Com_Error(errorParm_t level, const char *fmt, ...) {
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(s_errmsg, sizeof(s_errmsg), fmt, ap);
	va_end(ap);
	(void)level;
	if (s_abortArmed) {
		longjmp(s_abort, 1);
	}
	fprintf(stderr, "cmref: Com_Error outside guarded call: %s\n", s_errmsg);
	abort();
}

void QDECL Com_Printf(const char *fmt, ...) {
	if (s_verbose) {
		va_list ap;
		va_start(ap, fmt);
		vfprintf(stderr, fmt, ap);
		va_end(ap);
	}
}

void QDECL Com_DPrintf(const char *fmt, ...) {
	if (s_verbose > 1) {
		va_list ap;
		va_start(ap, fmt);
		vfprintf(stderr, fmt, ap);
		va_end(ap);
	}
}

/* the CM layer only ever reads ->integer of these */
static cvar_t s_cvNoAreas, s_cvNoCurves, s_cvPlayerCurveClip, s_cvDebugSurface, s_cvOther;

cvar_t *Cvar_Get(const char *name, const char *value, int flags) {
	(void)value; (void)flags;
	if (!strcmp(name, "cm_noAreas")) return &s_cvNoAreas;
	if (!strcmp(name, "cm_noCurves")) return &s_cvNoCurves;
	if (!strcmp(name, "cm_playerCurveClip")) return &s_cvPlayerCurveClip;
	if (!strcmp(name, "r_debugSurfaceUpdate")) return &s_cvDebugSurface;
	return &s_cvOther;
}

void Cmd_AddCommand(const char *cmd_name, xcommand_t function) {
	(void)cmd_name; (void)function;
}

static const void *s_fileData;
static int         s_fileLen;

int FS_ReadFile(const char *qpath, void **buffer) {
	(void)qpath;
	if (!s_fileData) {
		if (buffer) *buffer = NULL;
		return -1;
	}
	if (buffer) {
		/* the engine hands back a private copy; so do we */
		void *p = malloc((size_t)s_fileLen + 1);
		memcpy(p, s_fileData, (size_t)s_fileLen);
		((char *)p)[s_fileLen] = 0;
		*buffer = p;
	}
	return s_fileLen;
}

void FS_FreeFile(void *buffer) { free(buffer); }

/* one contiguous, lazily committed arena */
#define HUNK_BYTES ((size_t)3 << 30)
static unsigned char *s_hunk;
static size_t         s_hunkUsed;

void *Hunk_Alloc(int size) {
	size_t sz = ((size_t)size + 63) & ~(size_t)63;
	if (!s_hunk) {
		s_hunk = mmap(NULL, HUNK_BYTES, PROT_READ | PROT_WRITE,
		              MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
		if (s_hunk == MAP_FAILED) {
			s_hunk = NULL;
			Com_Error(ERR_FATAL, "cmref: hunk mmap failed");
		}
	}
	if (s_hunkUsed + sz > HUNK_BYTES) {
		Com_Error(ERR_DROP, "cmref: hunk exhausted");
	}
	void *p = s_hunk + s_hunkUsed;
	s_hunkUsed += sz;
	memset(p, 0, sz);
	return p;
}

#ifdef Z_Malloc
#undef Z_Malloc
#endif
void *Z_Malloc(int size) { return calloc(1, (size_t)size ? (size_t)size : 1); }
void *Z_MallocDebug(int size, char *label, char *file, int line) {
	(void)label; (void)file; (void)line;
	return calloc(1, (size_t)size ? (size_t)size : 1);
}
void Z_Free(void *ptr) { free(ptr); }

/* ------------------------------------------------------------------ */
/* guarded-call helper                                                  */
/* ------------------------------------------------------------------ */

#define GUARD_BEGIN()  do { s_abortArmed = 1; if (setjmp(s_abort)) { s_abortArmed = 0; return -1; } } while (0)
#define GUARD_END()    do { s_abortArmed = 0; } while (0)

API const char *cmref_last_error(void) { return s_errmsg; }
API void cmref_set_verbose(int v) { s_verbose = v; }

API void cmref_set_cvars(int noCurves, int playerCurveClip) {
	s_cvNoCurves.integer = noCurves;
	s_cvPlayerCurveClip.integer = playerCurveClip;
}

/* ------------------------------------------------------------------ */
/* map load                                                            */
/* ------------------------------------------------------------------ */

API int cmref_load(const void *bsp, int len) {
	static int serial;
	char name[64];
	int checksum = 0;

	GUARD_BEGIN();
	CM_ClearMap();
	s_hunkUsed = 0;
	s_cvPlayerCurveClip.integer = 1;
	s_cvNoCurves.integer = 0;
	s_cvNoAreas.integer = 0;
	s_cvDebugSurface.integer = 0;
	s_fileData = bsp;
	s_fileLen = len;
	snprintf(name, sizeof(name), "maps/synthetic%d.bsp", serial++);
	CM_LoadMap(name, qfalse, &checksum);
	s_fileData = NULL;
	GUARD_END();
	return cmWorlds[0].numNodes ? 0 : -2;
}

API int cmref_num_inline_models(void) { return CM_NumInlineModels(); }

/* ------------------------------------------------------------------ */
/* batch drivers around the reference's public entry points            */
/* ------------------------------------------------------------------ */

/* hull_stride / mask_stride: 0 = one shared hull / mask, 1 = per trace. */
API int cmref_box_trace_batch(int n, const float *starts, const float *ends,
                              const float *mins, const float *maxs, int hull_stride,
                              const int *models, const int *masks, int mask_stride,
                              const float *boxmins, const float *boxmaxs,
                              trace_t *out) {
	GUARD_BEGIN();
	for (int i = 0; i < n; i++) {
		const float *mn = mins ? mins + 3 * (size_t)i * hull_stride : NULL;
		const float *mx = maxs ? maxs + 3 * (size_t)i * hull_stride : NULL;
		int h = models ? models[i] : 0;
		if (h == BOX_MODEL_HANDLE && boxmins) {
			h = CM_TempBoxModel(boxmins + 3 * (size_t)i, boxmaxs + 3 * (size_t)i);
		}
		CM_BoxTrace(&out[i], starts + 3 * (size_t)i, ends + 3 * (size_t)i, mn, mx, h, masks[(size_t)i * mask_stride]);
	}
	GUARD_END();
	return 0;
}

/* models[i] == BOX_MODEL_HANDLE (1023): CM_TempBoxModel(boxmins[i], boxmaxs[i]) first */
API int cmref_transformed_box_trace_batch(int n, const float *starts, const float *ends,
                                          const float *mins, const float *maxs, int hull_stride,
                                          const int *models, const int *masks, int mask_stride,
                                          const float *origins, const float *angles,
                                          const float *boxmins, const float *boxmaxs,
                                          trace_t *out) {
	GUARD_BEGIN();
	for (int i = 0; i < n; i++) {
		const float *mn = mins ? mins + 3 * (size_t)i * hull_stride : NULL;
		const float *mx = maxs ? maxs + 3 * (size_t)i * hull_stride : NULL;
		int h = models[i];
		if (h == BOX_MODEL_HANDLE) {
			h = CM_TempBoxModel(boxmins + 3 * (size_t)i, boxmaxs + 3 * (size_t)i);
		}
		CM_TransformedBoxTrace(&out[i], starts + 3 * (size_t)i, ends + 3 * (size_t)i, mn, mx, h,
		                       masks[(size_t)i * mask_stride],
		                       origins + 3 * (size_t)i, angles + 3 * (size_t)i);
	}
	GUARD_END();
	return 0;
}

API int cmref_point_contents_batch(int n, const float *pts, const int *models,
                                   const float *boxmins, const float *boxmaxs, int *out) {
	GUARD_BEGIN();
	for (int i = 0; i < n; i++) {
		int h = models ? models[i] : 0;
		if (h == BOX_MODEL_HANDLE) {
			h = CM_TempBoxModel(boxmins + 3 * (size_t)i, boxmaxs + 3 * (size_t)i);
		}
		out[i] = CM_PointContents(pts + 3 * (size_t)i, h);
	}
	GUARD_END();
	return 0;
}

API int cmref_transformed_point_contents_batch(int n, const float *pts, const int *models,
                                               const float *origins, const float *angles,
                                               const float *boxmins, const float *boxmaxs, int *out) {
	GUARD_BEGIN();
	for (int i = 0; i < n; i++) {
		int h = models[i];
		if (h == BOX_MODEL_HANDLE) {
			h = CM_TempBoxModel(boxmins + 3 * (size_t)i, boxmaxs + 3 * (size_t)i);
		}
		out[i] = CM_TransformedPointContents(pts + 3 * (size_t)i, h,
		                                     origins + 3 * (size_t)i, angles + 3 * (size_t)i);
	}
	GUARD_END();
	return 0;
}

API int cmref_point_leafnum_batch(int n, const float *pts, int *out) {
	GUARD_BEGIN();
	for (int i = 0; i < n; i++) {
		out[i] = CM_PointLeafnum(pts + 3 * (size_t)i);
	}
	GUARD_END();
	return 0;
}

/* CM_BoxLeafnums: list is [n][listsize]; counts[i] and lastLeaf[i] per box */
API int cmref_box_leafnums_batch(int n, const float *mins, const float *maxs, int listsize,
                                 int *lists, int *counts, int *lastLeafs) {
	GUARD_BEGIN();
	for (int i = 0; i < n; i++) {
		counts[i] = CM_BoxLeafnums(mins + 3 * (size_t)i, maxs + 3 * (size_t)i,
		                           lists + (size_t)i * listsize, listsize, &lastLeafs[i]);
	}
	GUARD_END();
	return 0;
}

/* ---- clusters, areas, visibility (cm_load.c:838-860, cm_test.c:306-470) ---- */
API int cmref_leaf_cluster_area(int n, const int *leafs, int *cluster, int *area) {
	GUARD_BEGIN();
	for (int i = 0; i < n; i++) {
		cluster[i] = CM_LeafCluster(leafs[i]);
		area[i] = CM_LeafArea(leafs[i]);
	}
	GUARD_END();
	return 0;
}

API int cmref_num_clusters(void) { return CM_NumClusters(); }

/* the row CM_ClusterPVS(cluster) points at, (numClusters+7)/8 bytes of it (the engine's users index it by cluster) */
API int cmref_cluster_pvs(int cluster, int nbytes, unsigned char *out) {
	GUARD_BEGIN();
	memcpy(out, CM_ClusterPVS(cluster), (size_t)nbytes);
	GUARD_END();
	return 0;
}

API int cmref_adjust_area_portal(int area1, int area2, int open) {
	GUARD_BEGIN();
	CM_AdjustAreaPortalState(area1, area2, open ? qtrue : qfalse);
	GUARD_END();
	return 0;
}

API int cmref_areas_connected(int n, const int *a1, const int *a2, int *out) {
	GUARD_BEGIN();
	for (int i = 0; i < n; i++) out[i] = CM_AreasConnected(a1[i], a2[i]);
	GUARD_END();
	return 0;
}

/* CM_WriteAreaBits into a zeroed buffer; returns the byte count in *bytes */
API int cmref_write_area_bits(int area, unsigned char *buffer, int *bytes) {
	GUARD_BEGIN();
	*bytes = CM_WriteAreaBits(buffer, area);
	GUARD_END();
	return 0;
}

API void cmref_angle_vectors(const float *angles, float *fwd, float *right, float *up) {
	AngleVectors(angles, fwd, right, up);
}

/* ------------------------------------------------------------------ */
/* multi-process throughput run (the CPU baseline)                     */
/* ------------------------------------------------------------------ */

static double now_s(void) {
	struct timespec ts;
	clock_gettime(CLOCK_MONOTONIC, &ts);
	return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

/*
 * The CM layer is not re-entrant (cm.checkcount, per-brush stamps), so the
 * all-core figure forks one worker per requested core after the map is
 * loaded; each traces a contiguous slice, pinned to one CPU.  Returns the
 * wall time of the slowest worker in *seconds (plus per-worker times when
 * worker_seconds != NULL).  Results, when out != NULL, must point to a
 * MAP_SHARED region (cmref_shared_alloc) to be visible to the parent.
 */
API void *cmref_shared_alloc(size_t bytes) {
	void *p = mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
	return p == MAP_FAILED ? NULL : p;
}
API void cmref_shared_free(void *p, size_t bytes) { if (p) munmap(p, bytes); }

API int cmref_box_trace_bench(int n, const float *starts, const float *ends,
                              const float *mins, const float *maxs, int mask,
                              int nworkers, int repeats, trace_t *out_shared,
                              double *seconds, double *worker_seconds) {
	if (nworkers < 1) nworkers = 1;
	if (nworkers > 1024) nworkers = 1024;
	double *wt = cmref_shared_alloc(sizeof(double) * (size_t)nworkers);
	if (!wt) return -1;
	long ncpu = sysconf(_SC_NPROCESSORS_ONLN);
	cpu_set_t allowed;
	int have_aff = sched_getaffinity(0, sizeof(allowed), &allowed) == 0;
	int cpus[1024], ncpus = 0;
	if (have_aff) {
		for (int c = 0; c < CPU_SETSIZE && ncpus < 1024; c++)
			if (CPU_ISSET(c, &allowed)) cpus[ncpus++] = c;
	}
	if (ncpus == 0) {
		for (int c = 0; c < ncpu && c < 1024; c++) cpus[ncpus++] = c;
	}
	pid_t pids[1024];
	for (int w = 0; w < nworkers; w++) {
		wt[w] = -1.0;
		pid_t pid = fork();
		if (pid < 0) return -1;
		if (pid == 0) {
			cpu_set_t set;
			CPU_ZERO(&set);
			CPU_SET(cpus[w % ncpus], &set);
			sched_setaffinity(0, sizeof(set), &set);
			size_t lo = (size_t)n * (size_t)w / (size_t)nworkers;
			size_t hi = (size_t)n * (size_t)(w + 1) / (size_t)nworkers;
			trace_t tr;
			double best = 1e30;
			for (int r = 0; r < repeats; r++) {
				double t0 = now_s();
				for (size_t i = lo; i < hi; i++) {
					trace_t *dst = out_shared ? &out_shared[i] : &tr;
					CM_BoxTrace(dst, starts + 3 * i, ends + 3 * i, mins, maxs, 0, mask);
				}
				double t1 = now_s();
				if (t1 - t0 < best) best = t1 - t0;
			}
			wt[w] = best;
			_exit(0);
		}
		pids[w] = pid;
	}
	int rc = 0;
	for (int w = 0; w < nworkers; w++) {
		int st = 0;
		waitpid(pids[w], &st, 0);
		if (!WIFEXITED(st) || WEXITSTATUS(st) != 0) rc = -1;
	}
	double worst = 0.0;
	for (int w = 0; w < nworkers; w++) {
		if (wt[w] < 0) rc = -1;
		if (wt[w] > worst) worst = wt[w];
		if (worker_seconds) worker_seconds[w] = wt[w];
	}
	*seconds = worst;
	cmref_shared_free(wt, sizeof(double) * (size_t)nworkers);
	return rc;
}

/* ------------------------------------------------------------------ */
/* dump of the loaded clipMap_t (pointer graph -> indices)             */
/* ------------------------------------------------------------------ */

/* counts[] layout */
enum {
	CMREF_N_PLANES, CMREF_N_NODES, CMREF_N_LEAFS, CMREF_N_LEAFBRUSHES, CMREF_N_LEAFSURFACES,
	CMREF_N_BRUSHES, CMREF_N_BRUSHSIDES, CMREF_N_MODELS, CMREF_N_SURFACES, CMREF_N_SHADERS,
	CMREF_N_PATCHES, CMREF_N_PATCHPLANES, CMREF_N_FACETS, CMREF_N_BORDERS, CMREF_N_COUNTS
};

API void cmref_counts(int *counts) {
	clipMap_t *c = &cmWorlds[0];
	memset(counts, 0, sizeof(int) * CMREF_N_COUNTS);
	counts[CMREF_N_PLANES] = c->numPlanes + 12;          /* + box hull planes, cm_load.c:45-50 */
	counts[CMREF_N_NODES] = c->numNodes;
	counts[CMREF_N_LEAFS] = c->numLeafs;
	counts[CMREF_N_LEAFBRUSHES] = c->numLeafBrushes + 1; /* + box brush ref */
	counts[CMREF_N_LEAFSURFACES] = c->numLeafSurfaces;
	counts[CMREF_N_BRUSHES] = c->numBrushes + 1;         /* + box brush */
	counts[CMREF_N_BRUSHSIDES] = c->numBrushSides + 6;
	counts[CMREF_N_MODELS] = c->numSubModels;
	counts[CMREF_N_SURFACES] = c->numSurfaces;
	counts[CMREF_N_SHADERS] = c->numShaders;
	for (int i = 0; i < c->numSurfaces; i++) {
		cPatch_t *p = c->surfaces[i];
		if (!p) continue;
		counts[CMREF_N_PATCHES]++;
		counts[CMREF_N_PATCHPLANES] += p->pc->numPlanes;
		counts[CMREF_N_FACETS] += p->pc->numFacets;
		for (int f = 0; f < p->pc->numFacets; f++) counts[CMREF_N_BORDERS] += p->pc->facets[f].numBorders;
	}
}

/* planes: float[4] normal+dist, plus type/signbits bytes */
API void cmref_dump_planes(float *nd, unsigned char *type, unsigned char *signbits) {
	clipMap_t *c = &cmWorlds[0];
	int n = c->numPlanes + 12;
	for (int i = 0; i < n; i++) {
		nd[4 * i + 0] = c->planes[i].normal[0];
		nd[4 * i + 1] = c->planes[i].normal[1];
		nd[4 * i + 2] = c->planes[i].normal[2];
		nd[4 * i + 3] = c->planes[i].dist;
		type[i] = c->planes[i].type;
		signbits[i] = c->planes[i].signbits;
	}
}

/* nodes: int[3] = plane index, child0, child1 */
API void cmref_dump_nodes(int *out) {
	clipMap_t *c = &cmWorlds[0];
	for (int i = 0; i < c->numNodes; i++) {
		out[3 * i + 0] = (int)(c->nodes[i].plane - c->planes);
		out[3 * i + 1] = c->nodes[i].children[0];
		out[3 * i + 2] = c->nodes[i].children[1];
	}
}

/* leafs: int[6] = cluster, area, firstLeafBrush, numLeafBrushes, firstLeafSurface, numLeafSurfaces */
API void cmref_dump_leafs(int *out) {
	clipMap_t *c = &cmWorlds[0];
	memcpy(out, c->leafs, sizeof(int) * 6 * (size_t)c->numLeafs);
}

API void cmref_dump_leafbrushes(int *out) {
	clipMap_t *c = &cmWorlds[0];
	memcpy(out, c->leafbrushes, sizeof(int) * (size_t)(c->numLeafBrushes + 1));
}

API void cmref_dump_leafsurfaces(int *out) {
	clipMap_t *c = &cmWorlds[0];
	memcpy(out, c->leafsurfaces, sizeof(int) * (size_t)c->numLeafSurfaces);
}

/* brushes: int[4] = shaderNum, contents, firstSide, numsides ; float[6] bounds */
API void cmref_dump_brushes(int *meta, float *bounds) {
	clipMap_t *c = &cmWorlds[0];
	int n = c->numBrushes + 1;
	for (int i = 0; i < n; i++) {
		cbrush_t *b = &c->brushes[i];
		meta[4 * i + 0] = b->shaderNum;
		meta[4 * i + 1] = b->contents;
		meta[4 * i + 2] = b->sides ? (int)(b->sides - c->brushsides) : 0;
		meta[4 * i + 3] = b->numsides;
		memcpy(bounds + 6 * i, b->bounds, sizeof(float) * 6);
	}
}

/* brushsides: int[3] = plane index, surfaceFlags, shaderNum */
API void cmref_dump_brushsides(int *out) {
	clipMap_t *c = &cmWorlds[0];
	int n = c->numBrushSides + 6;
	for (int i = 0; i < n; i++) {
		out[3 * i + 0] = c->brushsides[i].plane ? (int)(c->brushsides[i].plane - c->planes) : 0;
		out[3 * i + 1] = c->brushsides[i].surfaceFlags;
		out[3 * i + 2] = c->brushsides[i].shaderNum;
	}
}

/*
 * models: float[6] mins,maxs ; int[2] numLeafBrushes, numLeafSurfaces per model.
 * Submodel index lists live on the hunk outside leafbrushes[] (cm_load.c:170-182);
 * brushlist/surflist receive them concatenated in model order.
 */
API void cmref_dump_models(float *bounds, int *nums) {
	clipMap_t *c = &cmWorlds[0];
	for (int i = 0; i < c->numSubModels; i++) {
		memcpy(bounds + 6 * i, c->cmodels[i].mins, sizeof(float) * 3);
		memcpy(bounds + 6 * i + 3, c->cmodels[i].maxs, sizeof(float) * 3);
		nums[2 * i + 0] = i ? c->cmodels[i].leaf.numLeafBrushes : 0;
		nums[2 * i + 1] = i ? c->cmodels[i].leaf.numLeafSurfaces : 0;
	}
}

API void cmref_dump_model_lists(int *brushlist, int *surflist) {
	clipMap_t *c = &cmWorlds[0];
	for (int i = 1; i < c->numSubModels; i++) {
		cLeaf_t *l = &c->cmodels[i].leaf;
		for (int k = 0; k < l->numLeafBrushes; k++) *brushlist++ = c->leafbrushes[l->firstLeafBrush + k];
		for (int k = 0; k < l->numLeafSurfaces; k++) *surflist++ = c->leafsurfaces[l->firstLeafSurface + k];
	}
}

/*
 * patches, in surface order.
 *   surf2patch[numSurfaces]  : patch ordinal or -1
 *   pmeta[patch][4]          : surfaceFlags, contents, numPlanes, numFacets
 *   pbounds[patch][6]
 *   planes[sum numPlanes][4] + psign[sum numPlanes]
 *   fmeta[sum numFacets][2]  : surfacePlane, numBorders
 *   borders[sum borders][3]  : plane, inward, noAdjust
 */
API void cmref_dump_patches(int *surf2patch, int *pmeta, float *pbounds, float *planes, int *psign,
                            int *fmeta, int *borders) {
	clipMap_t *c = &cmWorlds[0];
	int np = 0;
	for (int i = 0; i < c->numSurfaces; i++) {
		cPatch_t *p = c->surfaces[i];
		if (!p) { surf2patch[i] = -1; continue; }
		surf2patch[i] = np;
		pmeta[4 * np + 0] = p->surfaceFlags;
		pmeta[4 * np + 1] = p->contents;
		pmeta[4 * np + 2] = p->pc->numPlanes;
		pmeta[4 * np + 3] = p->pc->numFacets;
		memcpy(pbounds + 6 * np, p->pc->bounds, sizeof(float) * 6);
		for (int k = 0; k < p->pc->numPlanes; k++) {
			memcpy(planes, p->pc->planes[k].plane, sizeof(float) * 4);
			planes += 4;
			*psign++ = p->pc->planes[k].signbits;
		}
		for (int f = 0; f < p->pc->numFacets; f++) {
			facet_t *fc = &p->pc->facets[f];
			*fmeta++ = fc->surfacePlane;
			*fmeta++ = fc->numBorders;
			for (int j = 0; j < fc->numBorders; j++) {
				*borders++ = fc->borderPlanes[j];
				*borders++ = fc->borderInward[j];
				*borders++ = fc->borderNoAdjust[j];
			}
		}
		np++;
	}
}

API int cmref_sizeof_trace(void) { return (int)sizeof(trace_t); }
