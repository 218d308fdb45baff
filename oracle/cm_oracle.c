/*
 * oracle/cm_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the reference clip-map trace.  It is a checker only:
 * tests/ compare the CUDA engine with it, __graft_entry__.smoke() checks one
 * launch against it, and bench.py may time it as the "port" CPU baseline.
 * Nothing under threecore_b200/ links, imports or calls it.
 *
 * Parity pinning: the reference ships no golden vectors, so
 * this restatement is pinned against the UNMODIFIED reference sources compiled
 * into oracle/_ref/libcmref.so (oracle/Makefile) -- tests/test_oracle_vs_ref.py
 * runs both on the same generated maps/rays and requires byte-identical
 * trace_t records -- and against the committed fixtures in tests/golden/
 * that were produced by that library (tests/golden/make_golden.py).
 *
 * Written from the algorithm, not from the text, of
 *   code/qcommon/cm_trace.c   (sweep, position test, transformed trace)
 *   code/qcommon/cm_test.c    (point leaf, box leafs, point contents, bounds)
 *   code/qcommon/cm_patch.c   (run-time facet tests, :1220-1531)
 *   code/qcommon/q_math.c     (BoxOnPlaneSide :724, AngleVectors :999)
 * with iteration + explicit stacks in place of recursion, index arrays in
 * place of pointers, no global state and no check-count stamps (a repeated
 * brush/patch test cannot change the result: every update is guarded by a
 * strict comparison with the running fraction).
 *
 * Arithmetic follows the reference's evaluation widths operation by operation
 * (SURVEY.md 8a'): where the C source mixes float and double the same mix is
 * used here.  Build with -ffp-contract=off and without -ffast-math.
 */
#include "cm_oracle.h"

#include <math.h>
#include <string.h>

#define CLIP_EPS        0.125      /* SURFACE_CLIP_EPSILON, cm_local.h:177 (a double constant) */
#define BOUNDS_EPS      0.25f      /* BOUNDS_CLIP_EPSILON, cm_test.c:474 */
#define MAX_POS_LEAFS   1024       /* MAX_POSITION_LEAFS, cm_trace.c:190 */
#define STACK_MAX 4096          /* far above any tree the tests build; the reference recurses without a limit */

typedef struct {
	float start[3], end[3];        /* symmetrised sweep endpoints */
	float size[2][3];
	float corner[8][3];            /* offsets[signbits] */
	float extents[3];
	float lo[3], hi[3];            /* swept bounds */
	int   mask;
	int   isPoint;
	const float *boxmins, *boxmaxs; /* temp-box model geometry (handle 1023) */
	cmo_trace_t tr;
	cmo_info_t  info;
} work_t;

/* ---------------------------------------------------------------------- */
/* small helpers                                                           */
/* ---------------------------------------------------------------------- */

static float dot3f(const float *a, const float *b) {
	/* DotProduct, q_shared.h:489 : ((a0*b0 + a1*b1) + a2*b2) in float */
	float s = a[0] * b[0];
	s = s + a[1] * b[1];
	s = s + a[2] * b[2];
	return s;
}

static double dot3d(const float *a, const float *b) {
	/* DotProductDP, cm_local.h:31 */
	double s = (double)a[0] * (double)b[0];
	s = s + (double)a[1] * (double)b[1];
	s = s + (double)a[2] * (double)b[2];
	return s;
}

/* plane of a brush side; the temp-box sides are rebuilt from the item's box
 * (CM_TempBoxModel, cm_load.c:927-949: +axis planes carry maxs, -axis planes -mins) */
typedef struct { float n[3]; float d; uint8_t type, signbits; } plane_t;

static void fetch_plane(const cmo_map_t *m, const work_t *w, int planeIdx, plane_t *pl) {
	const float *p = m->planes + 4 * (size_t)planeIdx;
	pl->n[0] = p[0]; pl->n[1] = p[1]; pl->n[2] = p[2];
	pl->d = p[3];
	pl->type = m->planeType[planeIdx];
	pl->signbits = m->planeSignbits[planeIdx];
	if (planeIdx >= m->boxFirstPlane && w->boxmins) {
		int k = planeIdx - m->boxFirstPlane;   /* 0..11: pairs per face i = k>>1 */
		int face = k >> 1;                     /* 0:+x(maxs) 1:-x(mins) 2:+y 3:-y 4:+z 5:-z */
		int axis = face >> 1;
		float v = (face & 1) ? w->boxmins[axis] : w->boxmaxs[axis];
		pl->d = (k & 1) ? -v : v;
	}
}

static void brush_bounds(const cmo_map_t *m, const work_t *w, int b, float *mn, float *mx) {
	if (b == m->boxBrush && w->boxmins) {
		memcpy(mn, w->boxmins, 12);
		memcpy(mx, w->boxmaxs, 12);
	} else {
		memcpy(mn, m->brushBounds + 6 * (size_t)b, 12);
		memcpy(mx, m->brushBounds + 6 * (size_t)b + 3, 12);
	}
}

/* CM_BoundsIntersect, cm_test.c:481-494 */
static int bounds_overlap(const float *lo, const float *hi, const float *lo2, const float *hi2) {
	for (int i = 0; i < 3; i++) {
		if (hi[i] < lo2[i] - BOUNDS_EPS) return 0;
	}
	for (int i = 0; i < 3; i++) {
		if (lo[i] > hi2[i] + BOUNDS_EPS) return 0;
	}
	return 1;
}

/* CM_BoundsIntersectPoint, cm_test.c:502-515 */
static int bounds_has_point(const float *lo, const float *hi, const float *p) {
	for (int i = 0; i < 3; i++) {
		if (hi[i] < p[i] - BOUNDS_EPS) return 0;
	}
	for (int i = 0; i < 3; i++) {
		if (lo[i] > p[i] + BOUNDS_EPS) return 0;
	}
	return 1;
}

/* ---------------------------------------------------------------------- */
/* sweep against one brush -- CM_TraceThroughBrush, cm_trace.c:261-363     */
/* ---------------------------------------------------------------------- */
static void sweep_brush(const cmo_map_t *m, work_t *w, int b) {
	const int *bm = m->brushes + 3 * (size_t)b;
	int nsides = bm[2];
	if (nsides == 0) return;

	float enter = -1.0f, leave = 1.0f;
	int   lead = -1;            /* side index that owns `enter` */
	plane_t leadPlane;
	int   endOutside = 0, startOutside = 0;
	memset(&leadPlane, 0, sizeof(leadPlane));

	for (int s = 0; s < nsides; s++) {
		int side = bm[1] + s;
		plane_t pl;
		fetch_plane(m, w, m->brushsides[2 * (size_t)side], &pl);
		w->info.brushSides++;

		double dist = (double)pl.d - dot3d(w->corner[pl.signbits], pl.n);
		double d1 = dot3d(w->start, pl.n) - dist;
		double d2 = dot3d(w->end, pl.n) - dist;

		if (d2 > 0) endOutside = 1;
		if (d1 > 0) startOutside = 1;

		/* wholly in front of this face: misses the brush */
		if (d1 > 0 && (d2 >= CLIP_EPS || d2 >= d1)) return;
		/* wholly behind: face irrelevant */
		if (d1 <= 0 && d2 <= 0) continue;

		w->info.crossings++;
		if (d1 > d2) {
			float f = (float)((d1 - CLIP_EPS) / (d1 - d2));
			if (f < 0) f = 0;
			if (f > enter) { enter = f; lead = side; leadPlane = pl; }
		} else {
			float f = (float)((d1 + CLIP_EPS) / (d1 - d2));
			if (f > 1) f = 1;
			if (f < leave) leave = f;
		}
	}

	if (!startOutside) {
		w->tr.startsolid = 1;
		if (!endOutside) {
			w->tr.allsolid = 1;
			w->tr.fraction = 0;
			w->tr.contents = bm[0];
			w->info.hitBrush = b;
			w->info.hitPatch = -1;
		}
		return;
	}

	if (enter < leave && enter > -1 && enter < w->tr.fraction) {
		if (enter < 0) enter = 0;
		w->tr.fraction = enter;
		if (lead >= 0) {
			memcpy(w->tr.plane_normal, leadPlane.n, 12);
			w->tr.plane_dist = leadPlane.d;
			w->tr.plane_type = leadPlane.type;
			w->tr.plane_signbits = leadPlane.signbits;
			w->tr.plane_pad[0] = w->tr.plane_pad[1] = 0;
			w->tr.surfaceFlags = m->brushsides[2 * (size_t)lead + 1];
		}
		w->tr.contents = bm[0];
		w->info.hitBrush = b;
		w->info.hitPatch = -1;
	}
}

/* ---------------------------------------------------------------------- */
/* patch facets                                                            */
/* ---------------------------------------------------------------------- */

/* CM_CheckFacetPlane, cm_patch.c:1321-1360.  returns 0 when the sweep misses */
static int clip_facet_plane(const float *pl, const float *s, const float *e,
                            float *enter, float *leave, int *hit) {
	*hit = 0;
	float d1 = dot3f(s, pl) - pl[3];
	float d2 = dot3f(e, pl) - pl[3];
	if (d1 > 0 && (d2 >= CLIP_EPS || d2 >= d1)) return 0;
	if (d1 <= 0 && d2 <= 0) return 1;
	if (d1 > d2) {
		float f = (float)(((double)d1 - CLIP_EPS) / (double)(d1 - d2));
		if (f < 0) f = 0;
		if (f > *enter) { *enter = f; *hit = 1; }
	} else {
		float f = (float)(((double)d1 + CLIP_EPS) / (double)(d1 - d2));
		if (f > 1) f = 1;
		if (f < *leave) *leave = f;
	}
	return 1;
}

/* CM_TracePointThroughPatchCollide, cm_patch.c:1220-1313 */
static void sweep_patch_point(const cmo_map_t *m, work_t *w, int p) {
	const int *pm = m->patches + 6 * (size_t)p;
	const float *planes = m->patchPlanes + 4 * (size_t)pm[2];
	const int *psign = m->patchPlaneSignbits + pm[2];
	int np = pm[3];
	/* MAX_PATCH_PLANES = 2176, cm_patch.h:24 */
	static __thread unsigned char front[4096];
	static __thread float cross[4096];
	if (!m->playerCurveClip || !w->isPoint) return;
	if (np > 4096) np = 4096;

	for (int i = 0; i < np; i++) {
		const float *pl = planes + 4 * i;
		float off = dot3f(w->corner[psign[i]], pl);
		float d1 = dot3f(w->start, pl) - pl[3] + off;
		float d2 = dot3f(w->end, pl) - pl[3] + off;
		w->info.patchPlanes++;
		front[i] = !(d1 <= 0);
		if (d1 == d2) {
			cross[i] = 99999;
		} else {
			cross[i] = d1 / (d1 - d2);
			if (cross[i] <= 0) cross[i] = 99999;
		}
	}

	for (int f = 0; f < pm[5]; f++) {
		const int *fc = m->facets + 3 * (size_t)(pm[4] + f);
		w->info.facets++;
		if (!front[fc[0]]) continue;
		float t = cross[fc[0]];
		if (t < 0) continue;
		if (t > w->tr.fraction) continue;
		int j;
		for (j = 0; j < fc[1]; j++) {
			const int *bd = m->borders + 2 * (size_t)(fc[2] + j);
			int k = bd[0];
			if (front[k] ^ bd[1]) {
				if (cross[k] > t) break;
			} else {
				if (cross[k] < t) break;
			}
		}
		if (j != fc[1]) continue;

		const float *pl = planes + 4 * fc[0];
		float off = dot3f(w->corner[psign[fc[0]]], pl);
		float d1 = dot3f(w->start, pl) - pl[3] + off;
		float d2 = dot3f(w->end, pl) - pl[3] + off;
		float fr = (float)(((double)d1 - CLIP_EPS) / (double)(d1 - d2));
		if (fr < 0) fr = 0;
		w->tr.fraction = fr;
		/* only normal+dist are written: type/signbits keep whatever an earlier
		 * brush hit left there (cm_patch.c:1309-1310) */
		memcpy(w->tr.plane_normal, pl, 12);
		w->tr.plane_dist = pl[3];
	}
}

/* CM_TraceThroughPatchCollide, cm_patch.c:1368-1463 */
static void sweep_patch(const cmo_map_t *m, work_t *w, int p) {
	const int *pm = m->patches + 6 * (size_t)p;
	const float *pb = m->patchBounds + 6 * (size_t)p;
	if (!bounds_overlap(w->lo, w->hi, pb, pb + 3)) return;
	w->info.patches++;
	if (w->isPoint) {
		sweep_patch_point(m, w, p);
		return;
	}
	const float *planes = m->patchPlanes + 4 * (size_t)pm[2];
	const int *psign = m->patchPlaneSignbits + pm[2];
	float best[4] = {0, 0, 0, 0};

	for (int f = 0; f < pm[5]; f++) {
		const int *fc = m->facets + 3 * (size_t)(pm[4] + f);
		float enter = -1.0f, leave = 1.0f;
		int hit, hitnum = -1;
		float pl[4];
		w->info.facets++;

		const float *sp = planes + 4 * fc[0];
		pl[0] = sp[0]; pl[1] = sp[1]; pl[2] = sp[2];
		pl[3] = sp[3] - dot3f(w->corner[psign[fc[0]]], pl);
		if (!clip_facet_plane(pl, w->start, w->end, &enter, &leave, &hit)) continue;
		if (hit) memcpy(best, pl, 16);

		int j;
		for (j = 0; j < fc[1]; j++) {
			const int *bd = m->borders + 2 * (size_t)(fc[2] + j);
			const float *bp = planes + 4 * bd[0];
			w->info.patchPlanes++;
			if (bd[1]) {
				pl[0] = -bp[0]; pl[1] = -bp[1]; pl[2] = -bp[2]; pl[3] = -bp[3];
			} else {
				pl[0] = bp[0]; pl[1] = bp[1]; pl[2] = bp[2]; pl[3] = bp[3];
			}
			float off = dot3f(w->corner[psign[bd[0]]], pl);
			pl[3] = (float)((double)pl[3] + fabs((double)off));
			if (!clip_facet_plane(pl, w->start, w->end, &enter, &leave, &hit)) break;
			if (hit) { hitnum = j; memcpy(best, pl, 16); }
		}
		if (j < fc[1]) continue;
		if (hitnum == fc[1] - 1) continue;       /* never clip against the back side */

		if (enter < leave && enter >= 0 && enter < w->tr.fraction) {
			w->tr.fraction = enter;
			memcpy(w->tr.plane_normal, best, 12);
			w->tr.plane_dist = best[3];
		}
	}
}

/* CM_PositionTestInPatchCollide, cm_patch.c:1479-1531 */
static int box_in_patch(const cmo_map_t *m, work_t *w, int p) {
	const int *pm = m->patches + 6 * (size_t)p;
	const float *planes = m->patchPlanes + 4 * (size_t)pm[2];
	const int *psign = m->patchPlaneSignbits + pm[2];
	if (w->isPoint) return 0;
	for (int f = 0; f < pm[5]; f++) {
		const int *fc = m->facets + 3 * (size_t)(pm[4] + f);
		float pl[4];
		const float *sp = planes + 4 * fc[0];
		w->info.facets++;
		pl[0] = sp[0]; pl[1] = sp[1]; pl[2] = sp[2];
		pl[3] = sp[3] - dot3f(w->corner[psign[fc[0]]], pl);
		if (dot3f(pl, w->start) - pl[3] > 0.0f) continue;
		int j;
		for (j = 0; j < fc[1]; j++) {
			const int *bd = m->borders + 2 * (size_t)(fc[2] + j);
			const float *bp = planes + 4 * bd[0];
			if (bd[1]) {
				pl[0] = -bp[0]; pl[1] = -bp[1]; pl[2] = -bp[2]; pl[3] = -bp[3];
			} else {
				pl[0] = bp[0]; pl[1] = bp[1]; pl[2] = bp[2]; pl[3] = bp[3];
			}
			float off = dot3f(w->corner[psign[bd[0]]], pl);
			pl[3] = (float)((double)pl[3] + fabs((double)off));
			if (dot3f(pl, w->start) - pl[3] > 0.0f) break;
		}
		if (j < fc[1]) continue;
		return 1;
	}
	return 0;
}

/* ---------------------------------------------------------------------- */
/* leaf level                                                              */
/* ---------------------------------------------------------------------- */

/* CM_TraceThroughLeaf (+CM_TraceThroughPatch), cm_trace.c:370-427, 241-254 */
static void sweep_leaf(const cmo_map_t *m, work_t *w, const int *leaf) {
	w->info.leafs++;
	for (int k = 0; k < leaf[1]; k++) {
		int b = m->leafbrushes[leaf[0] + k];
		const int *bm = m->brushes + 3 * (size_t)b;
		w->info.brushRefs++;
		if (!(bm[0] & w->mask)) continue;
		float mn[3], mx[3];
		brush_bounds(m, w, b, mn, mx);
		w->info.brushBounds++;
		if (!bounds_overlap(w->lo, w->hi, mn, mx)) continue;
		sweep_brush(m, w, b);
		if (!w->tr.fraction) return;
	}
	if (m->noCurves) return;
	for (int k = 0; k < leaf[3]; k++) {
		int p = m->surf2patch[m->leafsurfaces[leaf[2] + k]];
		if (p < 0) continue;
		const int *pm = m->patches + 6 * (size_t)p;
		if (!(pm[1] & w->mask)) continue;
		float before = w->tr.fraction;
		sweep_patch(m, w, p);
		if (w->tr.fraction < before) {
			w->tr.surfaceFlags = pm[0];
			w->tr.contents = pm[1];
			w->info.hitPatch = p;
			w->info.hitBrush = -1;
		}
		if (!w->tr.fraction) return;
	}
}

/* CM_TestBoxInBrush, cm_trace.c:83-123 */
static void box_in_brush(const cmo_map_t *m, work_t *w, int b) {
	const int *bm = m->brushes + 3 * (size_t)b;
	if (!bm[2]) return;
	float mn[3], mx[3];
	brush_bounds(m, w, b, mn, mx);
	w->info.brushBounds++;
	for (int i = 0; i < 3; i++) {
		if (w->lo[i] > mx[i]) return;
	}
	for (int i = 0; i < 3; i++) {
		if (w->hi[i] < mn[i]) return;
	}
	for (int s = 6; s < bm[2]; s++) {
		plane_t pl;
		fetch_plane(m, w, m->brushsides[2 * (size_t)(bm[1] + s)], &pl);
		w->info.brushSides++;
		/* float dot and float subtract, THEN widened (cm_trace.c:111) */
		float distf = pl.d - dot3f(w->corner[pl.signbits], pl.n);
		double d1 = dot3d(w->start, pl.n) - (double)distf;
		if (d1 > 0) return;
	}
	w->tr.startsolid = w->tr.allsolid = 1;
	w->tr.fraction = 0;
	w->tr.contents = bm[0];
	w->info.hitBrush = b;
	w->info.hitPatch = -1;
}

/* CM_TestInLeaf, cm_trace.c:130-183 */
static void test_leaf(const cmo_map_t *m, work_t *w, const int *leaf) {
	w->info.leafs++;
	for (int k = 0; k < leaf[1]; k++) {
		int b = m->leafbrushes[leaf[0] + k];
		w->info.brushRefs++;
		if (!(m->brushes[3 * (size_t)b] & w->mask)) continue;
		box_in_brush(m, w, b);
		if (w->tr.allsolid) return;
	}
	if (m->noCurves) return;
	for (int k = 0; k < leaf[3]; k++) {
		int p = m->surf2patch[m->leafsurfaces[leaf[2] + k]];
		if (p < 0) continue;
		const int *pm = m->patches + 6 * (size_t)p;
		if (!(pm[1] & w->mask)) continue;
		if (box_in_patch(m, w, p)) {
			w->tr.startsolid = w->tr.allsolid = 1;
			w->tr.fraction = 0;
			w->tr.contents = pm[1];
			w->info.hitPatch = p;
			w->info.hitBrush = -1;
			return;
		}
	}
}

/* ---------------------------------------------------------------------- */
/* tree level                                                              */
/* ---------------------------------------------------------------------- */

/* BoxOnPlaneSide, q_math.c:724-758 */
static int box_plane_side(const float *lo, const float *hi, const float *pl, int type, int signbits) {
	if (type < 3) {
		if (pl[3] <= lo[type]) return 1;
		if (pl[3] >= hi[type]) return 2;
		return 3;
	}
	float acc[2] = {0, 0};
	if (signbits < 8) {
		for (int i = 0; i < 3; i++) {
			int b = (signbits >> i) & 1;
			acc[b] += pl[i] * hi[i];
			acc[!b] += pl[i] * lo[i];
		}
	}
	int s = 0;
	if (acc[0] >= pl[3]) s = 1;
	if (acc[1] < pl[3]) s |= 2;
	return s;
}

/* CM_PositionTest, cm_trace.c:191-226 with CM_BoxLeafnums_r / CM_StoreLeafs,
 * cm_test.c:131-156,73-88.  Leafs are tested in discovery order, so testing
 * as they are found and stopping after 1024 is the same computation. */
static void position_test_world(const cmo_map_t *m, work_t *w) {
	float lo[3], hi[3];
	for (int i = 0; i < 3; i++) {
		lo[i] = w->start[i] + w->size[0][i];
		hi[i] = w->start[i] + w->size[1][i];
		lo[i] -= 1;
		hi[i] += 1;
	}
	int stack[STACK_MAX];
	int sp = 0;
	int found = 0;
	int num = 0;
	for (;;) {
		if (num < 0) {
			if (found < MAX_POS_LEAFS) {
				found++;
				test_leaf(m, w, m->leafs + 4 * (size_t)(-1 - num));
				if (w->tr.allsolid) return;
			} else {
				return;
			}
			if (!sp) return;
			num = stack[--sp];
			continue;
		}
		const int *nd = m->nodes + 3 * (size_t)num;
		w->info.nodes++;
		int s = box_plane_side(lo, hi, m->planes + 4 * (size_t)nd[0], m->planeType[nd[0]], m->planeSignbits[nd[0]]);
		if (s == 1) {
			num = nd[1];
		} else if (s == 2) {
			num = nd[2];
		} else {
			if (sp < STACK_MAX) stack[sp++] = nd[2];
			num = nd[1];
		}
	}
}

typedef struct { int num; float f1, f2; float a[3], b[3]; } seg_t;

/* CM_TraceThroughTree, cm_trace.c:439-538 */
static void sweep_world(const cmo_map_t *m, work_t *w) {
	seg_t stack[STACK_MAX];
	int sp = 0;
	seg_t cur;
	cur.num = 0; cur.f1 = 0; cur.f2 = 1;
	memcpy(cur.a, w->start, 12);
	memcpy(cur.b, w->end, 12);

	for (;;) {
		int done = 0;
		if (w->tr.fraction <= cur.f1) {
			done = 1;                         /* something nearer was already hit */
		} else if (cur.num < 0) {
			sweep_leaf(m, w, m->leafs + 4 * (size_t)(-1 - cur.num));
			done = 1;
		}
		if (done) {
			if (!sp) return;
			cur = stack[--sp];
			continue;
		}

		const int *nd = m->nodes + 3 * (size_t)cur.num;
		const float *pl = m->planes + 4 * (size_t)nd[0];
		int type = m->planeType[nd[0]];
		double t1, t2, off;
		w->info.nodes++;
		if (type < 3) {
			t1 = (double)(cur.a[type] - pl[3]);      /* float subtract, then widened */
			t2 = (double)(cur.b[type] - pl[3]);
			off = (double)w->extents[type];
		} else {
			t1 = dot3d(pl, cur.a) - (double)pl[3];
			t2 = dot3d(pl, cur.b) - (double)pl[3];
			off = w->isPoint ? 0.0 : 2048.0;
		}

		if (t1 >= off + 1 && t2 >= off + 1) { cur.num = nd[1]; continue; }
		if (t1 < -off - 1 && t2 < -off - 1) { cur.num = nd[2]; continue; }

		int side;
		float fnear, ffar;
		if (t1 < t2) {
			float inv = (float)(1.0 / (t1 - t2));
			side = 1;
			ffar = (float)((t1 + off + CLIP_EPS) * (double)inv);
			fnear = (float)((t1 - off + CLIP_EPS) * (double)inv);
		} else if (t1 > t2) {
			float inv = (float)(1.0 / (t1 - t2));
			side = 0;
			ffar = (float)((t1 - off - CLIP_EPS) * (double)inv);
			fnear = (float)((t1 + off + CLIP_EPS) * (double)inv);
		} else {
			side = 0;
			fnear = 1;
			ffar = 0;
		}
		if (fnear < 0) fnear = 0; else if (fnear > 1) fnear = 1;
		if (ffar < 0) ffar = 0; else if (ffar > 1) ffar = 1;
		w->info.splits++;

		/* far part [mid(ffar), b] waits on the stack; near part [a, mid(fnear)] goes first */
		seg_t far;
		far.num = nd[1 + (side ^ 1)];
		far.f1 = cur.f1 + (cur.f2 - cur.f1) * ffar;
		far.f2 = cur.f2;
		for (int i = 0; i < 3; i++) {
			far.a[i] = cur.a[i] + ffar * (cur.b[i] - cur.a[i]);
			far.b[i] = cur.b[i];
		}
		float midf = cur.f1 + (cur.f2 - cur.f1) * fnear;
		float mid[3];
		for (int i = 0; i < 3; i++) {
			mid[i] = cur.a[i] + fnear * (cur.b[i] - cur.a[i]);
		}
		if (sp < STACK_MAX) stack[sp++] = far;
		cur.num = nd[1 + side];
		cur.f2 = midf;
		memcpy(cur.b, mid, 12);
	}
}

/* ---------------------------------------------------------------------- */
/* CM_Trace, cm_trace.c:545-687                                            */
/* ---------------------------------------------------------------------- */
static const int *model_leaf(const cmo_map_t *m, int model, int *tmp) {
	if (model == CMO_BOX_MODEL_HANDLE) {
		/* box_model.leaf, cm_load.c:888-891 */
		tmp[0] = m->boxLeafBrush; tmp[1] = 1; tmp[2] = 0; tmp[3] = 0;
		return tmp;
	}
	return m->modelLeafs + 4 * (size_t)model;
}

static void trace_core(const cmo_map_t *m, work_t *w, const float *start, const float *end,
                       const float *mins, const float *maxs, int model, int mask) {
	static const float zero3[3] = {0, 0, 0};
	memset(&w->tr, 0, sizeof(w->tr));
	memset(&w->info, 0, sizeof(w->info));
	w->info.hitBrush = w->info.hitPatch = -1;
	w->tr.fraction = 1;
	w->isPoint = 0;
	memset(w->extents, 0, sizeof(w->extents));
	if (!m->numNodes) return;
	if (!mins) mins = zero3;
	if (!maxs) maxs = zero3;
	w->mask = mask;

	for (int i = 0; i < 3; i++) {
		float c = (float)((double)(mins[i] + maxs[i]) * 0.5);
		w->size[0][i] = mins[i] - c;
		w->size[1][i] = maxs[i] - c;
		w->start[i] = start[i] + c;
		w->end[i] = end[i] + c;
	}
	for (int s = 0; s < 8; s++) {
		for (int i = 0; i < 3; i++) {
			w->corner[s][i] = w->size[(s >> i) & 1][i];
		}
	}
	for (int i = 0; i < 3; i++) {
		if (w->start[i] < w->end[i]) {
			w->lo[i] = w->start[i] + w->size[0][i];
			w->hi[i] = w->end[i] + w->size[1][i];
		} else {
			w->lo[i] = w->end[i] + w->size[0][i];
			w->hi[i] = w->start[i] + w->size[1][i];
		}
	}

	int tmp[4];
	if (start[0] == end[0] && start[1] == end[1] && start[2] == end[2]) {
		if (model) {
			test_leaf(m, w, model_leaf(m, model, tmp));
		} else {
			position_test_world(m, w);
		}
	} else {
		if (w->size[0][0] == 0 && w->size[0][1] == 0 && w->size[0][2] == 0) {
			w->isPoint = 1;
		} else {
			memcpy(w->extents, w->size[1], 12);
		}
		if (model) {
			sweep_leaf(m, w, model_leaf(m, model, tmp));
		} else {
			sweep_world(m, w);
		}
	}

	if (w->tr.fraction == 1) {
		memcpy(w->tr.endpos, end, 12);
	} else {
		for (int i = 0; i < 3; i++) {
			w->tr.endpos[i] = start[i] + w->tr.fraction * (end[i] - start[i]);
		}
	}
}

void cmo_box_trace(const cmo_map_t *m, cmo_trace_t *out, cmo_info_t *info,
                   const float *start, const float *end, const float *mins, const float *maxs,
                   int model, int mask, const float *boxmins, const float *boxmaxs) {
	work_t w;
	w.boxmins = (model == CMO_BOX_MODEL_HANDLE) ? boxmins : NULL;
	w.boxmaxs = (model == CMO_BOX_MODEL_HANDLE) ? boxmaxs : NULL;
	trace_core(m, &w, start, end, mins, maxs, model, mask);
	*out = w.tr;
	if (info) *info = w.info;
}

/* AngleVectors, q_math.c:999-1032: the angle is scaled in double, rounded to
 * float, sin/cos taken in double of that float and rounded to float; all
 * products and sums are float. */
void cmo_angle_vectors(const float *angles, float *forward, float *right, float *up) {
	const double k = M_PI * 2 / 360;
	float a, sy, cy, sp, cp, sr, cr;
	a = (float)((double)angles[1] * k); sy = (float)sin((double)a); cy = (float)cos((double)a);
	a = (float)((double)angles[0] * k); sp = (float)sin((double)a); cp = (float)cos((double)a);
	a = (float)((double)angles[2] * k); sr = (float)sin((double)a); cr = (float)cos((double)a);
	if (forward) {
		forward[0] = cp * cy;
		forward[1] = cp * sy;
		forward[2] = -sp;
	}
	if (right) {
		float nsr = -1 * sr, ncr = -1 * cr;
		right[0] = (nsr * sp) * cy + ncr * -sy;
		right[1] = (nsr * sp) * sy + ncr * cy;
		right[2] = nsr * cp;
	}
	if (up) {
		up[0] = (cr * sp) * cy + -sr * -sy;
		up[1] = (cr * sp) * sy + -sr * cy;
		up[2] = cr * cp;
	}
}

/* CM_TransformedBoxTrace, cm_trace.c:708-805 */
void cmo_transformed_box_trace(const cmo_map_t *m, cmo_trace_t *out, cmo_info_t *info,
                               const float *start, const float *end, const float *mins, const float *maxs,
                               int model, int mask, const float *origin, const float *angles,
                               const float *boxmins, const float *boxmaxs) {
	static const float zero3[3] = {0, 0, 0};
	float sl[3], el[3], smin[3], smax[3];
	float mat[3][3];
	if (!mins) mins = zero3;
	if (!maxs) maxs = zero3;
	for (int i = 0; i < 3; i++) {
		float c = (float)((double)(mins[i] + maxs[i]) * 0.5);
		smin[i] = mins[i] - c;
		smax[i] = maxs[i] - c;
		sl[i] = start[i] + c;
		el[i] = end[i] + c;
	}
	for (int i = 0; i < 3; i++) {
		sl[i] = sl[i] - origin[i];
		el[i] = el[i] - origin[i];
	}
	int rotated = (angles[0] || angles[1] || angles[2]);
	if (rotated) {
		/* CreateRotationMatrix: rows forward, -right, up (cm_trace.c:65-68) */
		cmo_angle_vectors(angles, mat[0], mat[1], mat[2]);
		for (int i = 0; i < 3; i++) mat[1][i] = -mat[1][i];
		float t[3];
		memcpy(t, sl, 12);
		for (int i = 0; i < 3; i++) sl[i] = dot3f(mat[i], t);
		memcpy(t, el, 12);
		for (int i = 0; i < 3; i++) el[i] = dot3f(mat[i], t);
	}
	work_t w;
	w.boxmins = (model == CMO_BOX_MODEL_HANDLE) ? boxmins : NULL;
	w.boxmaxs = (model == CMO_BOX_MODEL_HANDLE) ? boxmaxs : NULL;
	trace_core(m, &w, sl, el, smin, smax, model, mask);
	if (rotated && w.tr.fraction != 1.0) {
		float t[3], col[3];
		memcpy(t, w.tr.plane_normal, 12);
		for (int i = 0; i < 3; i++) {
			col[0] = mat[0][i]; col[1] = mat[1][i]; col[2] = mat[2][i];
			w.tr.plane_normal[i] = dot3f(col, t);
		}
	}
	for (int i = 0; i < 3; i++) {
		w.tr.endpos[i] = start[i] + w.tr.fraction * (end[i] - start[i]);
	}
	*out = w.tr;
	if (info) *info = w.info;
}

/* CM_PointLeafnum_r, cm_test.c:31-54 */
int cmo_point_leafnum(const cmo_map_t *m, const float *p) {
	if (!m->numNodes) return 0;
	int num = 0;
	while (num >= 0) {
		const int *nd = m->nodes + 3 * (size_t)num;
		const float *pl = m->planes + 4 * (size_t)nd[0];
		int type = m->planeType[nd[0]];
		float d;
		if (type < 3) d = p[type] - pl[3];
		else d = dot3f(pl, p) - pl[3];
		num = (d < 0) ? nd[2] : nd[1];
	}
	return -1 - num;
}

/* CM_PointContents, cm_test.c:217-264 */
int cmo_point_contents(const cmo_map_t *m, const float *p, int model,
                       const float *boxmins, const float *boxmaxs) {
	if (!m->numNodes) return 0;
	work_t w;                      /* only the temp-box fields are used */
	w.boxmins = (model == CMO_BOX_MODEL_HANDLE) ? boxmins : NULL;
	w.boxmaxs = (model == CMO_BOX_MODEL_HANDLE) ? boxmaxs : NULL;
	int tmp[4];
	const int *leaf = model ? model_leaf(m, model, tmp)
	                        : m->leafs + 4 * (size_t)cmo_point_leafnum(m, p);
	int contents = 0;
	for (int k = 0; k < leaf[1]; k++) {
		int b = m->leafbrushes[leaf[0] + k];
		const int *bm = m->brushes + 3 * (size_t)b;
		float mn[3], mx[3];
		brush_bounds(m, &w, b, mn, mx);
		if (!bounds_has_point(mn, mx, p)) continue;
		int s;
		for (s = 0; s < bm[2]; s++) {
			plane_t pl;
			fetch_plane(m, &w, m->brushsides[2 * (size_t)(bm[1] + s)], &pl);
			if (dot3f(p, pl.n) > pl.d) break;
		}
		if (s == bm[2]) contents |= bm[0];
	}
	return contents;
}

/* CM_TransformedPointContents, cm_test.c:274-294 */
int cmo_transformed_point_contents(const cmo_map_t *m, const float *p, int model,
                                   const float *origin, const float *angles,
                                   const float *boxmins, const float *boxmaxs) {
	float l[3];
	for (int i = 0; i < 3; i++) l[i] = p[i] - origin[i];
	if (angles[0] || angles[1] || angles[2]) {
		float f[3], r[3], u[3], t[3];
		cmo_angle_vectors(angles, f, r, u);
		memcpy(t, l, 12);
		l[0] = dot3f(t, f);
		l[1] = -dot3f(t, r);
		l[2] = dot3f(t, u);
	}
	return cmo_point_contents(m, l, model, boxmins, boxmaxs);
}

/* ---------------------------------------------------------------------- */
/* batch wrappers                                                          */
/* ---------------------------------------------------------------------- */
void cmo_box_trace_batch(const cmo_map_t *m, int n, const float *starts, const float *ends,
                         const float *mins, const float *maxs, int hull_stride,
                         const int *models, const int *masks, int mask_stride,
                         const float *boxmins, const float *boxmaxs,
                         cmo_trace_t *out, cmo_info_t *info) {
	for (int i = 0; i < n; i++) {
		size_t o = 3 * (size_t)i;
		cmo_box_trace(m, &out[i], info ? &info[i] : NULL, starts + o, ends + o,
		              mins ? mins + o * hull_stride : NULL, maxs ? maxs + o * hull_stride : NULL,
		              models ? models[i] : 0, masks[(size_t)i * mask_stride],
		              boxmins ? boxmins + o : NULL, boxmaxs ? boxmaxs + o : NULL);
	}
}

void cmo_transformed_box_trace_batch(const cmo_map_t *m, int n, const float *starts, const float *ends,
                                     const float *mins, const float *maxs, int hull_stride,
                                     const int *models, const int *masks, int mask_stride,
                                     const float *origins, const float *angles,
                                     const float *boxmins, const float *boxmaxs,
                                     cmo_trace_t *out, cmo_info_t *info) {
	for (int i = 0; i < n; i++) {
		size_t o = 3 * (size_t)i;
		cmo_transformed_box_trace(m, &out[i], info ? &info[i] : NULL, starts + o, ends + o,
		                          mins ? mins + o * hull_stride : NULL, maxs ? maxs + o * hull_stride : NULL,
		                          models[i], masks[(size_t)i * mask_stride], origins + o, angles + o,
		                          boxmins ? boxmins + o : NULL, boxmaxs ? boxmaxs + o : NULL);
	}
}

void cmo_point_contents_batch(const cmo_map_t *m, int n, const float *pts, const int *models,
                              const float *origins, const float *angles,
                              const float *boxmins, const float *boxmaxs, int *out) {
	for (int i = 0; i < n; i++) {
		size_t o = 3 * (size_t)i;
		int model = models ? models[i] : 0;
		if (origins) {
			out[i] = cmo_transformed_point_contents(m, pts + o, model, origins + o, angles + o,
			                                        boxmins ? boxmins + o : NULL, boxmaxs ? boxmaxs + o : NULL);
		} else {
			out[i] = cmo_point_contents(m, pts + o, model,
			                            boxmins ? boxmins + o : NULL, boxmaxs ? boxmaxs + o : NULL);
		}
	}
}

/* ---------------------------------------------------------------------- */
/* leaf queries, entity linking (PVS part), areas and visibility            */
/* ---------------------------------------------------------------------- */

/* CM_BoxLeafnums, cm_test.c:163-181 with CM_BoxLeafnums_r :131-156 and CM_StoreLeafs :73-88.
 * leafClusterArea is [numLeafs][2] (cluster, area) or NULL (then no leaf has a cluster). */
int cmo_box_leafnums(const cmo_map_t *m, const int *leafClusterArea, const float *mins, const float *maxs,
                     int *list, int listsize, int *lastLeaf) {
	int stack[STACK_MAX];
	int sp = 0, count = 0, last = 0, num = 0;
	for (;;) {
		while (num >= 0) {
			const int *nd = m->nodes + 3 * (size_t)num;
			int s = box_plane_side(mins, maxs, m->planes + 4 * (size_t)nd[0], m->planeType[nd[0]], m->planeSignbits[nd[0]]);
			if (s == 1) {
				num = nd[1];
			} else if (s == 2) {
				num = nd[2];
			} else {
				if (sp < STACK_MAX) stack[sp++] = nd[2];   /* children[1] waits while children[0] is walked */
				num = nd[1];
			}
		}
		int leaf = -1 - num;
		if (leafClusterArea && leafClusterArea[2 * (size_t)leaf] != -1) last = leaf;   /* even when the list is full */
		if (count < listsize) list[count++] = leaf;
		if (!sp) break;
		num = stack[--sp];
	}
	if (lastLeaf) *lastLeaf = last;
	return count;
}

void cmo_box_leafnums_batch(const cmo_map_t *m, const int *leafClusterArea, int n, const float *mins, const float *maxs,
                            int listsize, int *lists, int *counts, int *lastLeafs) {
	for (int i = 0; i < n; i++) {
		counts[i] = cmo_box_leafnums(m, leafClusterArea, mins + 3 * (size_t)i, maxs + 3 * (size_t)i,
		                             lists + (size_t)i * listsize, listsize, &lastLeafs[i]);
	}
}

/* The PVS part of SV_LinkEntity, server/sv_world.c:255-303: the clusters and areas the box absmin..absmax touches.
 * out: numLeafs, areanum, areanum2, numClusters, lastCluster, clusternums[16]  (CMO_ENT_CLUSTER_INTS ints) */
void cmo_entity_clusters(const cmo_map_t *m, const int *leafClusterArea, const float *absmin, const float *absmax, int *out) {
	int leafs[128];                                    /* MAX_TOTAL_ENT_LEAFS, sv_world.c:178 */
	int lastLeaf = 0;
	int numLeafs = cmo_box_leafnums(m, leafClusterArea, absmin, absmax, leafs, 128, &lastLeaf);
	int areanum = -1, areanum2 = -1, numClusters = 0, lastCluster = 0, i;
	int *clusters = out + 5;
	for (i = 0; i < 16; i++) clusters[i] = 0;
	for (i = 0; i < numLeafs; i++) {
		int area = leafClusterArea ? leafClusterArea[2 * (size_t)leafs[i] + 1] : -1;
		if (area != -1) {
			if (areanum != -1 && areanum != area) areanum2 = area;   /* doors may straddle two areas */
			else areanum = area;
		}
	}
	for (i = 0; i < numLeafs; i++) {
		int cluster = leafClusterArea ? leafClusterArea[2 * (size_t)leafs[i]] : -1;
		if (cluster != -1) {
			clusters[numClusters++] = cluster;
			if (numClusters == 16) break;                /* MAX_ENT_CLUSTERS, server.h:35 */
		}
	}
	if (i != numLeafs) lastCluster = leafClusterArea ? leafClusterArea[2 * (size_t)lastLeaf] : -1;
	out[0] = numLeafs; out[1] = areanum; out[2] = areanum2; out[3] = numClusters; out[4] = lastCluster;
}

void cmo_entity_clusters_batch(const cmo_map_t *m, const int *leafClusterArea, int n, const float *absmins,
                               const float *absmaxs, int *out) {
	for (int i = 0; i < n; i++) {
		cmo_entity_clusters(m, leafClusterArea, absmins + 3 * (size_t)i, absmaxs + 3 * (size_t)i,
		                    out + (size_t)i * CMO_ENT_CLUSTER_INTS);
	}
}

/* CM_FloodAreaConnections, cm_test.c:348-365 with CM_FloodArea_r :322-342: floodnum[area] from the reference
 * counts areaPortals[numAreas][numAreas].  Depth-first in index order like the recursion. */
void cmo_flood_areas(int numAreas, const int *areaPortals, int *floodnum) {
	int floodvalid[256];
	int stack[256], iter[256];
	int next = 0;
	if (numAreas > 256) numAreas = 256;                /* the product refuses maps with more areas */
	for (int i = 0; i < numAreas; i++) { floodvalid[i] = 0; floodnum[i] = 0; }
	for (int i = 0; i < numAreas; i++) {
		if (floodvalid[i]) continue;
		next++;
		int sp = 0;
		stack[0] = i; iter[0] = 0;
		floodnum[i] = next; floodvalid[i] = 1;
		while (sp >= 0) {
			int a = stack[sp];
			if (iter[sp] >= numAreas) { sp--; continue; }
			int j = iter[sp]++;
			if (areaPortals[(size_t)a * numAreas + j] > 0 && !floodvalid[j]) {
				floodnum[j] = next; floodvalid[j] = 1;
				sp++;
				stack[sp] = j; iter[sp] = 0;
			}
		}
	}
}

/* The PVS stage of SV_AddEntitiesVisibleFromPoint, server/sv_snapshot.c:327-355, for one (viewpoint, entity) pair:
 * area connection (CM_AreasConnected, cm_test.c:404-424, cm_noAreas = 0), then the entity's clusters against the
 * viewpoint cluster's row of the visibility matrix (CM_ClusterPVS, cm_test.c:306-312), then the overflow range.
 * vis == NULL: not vised, every bit set.  ent = the CMO_ENT_CLUSTER_INTS ints of cmo_entity_clusters. */
int cmo_entity_visible(int numAreas, const int *floodnum, int numClusters, int clusterBytes, const uint8_t *vis,
                       int clientArea, int clientCluster, const int *ent) {
	int areanum = ent[1], areanum2 = ent[2], n = ent[3], lastCluster = ent[4];
	const int *clusters = ent + 5;
	int connected = 0;
	(void)numAreas;
	if (clientArea >= 0 && areanum >= 0 && floodnum[clientArea] == floodnum[areanum]) connected = 1;
	if (!connected && clientArea >= 0 && areanum2 >= 0 && floodnum[clientArea] == floodnum[areanum2]) connected = 1;
	if (!connected) return 0;
	if (!n) return 0;
	const uint8_t *row = NULL;                          /* NULL: all ones */
	if (vis) row = (clientCluster < 0 || clientCluster >= numClusters) ? vis : vis + (size_t)clientCluster * clusterBytes;
	int i, l = 0;
	for (i = 0; i < n; i++) {
		l = clusters[i];
		if (!row || (row[l >> 3] & (1 << (l & 7)))) break;
	}
	if (i == n) {
		if (lastCluster) {
			for (; l <= lastCluster; l++) {
				if (!row || (row[l >> 3] & (1 << (l & 7)))) break;
			}
			if (l == lastCluster) return 0;             /* sv_snapshot.c:352, kept as written */
		} else {
			return 0;
		}
	}
	return 1;
}

/* ================================================================================================================
 * botlib's AAS tracer: AAS_TraceClientBBox (code/botlib/be_aas_sample.c:399-665) and AAS_PointAreaNum (:212-258),
 * restated over plain arrays.  planes [n][5] (normal, dist, type), nodes [n][3] (planenum, children[0], children[1];
 * node 0 is the dummy, the root is node 1), presence[area].  Entity collisions (passent >= 0, :472-487) are the
 * server's SV_Trace and not part of this path: passent is -1.  All arithmetic is float except the two places the
 * reference's literals make double: (front +- 0.125) / (front - back) (:624-625) and VectorMA(.., -0.125, ..) (:459,518).
 * Pinned to the unmodified reference by tests/test_aas.py.
 * ================================================================================================================ */
typedef struct { int startsolid; float fraction; float endpos[3]; int ent, lastarea, area, planenum; } cmo_aas_trace_t;

static float aas_dot(const float *a, const float *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

static void aas_blocked(cmo_aas_trace_t *tr, const float *planes, const float *start, const float *end, float *segStart,
                        int planenum, int area) {
	float v1[3], v2[3];
	int i;
	if (segStart[0] == start[0] && segStart[1] == start[1] && segStart[2] == start[2]) {   /* :444-451 */
		tr->startsolid = 1;
		tr->fraction = 0.0f;
		v1[0] = v1[1] = v1[2] = 0.0f;
	} else {
		float len2, length, ilength, l2;
		tr->startsolid = 0;
		for (i = 0; i < 3; i++) { v1[i] = end[i] - start[i]; v2[i] = segStart[i] - start[i]; }
		l2 = sqrtf(v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2]);                           /* VectorLength, q_shared.h:558 */
		len2 = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2];                               /* VectorNormalize, q_math.c:857-874 */
		length = len2;
		if (len2) {
			ilength = 1 / (float)sqrt(len2);
			length = len2 * ilength;
			v1[0] *= ilength; v1[1] *= ilength; v1[2] *= ilength;
		}
		tr->fraction = l2 / length;
		for (i = 0; i < 3; i++) segStart[i] = (float)(segStart[i] + v1[i] * (-0.125));     /* VectorMA with a double scale */
	}
	for (i = 0; i < 3; i++) tr->endpos[i] = segStart[i];
	tr->ent = 0;
	tr->area = area;
	tr->planenum = planenum;
	if (aas_dot(v1, planes + 5 * (size_t)planenum) > 0) tr->planenum ^= 1;                  /* :468-469 */
}

void cmo_aas_trace_client_bbox(const float *planes, const int *nodes, const int *presence, const float *start, const float *end,
                               int presencetype, cmo_aas_trace_t *out) {
	struct { float start[3], end[3]; int planenum, nodenum; } stack[127], *sp = stack;
	cmo_aas_trace_t tr;
	int i;
	memset(&tr, 0, sizeof(tr));
	for (i = 0; i < 3; i++) { sp->start[i] = start[i]; sp->end[i] = end[i]; }
	sp->planenum = 0;
	sp->nodenum = 1;
	sp++;
	for (;;) {
		sp--;
		if (sp < stack) {                                          /* nothing was hit, :421-434 */
			tr.startsolid = 0;
			tr.fraction = 1.0f;
			for (i = 0; i < 3; i++) tr.endpos[i] = end[i];
			tr.ent = 0; tr.area = 0; tr.planenum = 0;
			break;
		}
		const int nodenum = sp->nodenum;
		if (nodenum < 0) {                                         /* an area, :438-490 */
			if (!(presence[-nodenum] & presencetype)) { aas_blocked(&tr, planes, start, end, sp->start, sp->planenum, -nodenum); break; }
			tr.lastarea = -nodenum;
			continue;
		}
		if (!nodenum) { aas_blocked(&tr, planes, start, end, sp->start, sp->planenum, 0); break; }   /* solid leaf, :492-528 */
		const int *nd = nodes + 3 * (size_t)nodenum;
		const float *pl = planes + 5 * (size_t)nd[0];
		float cs[3], ce[3];
		for (i = 0; i < 3; i++) { cs[i] = sp->start[i]; ce[i] = sp->end[i]; }
		float front = aas_dot(cs, pl) - pl[3], back = aas_dot(ce, pl) - pl[3];
		if (front >= 0 && back >= 0) {                             /* ON_EPSILON is 0, :590-603 */
			sp->nodenum = nd[1];
			sp++;
			if (sp >= &stack[127]) break;                          /* "stack overflow": the trace as it stands */
		} else if (front < 0 && back < 0) {
			sp->nodenum = nd[2];
			sp++;
			if (sp >= &stack[127]) break;
		} else {
			const int tmpplanenum = sp->planenum;
			float frac, mid[3];
			if (front == back) front -= 0.001f;
			if (front < 0) frac = (float)((front + 0.125) / (front - back));
			else frac = (float)((front - 0.125) / (front - back));
			if (frac < 0) frac = 0.001f;
			else if (frac > 1) frac = 0.999f;
			for (i = 0; i < 3; i++) mid[i] = cs[i] + (ce[i] - cs[i]) * frac;
			const int side = front < 0;
			for (i = 0; i < 3; i++) sp->start[i] = mid[i];          /* the far part keeps its end */
			sp->planenum = nd[0];
			sp->nodenum = nd[1 + !side];
			sp++;
			if (sp >= &stack[127]) break;
			for (i = 0; i < 3; i++) { sp->start[i] = cs[i]; sp->end[i] = mid[i]; }
			sp->planenum = tmpplanenum;
			sp->nodenum = nd[1 + side];
			sp++;
			if (sp >= &stack[127]) break;
		}
	}
	*out = tr;
}

void cmo_aas_trace_client_bbox_batch(const float *planes, const int *nodes, const int *presence, int n, const float *starts,
                                     const float *ends, const int *presencetypes, int presence_stride, cmo_aas_trace_t *out) {
	int i;
	for (i = 0; i < n; i++)
		cmo_aas_trace_client_bbox(planes, nodes, presence, starts + 3 * (size_t)i, ends + 3 * (size_t)i,
		                          presencetypes[presence_stride ? i : 0], out + i);
}

void cmo_aas_point_area_num_batch(const float *planes, const int *nodes, int n, const float *points, int *out) {
	int i;
	for (i = 0; i < n; i++) {
		int nodenum = 1;
		while (nodenum > 0) {
			const int *nd = nodes + 3 * (size_t)nodenum;
			const float *pl = planes + 5 * (size_t)nd[0];
			const float dist = aas_dot(points + 3 * (size_t)i, pl) - pl[3];
			nodenum = dist > 0 ? nd[1] : nd[2];
		}
		out[i] = nodenum ? -nodenum : 0;
	}
}
