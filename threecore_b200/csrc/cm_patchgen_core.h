// cm_patchgen_core.h -- Bezier patch -> collision facets, ONE implementation for the host and the device.
//
// Counterpart of CM_GeneratePatchCollide (code/qcommon/cm_patch.c:62-1203, helpers from cm_polylib.c:153-283,421-531).
// It must produce the same plane list and facet list -- same order, same float bits -- as the reference, because the
// run-time facet tests are only bit-exact if they see bit-identical planes.  The algorithm:
//
//   1. control grid -> points on the curve: per direction, drop the middle control column of a quadratic segment when
//      it is within 16 units of the chord, otherwise split the segment (cm_patch.c:89-330); remove columns that coincide
//      within 0.1; do rows the same way after a transpose.
//   2. two triangles per grid cell -> planes, merged when the three corners are within 0.1 of an existing plane
//      (cm_patch.c:473-513).
//   3. per cell one quad facet (coplanar triangles) or two triangle facets; borders are the neighbouring triangles'
//      planes or, on the rim, a plane standing on the edge (cm_patch.c:969-1120); a border faces inward or outward by
//      the side the facet's own corners lie on (cm_patch.c:651-722).
//   4. facets whose borders do not bound a finite winding are dropped (cm_patch.c:731-777); the rest get axial and edge
//      bevels plus the back plane (cm_patch.c:785-952).
//
// Everything works in a fixed scratch block (PatchScratch, ~700 KB: the reference's own static limits MAX_GRID_SIZE,
// MAX_PATCH_PLANES, MAX_FACETS, MAX_POINTS_ON_WINDING), with error codes instead of exceptions, so that the very same
// code runs as one thread per patch on the GPU (patchCollideKernel, cm_gpu.cu) and on the host (cm_patchgen.cpp).
// Float/double evaluation widths follow the reference expression by expression; both builds forbid FMA contraction
// (-ffp-contract=off / -fmad=false), and sqrt, division and conversions are IEEE on both sides.
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define PG_HD __host__ __device__
#else
#define PG_HD
#endif

namespace cmgpu {
namespace pg {

constexpr int    kMaxGrid        = 129;        // MAX_GRID_SIZE, cm_patch.h:47
constexpr int    kMaxFacets      = 1024;       // MAX_FACETS, cm_patch.h:23
constexpr int    kMaxPatchPlanes = 2048 + 128; // MAX_PATCH_PLANES, cm_patch.h:24
constexpr int    kMaxWinding     = 64;         // MAX_POINTS_ON_WINDING, cm_polylib.h:31
constexpr int    kMaxBorders     = 4 + 6 + 16; // facet_t, cm_patch.h:31-37
constexpr float  kMapBounds      = 65535.0f;   // MAX_MAP_BOUNDS, cm_polylib.h:40
constexpr double kNormalEps      = 0.0001;     // NORMAL_EPSILON, cm_patch.c:369
constexpr double kDistEps        = 0.02;       // DIST_EPSILON, cm_patch.c:370
constexpr double kTriEps         = 0.1;        // PLANE_TRI_EPSILON, cm_patch.h:59
constexpr double kPointEps       = 0.1;        // POINT_EPSILON / WRAP_POINT_EPSILON
constexpr float  kSubdivide      = 16.0f;      // SUBDIVIDE_DISTANCE, cm_patch.h:58

// what the reference raises ERR_DROP for
enum Error : int {
	PG_OK = 0, PG_BAD_PARAMS, PG_EVEN_SIZE, PG_SOURCE_TOO_BIG, PG_GRID_TOO_BIG, PG_MAX_PLANES, PG_MAX_FACETS, PG_NO_AXIS,
	PG_WINDING_ESTIMATE, PG_WINDING_POINTS, PG_BAD_EDGE
};

inline const char *errorText(int e) {
	switch (e) {
	case PG_BAD_PARAMS:       return "CM_GeneratePatchFacets: bad parameters";
	case PG_EVEN_SIZE:        return "CM_GeneratePatchFacets: even sizes are invalid for quadratic meshes";
	case PG_SOURCE_TOO_BIG:   return "CM_GeneratePatchFacets: source is > MAX_GRID_SIZE";
	case PG_GRID_TOO_BIG:     return "CM_SubdivideGridColumns: grid exceeds MAX_GRID_SIZE";
	case PG_MAX_PLANES:       return "MAX_PATCH_PLANES";
	case PG_MAX_FACETS:       return "MAX_FACETS";
	case PG_NO_AXIS:          return "BaseWindingForPlane: no axis found";
	case PG_WINDING_ESTIMATE: return "ClipWinding: points exceeded estimate";
	case PG_WINDING_POINTS:   return "ClipWinding: MAX_POINTS_ON_WINDING";
	case PG_BAD_EDGE:         return "CM_EdgePlaneNum: bad k";
	default:                  return "";
	}
}

struct Plane { float p[4]; int signbits; };

struct Facet {
	int32_t surfacePlane;
	int32_t numBorders;
	int32_t borderPlanes[kMaxBorders];
	int32_t borderInward[kMaxBorders];
	int32_t borderNoAdjust[kMaxBorders];
};

struct V3 { float v[3]; };
struct Winding { int n; V3 p[kMaxWinding + 4]; };

// the builder's whole state; results: planes[0..numPlanes), facets[0..numFacets), bounds, error
struct PatchScratch {
	float  pts[kMaxGrid][kMaxGrid][3];
	int    cellPlanes[kMaxGrid][kMaxGrid][2];
	Plane  planes[kMaxPatchPlanes];
	Facet  facets[kMaxFacets];
	float  bounds[6];
	int    numPlanes, numFacets;
	int    w, h;
	int    wrapW, wrapH;
	int    error;
};

// ---- arithmetic helpers, widths as in the reference's macros -----------------------------------------------------
PG_HD inline float dotF(const float *a, const float *b) {               // DotProduct
	float s = a[0] * b[0];
	s = s + a[1] * b[1];
	s = s + a[2] * b[2];
	return s;
}
PG_HD inline double dotD(const float *a, const float *b) {              // DotProductDP / DotProductDPf
	double s = (double)a[0] * (double)b[0];
	s = s + (double)a[1] * (double)b[1];
	s = s + (double)a[2] * (double)b[2];
	return s;
}
PG_HD inline void crossD(const float *a, const float *b, float *out) {  // CrossProductDP, cm_local.h:42-49
	const double a0 = a[0], a1 = a[1], a2 = a[2], b0 = b[0], b1 = b[1], b2 = b[2];
	const float c0 = (float)(a1 * b2 - a2 * b1);
	const float c1 = (float)(a2 * b0 - a0 * b2);
	const float c2 = (float)(a0 * b1 - a1 * b0);
	out[0] = c0; out[1] = c1; out[2] = c2;
}
PG_HD inline float normalizeD(float *v) {                                // VectorNormalizeDP, cm_local.h:51-68
	const double d0 = v[0], d1 = v[1], d2 = v[2];
	double length = d0 * d0 + d1 * d1 + d2 * d2;
	if (length) {
		const double inv = 1.0 / sqrt(length);
		length *= inv;
		v[0] = (float)(d0 * inv);
		v[1] = (float)(d1 * inv);
		v[2] = (float)(d2 * inv);
	}
	return (float)length;
}
PG_HD inline bool within(float d, double eps) { return !((double)d < -eps || (double)d > eps); }
PG_HD inline void copy3(float *dst, const float *src) { dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; }

// ---- windings: the parts of cm_polylib.c the facet validation needs -------------------------------------------------
PG_HD inline bool baseWinding(const float *normal, float dist, Winding &w, int &error) {   // BaseWindingForPlane, cm_polylib.c:197-263
	int axis = -1;
	float best = -kMapBounds;
	for (int i = 0; i < 3; i++) {
		const float v = (float)fabs((double)normal[i]);
		if (v > best) { axis = i; best = v; }
	}
	if (axis < 0) { error = PG_NO_AXIS; return false; }
	float up[3] = {0, 0, 0};
	up[axis == 2 ? 0 : 2] = 1.0f;
	const double d = dotD(up, normal);
	for (int i = 0; i < 3; i++) up[i] = (float)((double)up[i] + (double)normal[i] * -d);
	normalizeD(up);
	float org[3], right[3];
	for (int i = 0; i < 3; i++) org[i] = normal[i] * dist;
	crossD(up, normal, right);
	for (int i = 0; i < 3; i++) { up[i] = up[i] * kMapBounds; right[i] = right[i] * kMapBounds; }
	w.n = 4;
	for (int i = 0; i < 3; i++) {
		w.p[0].v[i] = (org[i] - right[i]) + up[i];
		w.p[1].v[i] = (org[i] + right[i]) + up[i];
		w.p[2].v[i] = (org[i] + right[i]) - up[i];
		w.p[3].v[i] = (org[i] - right[i]) - up[i];
	}
	return true;
}

// keep the part in front of (normal, dist); returns false when nothing is left.  ChopWindingInPlace, cm_polylib.c:421-531
PG_HD inline bool chopWinding(Winding &w, const float *normal, float dist, float epsilon, int &error) {
	const int n = w.n;
	float dists[kMaxWinding + 4 + 1];
	int sides[kMaxWinding + 4 + 1];
	int counts[3] = {0, 0, 0};
	for (int i = 0; i < n; i++) {
		const double dot = dotD(w.p[i].v, normal) - (double)dist;
		dists[i] = (float)dot;
		sides[i] = dot > (double)epsilon ? 0 : (dot < -(double)epsilon ? 1 : 2);   // front, back, on
		counts[sides[i]]++;
	}
	sides[n] = sides[0];
	dists[n] = dists[0];
	if (!counts[0]) { w.n = 0; return false; }
	if (!counts[1]) return true;
	Winding out;
	out.n = 0;
	const int maxOut = n + 4;
	for (int i = 0; i < n; i++) {
		const float *p1 = w.p[i].v;
		if (sides[i] == 2) {
			if (out.n >= kMaxWinding + 4) { error = PG_WINDING_POINTS; return false; }
			out.p[out.n++] = w.p[i];
			continue;
		}
		if (sides[i] == 0) {
			if (out.n >= kMaxWinding + 4) { error = PG_WINDING_POINTS; return false; }
			out.p[out.n++] = w.p[i];
		}
		if (sides[i + 1] == 2 || sides[i + 1] == sides[i]) continue;
		const float *p2 = w.p[(i + 1) % n].v;
		const double a = dists[i], b = dists[i + 1];
		const double t = a / (a - b);
		V3 mid;
		for (int j = 0; j < 3; j++) {
			if (normal[j] == 1.0f) mid.v[j] = dist;
			else if (normal[j] == -1.0f) mid.v[j] = -dist;
			else mid.v[j] = (float)((double)p1[j] + t * ((double)p2[j] - (double)p1[j]));
		}
		if (out.n >= kMaxWinding + 4) { error = PG_WINDING_POINTS; return false; }
		out.p[out.n++] = mid;
	}
	if (out.n > maxOut) { error = PG_WINDING_ESTIMATE; return false; }
	if (out.n > kMaxWinding) { error = PG_WINDING_POINTS; return false; }
	w.n = out.n;
	for (int i = 0; i < out.n; i++) w.p[i] = out.p[i];
	return true;
}

PG_HD inline void windingBounds(const Winding &w, float *mins, float *maxs) {   // cm_polylib.c:153-172
	for (int j = 0; j < 3; j++) { mins[j] = kMapBounds; maxs[j] = -kMapBounds; }
	for (int i = 0; i < w.n; i++) {
		for (int j = 0; j < 3; j++) {
			if (w.p[i].v[j] < mins[j]) mins[j] = w.p[i].v[j];
			if (w.p[i].v[j] > maxs[j]) maxs[j] = w.p[i].v[j];
		}
	}
}

// ---- the builder ------------------------------------------------------------------------------------------------------
struct Builder {
	PatchScratch &s;
	PG_HD explicit Builder(PatchScratch &scratch) : s(scratch) {}

	PG_HD float *P(int i, int j) { return s.pts[i][j]; }

	// ---------------- step 1: grid refinement -----------------
	PG_HD static bool curveIsBent(const float *a, const float *b, const float *c) {   // CM_NeedsSubdivision, cm_patch.c:89-112
		float delta[3];
		for (int i = 0; i < 3; i++) {
			const float chord = (float)(0.5 * (double)(a[i] + c[i]));
			const float curve = (float)(0.5 * (0.5 * (double)(a[i] + b[i]) + 0.5 * (double)(b[i] + c[i])));
			delta[i] = curve - chord;
		}
		const float len = sqrtf(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
		return len >= kSubdivide;
	}

	PG_HD void removeColumn(int col) {
		for (int j = 0; j < s.h; j++)
			for (int k = col + 1; k < s.w; k++) copy3(P(k - 1, j), P(k, j));
		s.w--;
	}

	PG_HD void subdivideColumns() {                                       // CM_SubdivideGridColumns, cm_patch.c:207-270
		for (int i = 0; i < s.w - 2;) {
			bool bent = false;
			for (int j = 0; j < s.h && !bent; j++) bent = curveIsBent(P(i, j), P(i + 1, j), P(i + 2, j));
			if (!bent) {
				removeColumn(i + 1);      // the middle control column adds nothing: keep the chord
				i++;
				continue;
			}
			if (s.w + 2 > kMaxGrid) { s.error = PG_GRID_TOO_BIG; return; }
			for (int j = 0; j < s.h; j++) {
				float a[3], b[3], c[3];
				copy3(a, P(i, j)); copy3(b, P(i + 1, j)); copy3(c, P(i + 2, j));
				for (int k = s.w - 1; k > i + 1; k--) copy3(P(k + 2, j), P(k, j));
				for (int k = 0; k < 3; k++) {                            // CM_Subdivide, cm_patch.c:123-131
					const float q1 = (float)(0.5 * (double)(a[k] + b[k]));
					const float q3 = (float)(0.5 * (double)(b[k] + c[k]));
					P(i + 1, j)[k] = q1;
					P(i + 3, j)[k] = q3;
					P(i + 2, j)[k] = (float)(0.5 * (double)(q1 + q3));
				}
			}
			s.w += 2;
			// the new middle column at i+1 is examined again
		}
	}

	PG_HD void dropDuplicateColumns() {                                   // CM_RemoveDegenerateColumns, cm_patch.c:330-355
		for (int i = 0; i < s.w - 1; i++) {
			bool same = true;
			for (int j = 0; j < s.h && same; j++)
				for (int k = 0; k < 3; k++)
					if (!within(P(i, j)[k] - P(i + 1, j)[k], kPointEps)) { same = false; break; }
			if (!same) continue;
			removeColumn(i + 1);
			i--;
		}
	}

	PG_HD void refineColumns() {
		// CM_SetGridWrapWidth, cm_patch.c:176-196: first and last column coincide within 0.1
		bool wrap = true;
		for (int j = 0; j < s.h && wrap; j++)
			for (int k = 0; k < 3; k++)
				if (!within(P(0, j)[k] - P(s.w - 1, j)[k], kPointEps)) { wrap = false; break; }
		s.wrapW = wrap ? 1 : 0;
		subdivideColumns();
		if (s.error) return;
		dropDuplicateColumns();
	}

	PG_HD void transpose() {                                              // CM_TransposeGrid, cm_patch.c:140-169
		const int n = s.w > s.h ? s.w : s.h;
		for (int i = 0; i < n; i++)
			for (int j = i + 1; j < n; j++) {
				float t[3];
				copy3(t, P(i, j)); copy3(P(i, j), P(j, i)); copy3(P(j, i), t);
			}
		const int tw = s.w; s.w = s.h; s.h = tw;
		const int tr = s.wrapW; s.wrapW = s.wrapH; s.wrapH = tr;
	}

	// ---------------- step 2: planes -----------------
	PG_HD static int signbitsOf(const float *n) { return (n[0] < 0 ? 1 : 0) | (n[1] < 0 ? 2 : 0) | (n[2] < 0 ? 4 : 0); }

	PG_HD int addPlane(const float *pl) {
		if (s.numPlanes == kMaxPatchPlanes) { s.error = PG_MAX_PLANES; return -1; }
		Plane &np = s.planes[s.numPlanes];
		np.p[0] = pl[0]; np.p[1] = pl[1]; np.p[2] = pl[2]; np.p[3] = pl[3];
		np.signbits = signbitsOf(pl);
		return s.numPlanes++;
	}

	// plane through three points, merged into an existing one when all three lie within 0.1 of it.  CM_FindPlane
	PG_HD int planeOfTriangle(const float *a, const float *b, const float *c) {
		float e1[3], e2[3], pl[4];
		for (int i = 0; i < 3; i++) { e1[i] = b[i] - a[i]; e2[i] = c[i] - a[i]; }
		crossD(e2, e1, pl);                                           // CM_PlaneFromPoints, cm_patch.c:62-75
		if (normalizeD(pl) == 0.0f) return -1;
		pl[3] = (float)dotD(a, pl);
		for (int i = 0; i < s.numPlanes; i++) {
			const float *q = s.planes[i].p;
			if (dotF(pl, q) < 0) continue;
			if (!within(dotF(a, q) - q[3], kTriEps)) continue;
			if (!within(dotF(b, q) - q[3], kTriEps)) continue;
			if (!within(dotF(c, q) - q[3], kTriEps)) continue;
			return i;
		}
		return addPlane(pl);
	}

	PG_HD static bool nearF(float a, float b, double eps) { return fabs((double)(a - b)) < eps; }

	// CM_PlaneEqual, cm_patch.c:377-405: same plane, or the same plane facing the other way
	PG_HD static bool samePlane(const Plane &p, const float *pl, int *flipped) {
		if (nearF(p.p[0], pl[0], kNormalEps) && nearF(p.p[1], pl[1], kNormalEps) && nearF(p.p[2], pl[2], kNormalEps) &&
		    nearF(p.p[3], pl[3], kDistEps)) { *flipped = 0; return true; }
		const float inv[4] = {-pl[0], -pl[1], -pl[2], -pl[3]};
		if (nearF(p.p[0], inv[0], kNormalEps) && nearF(p.p[1], inv[1], kNormalEps) && nearF(p.p[2], inv[2], kNormalEps) &&
		    nearF(p.p[3], inv[3], kDistEps)) { *flipped = 1; return true; }
		return false;
	}

	PG_HD int planeLike(const float *pl, int *flipped) {                  // CM_FindPlane2, cm_patch.c:439-460
		for (int i = 0; i < s.numPlanes; i++)
			if (samePlane(s.planes[i], pl, flipped)) return i;
		*flipped = 0;
		return addPlane(pl);
	}

	PG_HD int sideOfPlane(const float *p, int plane) const {              // CM_PointOnPlaneSide: 0 front, 1 back, 2 on
		if (plane == -1) return 2;
		const float *q = s.planes[plane].p;
		const double d = dotD(p, q) - (double)q[3];
		if (d > kTriEps) return 0;
		if (d < -kTriEps) return 1;
		return 2;
	}

	PG_HD int cellPlane(int i, int j, int tri) const {                    // CM_GridPlane, cm_patch.c:545-560
		const int p = s.cellPlanes[i][j][tri];
		if (p != -1) return p;
		return s.cellPlanes[i][j][!tri];
	}

	// a plane standing on one edge of cell (i,j), perpendicular to that cell's surface.  CM_EdgePlaneNum, cm_patch.c:568-642
	PG_HD int rimPlane(int i, int j, int which) {
		const float *a, *b;
		int tri;
		bool reversed = false;
		switch (which) {
		case 0: a = P(i, j);         b = P(i + 1, j);     tri = 0; break;                  // top
		case 2: a = P(i, j + 1);     b = P(i + 1, j + 1); tri = 1; reversed = true; break; // bottom
		case 3: a = P(i, j);         b = P(i, j + 1);     tri = 1; reversed = true; break; // left
		case 1: a = P(i + 1, j);     b = P(i + 1, j + 1); tri = 0; break;                  // right
		case 4: a = P(i + 1, j + 1); b = P(i, j);         tri = 0; break;                  // diagonal, leaving triangle 0
		case 5: a = P(i, j);         b = P(i + 1, j + 1); tri = 1; break;                  // diagonal, leaving triangle 1
		default: s.error = PG_BAD_EDGE; return -1;
		}
		const int p = cellPlane(i, j, tri);
		if (p == -1) return -1;
		float up[3];
		for (int k = 0; k < 3; k++) up[k] = a[k] + s.planes[p].p[k] * 4.0f;   // VectorMA( p1, 4, normal, up )
		return reversed ? planeOfTriangle(b, a, up) : planeOfTriangle(a, b, up);
	}

	// ---------------- step 3/4: facets -----------------
	PG_HD void orientBorders(Facet &f, int i, int j, int which) {      // CM_SetBorderInward, cm_patch.c:651-722
		const float *corner[4];
		int n;
		if (which == -1) { corner[0] = P(i, j); corner[1] = P(i + 1, j); corner[2] = P(i + 1, j + 1); corner[3] = P(i, j + 1); n = 4; }
		else if (which == 0) { corner[0] = P(i, j); corner[1] = P(i + 1, j); corner[2] = P(i + 1, j + 1); corner[3] = nullptr; n = 3; }
		else { corner[0] = P(i + 1, j + 1); corner[1] = P(i, j + 1); corner[2] = P(i, j); corner[3] = nullptr; n = 3; }
		for (int k = 0; k < f.numBorders; k++) {
			int front = 0, back = 0;
			for (int l = 0; l < n; l++) {
				const int sd = sideOfPlane(corner[l], f.borderPlanes[k]);
				if (sd == 0) front++; else if (sd == 1) back++;
			}
			if (front && !back) f.borderInward[k] = 1;
			else if (back && !front) f.borderInward[k] = 0;
			else if (!front && !back) f.borderPlanes[k] = -1;            // flat side border
			else f.borderInward[k] = 0;                                   // bisecting border ("mixed plane sides")
		}
	}

	// the facet's surface polygon cut by its borders; false when a border is missing or nothing is left
	PG_HD bool facetWinding(const Facet &f, bool skipSurfaceBorders, bool failOnMissing, Winding &w) {
		const float *sp = s.planes[f.surfacePlane].p;
		if (!baseWinding(sp, sp[3], w, s.error)) return false;
		for (int j = 0; j < f.numBorders; j++) {
			if (f.borderPlanes[j] == -1) { if (failOnMissing) return false; continue; }
			if (skipSurfaceBorders && f.borderPlanes[j] == f.surfacePlane) continue;
			float pl[4];
			const float *bp = s.planes[f.borderPlanes[j]].p;
			pl[0] = bp[0]; pl[1] = bp[1]; pl[2] = bp[2]; pl[3] = bp[3];
			if (!f.borderInward[j]) {
				for (int k = 0; k < 3; k++) pl[k] = 0.0f - pl[k];          // VectorSubtract( vec3_origin, plane, plane )
				pl[3] = -pl[3];
			}
			if (!chopWinding(w, pl, pl[3], 0.1f, s.error)) return false;
		}
		return true;
	}

	PG_HD bool facetIsBounded(const Facet &f) {                          // CM_ValidateFacet, cm_patch.c:731-777
		if (f.surfacePlane == -1) return false;
		Winding w;
		if (!facetWinding(f, false, true, w)) return false;
		float mins[3], maxs[3];
		windingBounds(w, mins, maxs);
		for (int j = 0; j < 3; j++) {
			if (maxs[j] - mins[j] > kMapBounds) return false;
			if (mins[j] >= kMapBounds) return false;
			if (maxs[j] <= -kMapBounds) return false;
		}
		return true;
	}

	PG_HD bool haveBorderLike(const Facet &f, const float *pl, int *flipped) const {
		for (int i = 0; i < f.numBorders; i++)
			if (samePlane(s.planes[f.borderPlanes[i]], pl, flipped)) return true;
		return false;
	}

	PG_HD void addBevels(Facet &f) {                                    // CM_AddFacetBevels, cm_patch.c:785-952
		Winding w;
		if (!facetWinding(f, true, false, w)) return;
		float mins[3], maxs[3];
		windingBounds(w, mins, maxs);
		int flipped;
		// the six axial planes through the facet's bounds
		for (int axis = 0; axis < 3; axis++) {
			for (int dir = -1; dir <= 1; dir += 2) {
				float pl[4] = {0, 0, 0, 0};
				pl[axis] = (float)dir;
				pl[3] = dir == 1 ? maxs[axis] : -mins[axis];
				if (samePlane(s.planes[f.surfacePlane], pl, &flipped)) continue;
				if (haveBorderLike(f, pl, &flipped)) continue;
				if (f.numBorders >= kMaxBorders) continue;            // "too many bevels"
				f.borderPlanes[f.numBorders] = planeLike(pl, &flipped);
				if (s.error) return;
				f.borderNoAdjust[f.numBorders] = 0;
				f.borderInward[f.numBorders] = flipped;
				f.numBorders++;
			}
		}
		// slanted bevels along the non-axial edges of the facet polygon
		const int np = w.n;
		for (int j = 0; j < np; j++) {
			const int k = (j + 1) % np;
			float edge[3];
			for (int c = 0; c < 3; c++) edge[c] = (float)((double)w.p[j].v[c] - (double)w.p[k].v[c]);
			if (normalizeD(edge) < 0.5) continue;                          // degenerate edge
			for (int c = 0; c < 3; c++) {                                   // CM_SnapVector, cm_patch.c:413-431
				if (fabs((double)(edge[c] - 1)) < kNormalEps) { edge[0] = edge[1] = edge[2] = 0; edge[c] = 1; break; }
				if (fabs((double)(edge[c] - -1)) < kNormalEps) { edge[0] = edge[1] = edge[2] = 0; edge[c] = -1; break; }
			}
			if (edge[0] == -1 || edge[0] == 1 || edge[1] == -1 || edge[1] == 1 || edge[2] == -1 || edge[2] == 1) continue;
			for (int axis = 0; axis < 3; axis++) {
				for (int dir = -1; dir <= 1; dir += 2) {
					float unit[3] = {0, 0, 0}, pl[4];
					unit[axis] = (float)dir;
					crossD(edge, unit, pl);
					if (normalizeD(pl) < 0.5) continue;
					pl[3] = (float)dotD(w.p[j].v, pl);
					bool behind = true;                                          // every polygon point must be behind the bevel
					for (int l = 0; l < np; l++) {
						if (dotD(w.p[l].v, pl) - (double)pl[3] > 0.1) { behind = false; break; }
					}
					if (!behind) continue;
					if (samePlane(s.planes[f.surfacePlane], pl, &flipped)) continue;
					if (haveBorderLike(f, pl, &flipped)) continue;
					if (f.numBorders >= kMaxBorders) continue;
					const int slot = f.numBorders;
					f.borderPlanes[slot] = planeLike(pl, &flipped);
					if (s.error) return;
					f.borderNoAdjust[slot] = 0;
					f.borderInward[slot] = flipped;
					// the bevel must leave some of the polygon: otherwise the slot is not claimed
					// (the plane stays in the list, as in the reference)
					float cut[4];
					const float *bp = s.planes[f.borderPlanes[slot]].p;
					cut[0] = bp[0]; cut[1] = bp[1]; cut[2] = bp[2]; cut[3] = bp[3];
					if (!f.borderInward[slot]) { cut[0] = -cut[0]; cut[1] = -cut[1]; cut[2] = -cut[2]; cut[3] = -cut[3]; }
					Winding w2;
					w2.n = w.n;
					for (int l = 0; l < w.n; l++) w2.p[l] = w.p[l];
					const bool left = chopWinding(w2, cut, cut[3], 0.1f, s.error);
					if (s.error) return;
					if (!left) continue;
					f.numBorders++;
				}
			}
		}
		// the back side
		if (f.numBorders >= kMaxBorders) return;
		f.borderPlanes[f.numBorders] = f.surfacePlane;
		f.borderNoAdjust[f.numBorders] = 0;
		f.borderInward[f.numBorders] = 1;
		f.numBorders++;
	}

	PG_HD void keepIfBounded(Facet &f) {
		const bool bounded = facetIsBounded(f);
		if (s.error) return;
		if (bounded) {
			addBevels(f);
			if (s.error) return;
			if (s.numFacets == kMaxFacets) { s.error = PG_MAX_FACETS; return; }
			s.facets[s.numFacets++] = f;
		}
	}

	PG_HD static void clearFacet(Facet &f) {
		f.surfacePlane = 0; f.numBorders = 0;
		for (int k = 0; k < kMaxBorders; k++) { f.borderPlanes[k] = 0; f.borderInward[k] = 0; f.borderNoAdjust[k] = 0; }
	}

	PG_HD void buildFacets() {                                               // CM_PatchCollideFromGrid, cm_patch.c:969-1120
		s.numPlanes = 0;
		s.numFacets = 0;
		for (int i = 0; i < s.w - 1; i++)
			for (int j = 0; j < s.h - 1; j++) {
				s.cellPlanes[i][j][0] = planeOfTriangle(P(i, j), P(i + 1, j), P(i + 1, j + 1));
				if (s.error) return;
				s.cellPlanes[i][j][1] = planeOfTriangle(P(i + 1, j + 1), P(i, j + 1), P(i, j));
				if (s.error) return;
			}
		for (int i = 0; i < s.w - 1; i++) {
			for (int j = 0; j < s.h - 1; j++) {
				const int tri0 = s.cellPlanes[i][j][0], tri1 = s.cellPlanes[i][j][1];
				// neighbour across each edge (wrapping when the patch closes on itself), else a rim plane
				int top = -1, bottom = -1, left = -1, right = -1;
				if (j > 0) top = s.cellPlanes[i][j - 1][1]; else if (s.wrapH) top = s.cellPlanes[i][s.h - 2][1];
				const int topSame = (top == tri0);
				if (top == -1 || topSame) top = rimPlane(i, j, 0);
				if (j < s.h - 2) bottom = s.cellPlanes[i][j + 1][0]; else if (s.wrapH) bottom = s.cellPlanes[i][0][0];
				const int bottomSame = (bottom == tri1);
				if (bottom == -1 || bottomSame) bottom = rimPlane(i, j, 2);
				if (i > 0) left = s.cellPlanes[i - 1][j][0]; else if (s.wrapW) left = s.cellPlanes[s.w - 2][j][0];
				const int leftSame = (left == tri1);
				if (left == -1 || leftSame) left = rimPlane(i, j, 3);
				if (i < s.w - 2) right = s.cellPlanes[i + 1][j][1]; else if (s.wrapW) right = s.cellPlanes[0][j][1];
				const int rightSame = (right == tri0);
				if (right == -1 || rightSame) right = rimPlane(i, j, 1);
				if (s.error) return;

				if (s.numFacets == kMaxFacets) { s.error = PG_MAX_FACETS; return; }
				Facet f;
				clearFacet(f);
				if (tri0 == tri1) {
					if (tri0 == -1) continue;                                // degenerate cell
					f.surfacePlane = tri0;
					f.numBorders = 4;
					f.borderPlanes[0] = top;    f.borderNoAdjust[0] = topSame;
					f.borderPlanes[1] = right;  f.borderNoAdjust[1] = rightSame;
					f.borderPlanes[2] = bottom; f.borderNoAdjust[2] = bottomSame;
					f.borderPlanes[3] = left;   f.borderNoAdjust[3] = leftSame;
					orientBorders(f, i, j, -1);
					keepIfBounded(f);
					if (s.error) return;
				} else {
					f.surfacePlane = tri0;
					f.numBorders = 3;
					f.borderPlanes[0] = top;   f.borderNoAdjust[0] = topSame;
					f.borderPlanes[1] = right; f.borderNoAdjust[1] = rightSame;
					f.borderPlanes[2] = tri1;
					if (f.borderPlanes[2] == -1) {
						f.borderPlanes[2] = bottom;
						if (f.borderPlanes[2] == -1) f.borderPlanes[2] = rimPlane(i, j, 4);
					}
					if (s.error) return;
					orientBorders(f, i, j, 0);
					keepIfBounded(f);
					if (s.error) return;

					if (s.numFacets == kMaxFacets) { s.error = PG_MAX_FACETS; return; }
					clearFacet(f);
					f.surfacePlane = tri1;
					f.numBorders = 3;
					f.borderPlanes[0] = bottom; f.borderNoAdjust[0] = bottomSame;
					f.borderPlanes[1] = left;   f.borderNoAdjust[1] = leftSame;
					f.borderPlanes[2] = tri0;
					if (f.borderPlanes[2] == -1) {
						f.borderPlanes[2] = top;
						if (f.borderPlanes[2] == -1) f.borderPlanes[2] = rimPlane(i, j, 5);
					}
					if (s.error) return;
					orientBorders(f, i, j, 1);
					keepIfBounded(f);
					if (s.error) return;
				}
			}
		}
	}

	// points: [height*width][3] control grid, row-major as stored in the BSP (cm_load.c:567-572)
	PG_HD void run(int width, int height, const float *points) {
		s.error = PG_OK;
		s.numPlanes = s.numFacets = 0;
		if (width <= 2 || height <= 2 || !points) { s.error = PG_BAD_PARAMS; return; }
		if (!(width & 1) || !(height & 1)) { s.error = PG_EVEN_SIZE; return; }
		if (width > kMaxGrid || height > kMaxGrid) { s.error = PG_SOURCE_TOO_BIG; return; }
		s.w = width; s.h = height;
		for (int i = 0; i < width; i++)
			for (int j = 0; j < height; j++) copy3(P(i, j), points + 3 * ((size_t)j * width + i));
		s.wrapW = s.wrapH = 0;
		refineColumns();
		if (s.error) return;
		transpose();
		refineColumns();
		if (s.error) return;

		float mins[3] = {99999, 99999, 99999}, maxs[3] = {-99999, -99999, -99999};   // ClearBounds / AddPointToBounds
		for (int i = 0; i < s.w; i++)
			for (int j = 0; j < s.h; j++)
				for (int k = 0; k < 3; k++) {
					const float v = P(i, j)[k];
					if (v < mins[k]) mins[k] = v;
					if (v > maxs[k]) maxs[k] = v;
				}
		buildFacets();
		for (int k = 0; k < 3; k++) { s.bounds[k] = mins[k] - 1; s.bounds[3 + k] = maxs[k] + 1; }   // cm_patch.c:1190-1196
	}
};

}  // namespace pg
}  // namespace cmgpu
