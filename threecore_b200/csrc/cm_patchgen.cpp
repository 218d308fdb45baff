// cm_patchgen.cpp -- load-time Bezier patch -> collision facets, host only.
//
// Produces, for one MST_PATCH surface, the same plane list and facet list (same order, same
// float bits) that the reference builds in CM_GeneratePatchCollide
// (code/qcommon/cm_patch.c:62-1203, helpers from cm_polylib.c:153-283,421-531), because the
// run-time facet tests are only bit-exact if they see bit-identical planes.  The algorithm:
//
//   1. control grid -> points on the curve: per direction, drop the middle control column of a
//      quadratic segment when it is within 16 units of the chord, otherwise split the segment
//      (cm_patch.c:89-330); remove columns that coincide within 0.1; do rows the same way after a
//      transpose.
//   2. two triangles per grid cell -> planes, merged when the three corners are within 0.1 of an
//      existing plane (cm_patch.c:473-513).
//   3. per cell one quad facet (coplanar triangles) or two triangle facets; borders are the
//      neighbouring triangles' planes or, on the rim, a plane standing on the edge
//      (cm_patch.c:969-1120); a border faces inward or outward by the side the facet's own corners
//      lie on (cm_patch.c:651-722).
//   4. facets whose borders do not bound a finite winding are dropped (cm_patch.c:731-777); the
//      rest get axial and edge bevels plus the back plane (cm_patch.c:785-952).
//
// Float/double evaluation widths follow the reference expression by expression (the reference is
// built without FMA contraction; this file is compiled with -ffp-contract=off).
#include "cm_patchgen.h"

#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <memory>
#include <stdexcept>
#include <vector>

namespace cmgpu {
namespace {

constexpr int   kMaxGrid        = 129;        // MAX_GRID_SIZE, cm_patch.h:47
constexpr int   kMaxFacets      = 1024;       // MAX_FACETS, cm_patch.h:23
constexpr int   kMaxPatchPlanes = 2048 + 128; // MAX_PATCH_PLANES, cm_patch.h:24
constexpr int   kMaxWinding     = 64;         // MAX_POINTS_ON_WINDING, cm_polylib.h:31
constexpr float kMapBounds      = 65535.0f;   // MAX_MAP_BOUNDS, cm_polylib.h:40
constexpr double kNormalEps     = 0.0001;     // NORMAL_EPSILON, cm_patch.c:369
constexpr double kDistEps       = 0.02;       // DIST_EPSILON, cm_patch.c:370
constexpr double kTriEps        = 0.1;        // PLANE_TRI_EPSILON, cm_patch.h:59
constexpr double kPointEps      = 0.1;        // POINT_EPSILON / WRAP_POINT_EPSILON
constexpr float kSubdivide      = 16.0f;      // SUBDIVIDE_DISTANCE, cm_patch.h:58

struct Drop : std::runtime_error {
	using std::runtime_error::runtime_error;
};

struct V3 { float v[3]; };
using Winding = std::vector<V3>;

// ---- arithmetic helpers, widths as in the reference's macros -------------------------------
inline float dotF(const float *a, const float *b) {               // DotProduct
	float s = a[0] * b[0];
	s = s + a[1] * b[1];
	s = s + a[2] * b[2];
	return s;
}
inline double dotD(const float *a, const float *b) {              // DotProductDP / DotProductDPf
	double s = (double)a[0] * (double)b[0];
	s = s + (double)a[1] * (double)b[1];
	s = s + (double)a[2] * (double)b[2];
	return s;
}
inline void crossD(const float *a, const float *b, float *out) {  // CrossProductDP, cm_local.h:42-49
	const double a0 = a[0], a1 = a[1], a2 = a[2], b0 = b[0], b1 = b[1], b2 = b[2];
	const float c0 = (float)(a1 * b2 - a2 * b1);
	const float c1 = (float)(a2 * b0 - a0 * b2);
	const float c2 = (float)(a0 * b1 - a1 * b0);
	out[0] = c0; out[1] = c1; out[2] = c2;
}
inline float normalizeD(float *v) {                                // VectorNormalizeDP, cm_local.h:51-68
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
inline bool within(float d, double eps) { return !((double)d < -eps || (double)d > eps); }

// ---- windings: the parts of cm_polylib.c the facet validation needs -------------------------
Winding baseWinding(const float *normal, float dist) {             // BaseWindingForPlane, cm_polylib.c:197-263
	int axis = -1;
	float best = -kMapBounds;
	for (int i = 0; i < 3; i++) {
		const float v = (float)fabs((double)normal[i]);
		if (v > best) { axis = i; best = v; }
	}
	if (axis < 0) throw Drop("BaseWindingForPlane: no axis found");
	float up[3] = {0, 0, 0};
	up[axis == 2 ? 0 : 2] = 1.0f;
	const double d = dotD(up, normal);
	for (int i = 0; i < 3; i++) up[i] = (float)((double)up[i] + (double)normal[i] * -d);
	normalizeD(up);
	float org[3], right[3];
	for (int i = 0; i < 3; i++) org[i] = normal[i] * dist;
	crossD(up, normal, right);
	for (int i = 0; i < 3; i++) { up[i] = up[i] * kMapBounds; right[i] = right[i] * kMapBounds; }
	Winding w(4);
	for (int i = 0; i < 3; i++) {
		w[0].v[i] = (org[i] - right[i]) + up[i];
		w[1].v[i] = (org[i] + right[i]) + up[i];
		w[2].v[i] = (org[i] + right[i]) - up[i];
		w[3].v[i] = (org[i] - right[i]) - up[i];
	}
	return w;
}

// keep the part in front of (normal, dist); returns false when nothing is left.  ChopWindingInPlace, cm_polylib.c:421-531
bool chopWinding(Winding &w, const float *normal, float dist, float epsilon) {
	const int n = (int)w.size();
	float dists[kMaxWinding + 4 + 1];
	int sides[kMaxWinding + 4 + 1];
	int counts[3] = {0, 0, 0};
	for (int i = 0; i < n; i++) {
		const double dot = dotD(w[i].v, normal) - (double)dist;
		dists[i] = (float)dot;
		sides[i] = dot > (double)epsilon ? 0 : (dot < -(double)epsilon ? 1 : 2);   // front, back, on
		counts[sides[i]]++;
	}
	sides[n] = sides[0];
	dists[n] = dists[0];
	if (!counts[0]) { w.clear(); return false; }
	if (!counts[1]) return true;
	Winding out;
	out.reserve((size_t)n + 4);
	for (int i = 0; i < n; i++) {
		const float *p1 = w[i].v;
		if (sides[i] == 2) { out.push_back(w[i]); continue; }
		if (sides[i] == 0) out.push_back(w[i]);
		if (sides[i + 1] == 2 || sides[i + 1] == sides[i]) continue;
		const float *p2 = w[(i + 1) % n].v;
		const double a = dists[i], b = dists[i + 1];
		const double t = a / (a - b);
		V3 mid;
		for (int j = 0; j < 3; j++) {
			if (normal[j] == 1.0f) mid.v[j] = dist;
			else if (normal[j] == -1.0f) mid.v[j] = -dist;
			else mid.v[j] = (float)((double)p1[j] + t * ((double)p2[j] - (double)p1[j]));
		}
		out.push_back(mid);
	}
	if ((int)out.size() > n + 4) throw Drop("ClipWinding: points exceeded estimate");
	if ((int)out.size() > kMaxWinding) throw Drop("ClipWinding: MAX_POINTS_ON_WINDING");
	w.swap(out);
	return true;
}

void windingBounds(const Winding &w, float *mins, float *maxs) {   // cm_polylib.c:153-172
	for (int j = 0; j < 3; j++) { mins[j] = kMapBounds; maxs[j] = -kMapBounds; }
	for (const V3 &p : w) {
		for (int j = 0; j < 3; j++) {
			if (p.v[j] < mins[j]) mins[j] = p.v[j];
			if (p.v[j] > maxs[j]) maxs[j] = p.v[j];
		}
	}
}

// ---- the builder -----------------------------------------------------------------------------------
struct Plane { float p[4]; int signbits; };

class Builder {
public:
	Builder() : pts_(new float[kMaxGrid][kMaxGrid][3]), cellPlanes_(new int[kMaxGrid][kMaxGrid][2]) {}

	void run(int width, int height, const float *points, PatchCollide *out) {
		w_ = width; h_ = height;
		for (int i = 0; i < width; i++)
			for (int j = 0; j < height; j++) memcpy(P(i, j), points + 3 * ((size_t)j * width + i), 12);
		wrapW_ = wrapH_ = false;
		refineColumns();
		transpose();
		refineColumns();

		float mins[3] = {99999, 99999, 99999}, maxs[3] = {-99999, -99999, -99999};   // ClearBounds / AddPointToBounds
		for (int i = 0; i < w_; i++)
			for (int j = 0; j < h_; j++)
				for (int k = 0; k < 3; k++) {
					const float v = P(i, j)[k];
					if (v < mins[k]) mins[k] = v;
					if (v > maxs[k]) maxs[k] = v;
				}
		buildFacets();
		for (int k = 0; k < 3; k++) { out->bounds[k] = mins[k] - 1; out->bounds[3 + k] = maxs[k] + 1; }   // cm_patch.c:1190-1196
		out->planes.clear(); out->signbits.clear();
		for (const Plane &pl : planes_) {
			out->planes.insert(out->planes.end(), pl.p, pl.p + 4);
			out->signbits.push_back(pl.signbits);
		}
		out->facets = facets_;
	}

private:
	std::unique_ptr<float[][kMaxGrid][3]> pts_;
	std::unique_ptr<int[][kMaxGrid][2]> cellPlanes_;
	int w_ = 0, h_ = 0;
	bool wrapW_ = false, wrapH_ = false;
	std::vector<Plane> planes_;
	std::vector<PatchFacet> facets_;

	float *P(int i, int j) { return pts_[i][j]; }
	const float *P(int i, int j) const { return pts_[i][j]; }

	// ---------------- step 1: grid refinement -----------------
	void refineColumns() {
		// CM_SetGridWrapWidth, cm_patch.c:176-196: first and last column coincide within 0.1
		bool wrap = true;
		for (int j = 0; j < h_ && wrap; j++)
			for (int k = 0; k < 3; k++)
				if (!within(P(0, j)[k] - P(w_ - 1, j)[k], kPointEps)) { wrap = false; break; }
		wrapW_ = wrap;
		subdivideColumns();
		dropDuplicateColumns();
	}

	static bool curveIsBent(const float *a, const float *b, const float *c) {   // CM_NeedsSubdivision, cm_patch.c:89-112
		float delta[3];
		for (int i = 0; i < 3; i++) {
			const float chord = (float)(0.5 * (double)(a[i] + c[i]));
			const float curve = (float)(0.5 * (0.5 * (double)(a[i] + b[i]) + 0.5 * (double)(b[i] + c[i])));
			delta[i] = curve - chord;
		}
		const float len = sqrtf(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
		return len >= kSubdivide;
	}

	void removeColumn(int col) {
		for (int j = 0; j < h_; j++)
			for (int k = col + 1; k < w_; k++) memcpy(P(k - 1, j), P(k, j), 12);
		w_--;
	}

	void subdivideColumns() {                                       // CM_SubdivideGridColumns, cm_patch.c:207-270
		for (int i = 0; i < w_ - 2;) {
			bool bent = false;
			for (int j = 0; j < h_ && !bent; j++) bent = curveIsBent(P(i, j), P(i + 1, j), P(i + 2, j));
			if (!bent) {
				removeColumn(i + 1);      // the middle control column adds nothing: keep the chord
				i++;
				continue;
			}
			if (w_ + 2 > kMaxGrid) throw Drop("CM_SubdivideGridColumns: grid exceeds MAX_GRID_SIZE");
			for (int j = 0; j < h_; j++) {
				float a[3], b[3], c[3];
				memcpy(a, P(i, j), 12); memcpy(b, P(i + 1, j), 12); memcpy(c, P(i + 2, j), 12);
				for (int k = w_ - 1; k > i + 1; k--) memcpy(P(k + 2, j), P(k, j), 12);
				for (int k = 0; k < 3; k++) {                            // CM_Subdivide, cm_patch.c:123-131
					const float q1 = (float)(0.5 * (double)(a[k] + b[k]));
					const float q3 = (float)(0.5 * (double)(b[k] + c[k]));
					P(i + 1, j)[k] = q1;
					P(i + 3, j)[k] = q3;
					P(i + 2, j)[k] = (float)(0.5 * (double)(q1 + q3));
				}
			}
			w_ += 2;
			// the new middle column at i+1 is examined again
		}
	}

	void dropDuplicateColumns() {                                   // CM_RemoveDegenerateColumns, cm_patch.c:330-355
		for (int i = 0; i < w_ - 1; i++) {
			bool same = true;
			for (int j = 0; j < h_ && same; j++)
				for (int k = 0; k < 3; k++)
					if (!within(P(i, j)[k] - P(i + 1, j)[k], kPointEps)) { same = false; break; }
			if (!same) continue;
			removeColumn(i + 1);
			i--;
		}
	}

	void transpose() {                                              // CM_TransposeGrid, cm_patch.c:140-169
		const int n = w_ > h_ ? w_ : h_;
		for (int i = 0; i < n; i++)
			for (int j = i + 1; j < n; j++) {
				float t[3];
				memcpy(t, P(i, j), 12); memcpy(P(i, j), P(j, i), 12); memcpy(P(j, i), t, 12);
			}
		std::swap(w_, h_);
		std::swap(wrapW_, wrapH_);
	}

	// ---------------- step 2: planes -----------------
	static int signbitsOf(const float *n) { return (n[0] < 0 ? 1 : 0) | (n[1] < 0 ? 2 : 0) | (n[2] < 0 ? 4 : 0); }

	int addPlane(const float *pl) {
		if ((int)planes_.size() == kMaxPatchPlanes) throw Drop("MAX_PATCH_PLANES");
		Plane np;
		memcpy(np.p, pl, 16);
		np.signbits = signbitsOf(pl);
		planes_.push_back(np);
		return (int)planes_.size() - 1;
	}

	// plane through three points, merged into an existing one when all three lie within 0.1 of it.  CM_FindPlane
	int planeOfTriangle(const float *a, const float *b, const float *c) {
		float e1[3], e2[3], pl[4];
		for (int i = 0; i < 3; i++) { e1[i] = b[i] - a[i]; e2[i] = c[i] - a[i]; }
		crossD(e2, e1, pl);                                           // CM_PlaneFromPoints, cm_patch.c:62-75
		if (normalizeD(pl) == 0.0f) return -1;
		pl[3] = (float)dotD(a, pl);
		for (int i = 0; i < (int)planes_.size(); i++) {
			const float *q = planes_[i].p;
			if (dotF(pl, q) < 0) continue;
			if (!within(dotF(a, q) - q[3], kTriEps)) continue;
			if (!within(dotF(b, q) - q[3], kTriEps)) continue;
			if (!within(dotF(c, q) - q[3], kTriEps)) continue;
			return i;
		}
		return addPlane(pl);
	}

	// CM_PlaneEqual, cm_patch.c:377-405: same plane, or the same plane facing the other way
	static bool samePlane(const Plane &p, const float *pl, int *flipped) {
		auto near = [](float a, float b, double eps) { return fabs((double)(a - b)) < eps; };
		if (near(p.p[0], pl[0], kNormalEps) && near(p.p[1], pl[1], kNormalEps) && near(p.p[2], pl[2], kNormalEps) &&
		    near(p.p[3], pl[3], kDistEps)) { *flipped = 0; return true; }
		const float inv[4] = {-pl[0], -pl[1], -pl[2], -pl[3]};
		if (near(p.p[0], inv[0], kNormalEps) && near(p.p[1], inv[1], kNormalEps) && near(p.p[2], inv[2], kNormalEps) &&
		    near(p.p[3], inv[3], kDistEps)) { *flipped = 1; return true; }
		return false;
	}

	int planeLike(const float *pl, int *flipped) {                  // CM_FindPlane2, cm_patch.c:439-460
		for (int i = 0; i < (int)planes_.size(); i++)
			if (samePlane(planes_[i], pl, flipped)) return i;
		*flipped = 0;
		return addPlane(pl);
	}

	int sideOfPlane(const float *p, int plane) const {              // CM_PointOnPlaneSide: 0 front, 1 back, 2 on
		if (plane == -1) return 2;
		const float *q = planes_[plane].p;
		const double d = dotD(p, q) - (double)q[3];
		if (d > kTriEps) return 0;
		if (d < -kTriEps) return 1;
		return 2;
	}

	int cellPlane(int i, int j, int tri) const {                    // CM_GridPlane, cm_patch.c:545-560
		int p = cellPlanes_[i][j][tri];
		if (p != -1) return p;
		return cellPlanes_[i][j][!tri];
	}

	// a plane standing on one edge of cell (i,j), perpendicular to that cell's surface.  CM_EdgePlaneNum, cm_patch.c:568-642
	int rimPlane(int i, int j, int which) {
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
		default: throw Drop("CM_EdgePlaneNum: bad k");
		}
		const int p = cellPlane(i, j, tri);
		if (p == -1) return -1;
		float up[3];
		for (int k = 0; k < 3; k++) up[k] = a[k] + planes_[p].p[k] * 4.0f;   // VectorMA( p1, 4, normal, up )
		return reversed ? planeOfTriangle(b, a, up) : planeOfTriangle(a, b, up);
	}

	// ---------------- step 3/4: facets -----------------
	void orientBorders(PatchFacet &f, int i, int j, int which) {      // CM_SetBorderInward, cm_patch.c:651-722
		const float *corner[4];
		int n;
		if (which == -1) { corner[0] = P(i, j); corner[1] = P(i + 1, j); corner[2] = P(i + 1, j + 1); corner[3] = P(i, j + 1); n = 4; }
		else if (which == 0) { corner[0] = P(i, j); corner[1] = P(i + 1, j); corner[2] = P(i + 1, j + 1); n = 3; }
		else { corner[0] = P(i + 1, j + 1); corner[1] = P(i, j + 1); corner[2] = P(i, j); n = 3; }
		for (int k = 0; k < f.numBorders; k++) {
			int front = 0, back = 0;
			for (int l = 0; l < n; l++) {
				const int s = sideOfPlane(corner[l], f.borderPlanes[k]);
				if (s == 0) front++; else if (s == 1) back++;
			}
			if (front && !back) f.borderInward[k] = 1;
			else if (back && !front) f.borderInward[k] = 0;
			else if (!front && !back) f.borderPlanes[k] = -1;            // flat side border
			else f.borderInward[k] = 0;                                   // bisecting border ("mixed plane sides")
		}
	}

	// the facet's surface polygon cut by its borders; false when a border is missing or nothing is left
	bool facetWinding(const PatchFacet &f, bool skipSurfaceBorders, bool failOnMissing, Winding &w) const {
		const float *sp = planes_[f.surfacePlane].p;
		w = baseWinding(sp, sp[3]);
		for (int j = 0; j < f.numBorders; j++) {
			if (f.borderPlanes[j] == -1) { if (failOnMissing) return false; continue; }
			if (skipSurfaceBorders && f.borderPlanes[j] == f.surfacePlane) continue;
			float pl[4];
			memcpy(pl, planes_[f.borderPlanes[j]].p, 16);
			if (!f.borderInward[j]) {
				for (int k = 0; k < 3; k++) pl[k] = 0.0f - pl[k];          // VectorSubtract( vec3_origin, plane, plane )
				pl[3] = -pl[3];
			}
			if (!chopWinding(w, pl, pl[3], 0.1f)) return false;
		}
		return true;
	}

	bool facetIsBounded(const PatchFacet &f) const {                  // CM_ValidateFacet, cm_patch.c:731-777
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

	bool haveBorderLike(const PatchFacet &f, const float *pl, int *flipped) const {
		for (int i = 0; i < f.numBorders; i++)
			if (samePlane(planes_[f.borderPlanes[i]], pl, flipped)) return true;
		return false;
	}

	void addBevels(PatchFacet &f) {                                    // CM_AddFacetBevels, cm_patch.c:785-952
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
				if (samePlane(planes_[f.surfacePlane], pl, &flipped)) continue;
				if (haveBorderLike(f, pl, &flipped)) continue;
				if (f.numBorders >= kMaxFacetBorders) continue;            // "too many bevels"
				f.borderPlanes[f.numBorders] = planeLike(pl, &flipped);
				f.borderNoAdjust[f.numBorders] = 0;
				f.borderInward[f.numBorders] = flipped;
				f.numBorders++;
			}
		}
		// slanted bevels along the non-axial edges of the facet polygon
		const int np = (int)w.size();
		for (int j = 0; j < np; j++) {
			const int k = (j + 1) % np;
			float edge[3];
			for (int c = 0; c < 3; c++) edge[c] = (float)((double)w[j].v[c] - (double)w[k].v[c]);
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
					pl[3] = (float)dotD(w[j].v, pl);
					bool behind = true;                                          // every polygon point must be behind the bevel
					for (int l = 0; l < np; l++) {
						if (dotD(w[l].v, pl) - (double)pl[3] > 0.1) { behind = false; break; }
					}
					if (!behind) continue;
					if (samePlane(planes_[f.surfacePlane], pl, &flipped)) continue;
					if (haveBorderLike(f, pl, &flipped)) continue;
					if (f.numBorders >= kMaxFacetBorders) continue;
					const int slot = f.numBorders;
					f.borderPlanes[slot] = planeLike(pl, &flipped);
					f.borderNoAdjust[slot] = 0;
					f.borderInward[slot] = flipped;
					// the bevel must leave some of the polygon: otherwise the slot is not claimed
					// (the plane stays in the list, as in the reference)
					float cut[4];
					memcpy(cut, planes_[f.borderPlanes[slot]].p, 16);
					if (!f.borderInward[slot]) { cut[0] = -cut[0]; cut[1] = -cut[1]; cut[2] = -cut[2]; cut[3] = -cut[3]; }
					Winding w2 = w;
					if (!chopWinding(w2, cut, cut[3], 0.1f)) continue;
					f.numBorders++;
				}
			}
		}
		// the back side
		if (f.numBorders >= kMaxFacetBorders) return;
		f.borderPlanes[f.numBorders] = f.surfacePlane;
		f.borderNoAdjust[f.numBorders] = 0;
		f.borderInward[f.numBorders] = 1;
		f.numBorders++;
	}

	void keepIfBounded(PatchFacet &f) {
		if (facetIsBounded(f)) {
			addBevels(f);
			if ((int)facets_.size() == kMaxFacets) throw Drop("MAX_FACETS");
			facets_.push_back(f);
		}
	}

	void buildFacets() {                                               // CM_PatchCollideFromGrid, cm_patch.c:969-1120
		planes_.clear();
		facets_.clear();
		for (int i = 0; i < w_ - 1; i++)
			for (int j = 0; j < h_ - 1; j++) {
				cellPlanes_[i][j][0] = planeOfTriangle(P(i, j), P(i + 1, j), P(i + 1, j + 1));
				cellPlanes_[i][j][1] = planeOfTriangle(P(i + 1, j + 1), P(i, j + 1), P(i, j));
			}
		for (int i = 0; i < w_ - 1; i++) {
			for (int j = 0; j < h_ - 1; j++) {
				const int tri0 = cellPlanes_[i][j][0], tri1 = cellPlanes_[i][j][1];
				// neighbour across each edge (wrapping when the patch closes on itself), else a rim plane
				int top = -1, bottom = -1, left = -1, right = -1;
				if (j > 0) top = cellPlanes_[i][j - 1][1]; else if (wrapH_) top = cellPlanes_[i][h_ - 2][1];
				const int topSame = (top == tri0);
				if (top == -1 || topSame) top = rimPlane(i, j, 0);
				if (j < h_ - 2) bottom = cellPlanes_[i][j + 1][0]; else if (wrapH_) bottom = cellPlanes_[i][0][0];
				const int bottomSame = (bottom == tri1);
				if (bottom == -1 || bottomSame) bottom = rimPlane(i, j, 2);
				if (i > 0) left = cellPlanes_[i - 1][j][0]; else if (wrapW_) left = cellPlanes_[w_ - 2][j][0];
				const int leftSame = (left == tri1);
				if (left == -1 || leftSame) left = rimPlane(i, j, 3);
				if (i < w_ - 2) right = cellPlanes_[i + 1][j][1]; else if (wrapW_) right = cellPlanes_[0][j][1];
				const int rightSame = (right == tri0);
				if (right == -1 || rightSame) right = rimPlane(i, j, 1);

				if ((int)facets_.size() == kMaxFacets) throw Drop("MAX_FACETS");
				PatchFacet f;
				memset(&f, 0, sizeof(f));
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
					orientBorders(f, i, j, 0);
					keepIfBounded(f);

					if ((int)facets_.size() == kMaxFacets) throw Drop("MAX_FACETS");
					memset(&f, 0, sizeof(f));
					f.surfacePlane = tri1;
					f.numBorders = 3;
					f.borderPlanes[0] = bottom; f.borderNoAdjust[0] = bottomSame;
					f.borderPlanes[1] = left;   f.borderNoAdjust[1] = leftSame;
					f.borderPlanes[2] = tri0;
					if (f.borderPlanes[2] == -1) {
						f.borderPlanes[2] = top;
						if (f.borderPlanes[2] == -1) f.borderPlanes[2] = rimPlane(i, j, 5);
					}
					orientBorders(f, i, j, 1);
					keepIfBounded(f);
				}
			}
		}
	}
};

}  // namespace

bool generatePatchCollide(int width, int height, const float *points, PatchCollide *out, char *err, size_t errlen) {
	auto fail = [&](const char *msg) {
		if (err && errlen) snprintf(err, errlen, "%s", msg);
		return false;
	};
	if (width <= 2 || height <= 2 || !points) return fail("CM_GeneratePatchFacets: bad parameters");
	if (!(width & 1) || !(height & 1)) return fail("CM_GeneratePatchFacets: even sizes are invalid for quadratic meshes");
	if (width > kMaxGrid || height > kMaxGrid) return fail("CM_GeneratePatchFacets: source is > MAX_GRID_SIZE");
	try {
		Builder b;
		b.run(width, height, points, out);
	} catch (const Drop &d) {
		return fail(d.what());
	}
	return true;
}

}  // namespace cmgpu
