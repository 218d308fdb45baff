// cm_kernels.cuh -- sm_100a device code of the batched clip-map trace engine.
//
// One thread owns one work item (a CM_BoxTrace / CM_TransformedBoxTrace /
// CM_PointContents call of the reference, code/qcommon/cm_trace.c, cm_test.c,
// cm_patch.c) and walks the flattened BSP with an explicit per-thread stack.
// The map is read-only, index-based and packed in 16-byte records so that every
// node / leaf / brush-box / side-plane fetch is one LDG.128 that hits L1/L2.
//
// Arithmetic contract (SURVEY.md 8a'): the reference mixes float and double in
// the same expressions and is compiled without FMA contraction.  This TU is
// built with -fmad=false; every float op below is therefore a separately
// rounded IEEE op, exactly like the SSE2 code of the reference.  Double dot
// products of float inputs use explicit fma(): each float*float product is
// exact in double, so fma(a,b,c) == (a*b)+c bit for bit and only saves issue
// slots.  Divisions are IEEE (no fast-math), conversions round-to-nearest.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace cmgpu {

constexpr int    kMaxDepth       = 64;      // traversal stack entries per item (CMGPU_ERR_LIMIT beyond)
constexpr int    kMaxPosLeafs    = 1024;    // MAX_POSITION_LEAFS, cm_trace.c:190
constexpr int    kBoxModelHandle = 1023;    // BOX_MODEL_HANDLE, cm_local.h:28
constexpr double kClipEps        = 0.125;   // SURFACE_CLIP_EPSILON, cm_local.h:177
constexpr float  kBoundsEps      = 0.25f;   // BOUNDS_CLIP_EPSILON, cm_test.c:474

// ---- device-resident map (structure of arrays, 16-byte records) -------------
struct DevMap {
	const int4   *nodes;        // child0, child1, plane type (0..2 axial, 3 general), float bits of dist
	const float4 *nodePlanes;   // normal.xyz, dist (read only for type 3)
	const int4   *leafs;        // firstLeafBrush, numLeafBrushes, firstLeafSurface, numLeafSurfaces
	const int    *leafbrushes;
	const int    *leafsurfaces;
	const float4 *brushBox;     // [2b] mins.xyz + contents bits, [2b+1] maxs.xyz + firstSide bits
	const int    *brushNumSides;
	const float4 *sidePlanes;   // normal.xyz, dist per brush side (plane copied next to its side)
	const int2   *sideMeta;     // surfaceFlags, type | signbits << 8
	const int4   *modelLeafs;   // pseudo leaf per inline model
	const int    *surf2patch;
	const int4   *patchInfo;    // [2p] surfaceFlags, contents, firstPlane, numPlanes ; [2p+1] firstFacet, numFacets, 0, 0
	const float4 *patchBox;     // [2p] mins, [2p+1] maxs
	const float4 *patchPlanes;
	const int4   *facets;       // surfacePlane, numBorders, firstBorder, 0
	const int    *borders;      // plane | inward << 31
	int numNodes;
	int numModels;
	int boxBrush;
	int boxLeafBrush;
	int noCurves;
	int playerCurveClip;
};

struct Counters {
	unsigned long long traces, nodes, leafs, brushRefs, brushBounds, brushSides, crossings, splits,
	                   patches, patchPlanes, facets;
};

// ---- one batch ----------------------------------------------------------------
struct TraceBatch {
	int n;
	const float *starts, *ends;          // [n][3]
	const float *mins, *maxs;            // [1][3] or [n][3]; may be null
	int hullStride;
	const int *models;                   // may be null (world)
	const int *masks;
	int maskStride;
	const float *origins;                // null: untransformed (CM_BoxTrace)
	const float *rotations;              // [*][9]
	const int *rotIndex;                 // -1: not rotated
	const float *boxmins, *boxmaxs;      // temp-box geometry per item; may be null
	uint2 *out;                          // 56-byte trace records viewed as 7 x 8 bytes
	int *hitIds;                         // may be null
	int *status;                         // device status word
	Counters *counters;                  // only read by the counting variant
};

struct PointBatch {
	int n;
	const float *points;
	const int *models;
	const float *origins;
	const float *rotations;
	const int *rotIndex;
	const float *boxmins, *boxmaxs;
	int *out;
};

// ---- per-item working set (registers) ----------------------------------------------
struct Work {
	float sx, sy, sz, ex, ey, ez;        // symmetrised sweep end points (tw.start / tw.end)
	float n0x, n0y, n0z, p0x, p0y, p0z;  // tw.size[0], tw.size[1]
	float extx, exty, extz;              // tw.extents
	float lox, loy, loz, hix, hiy, hiz;  // tw.bounds
	int   mask;
	bool  isPoint;
	bool  isBox;                         // model == 1023
	float bmnx, bmny, bmnz, bmxx, bmxy, bmxz;   // temp-box mins/maxs
	// running result (tw.trace)
	int   allsolid, startsolid;
	float fraction;
	float pnx, pny, pnz, pd;
	int   ptypesign;                     // plane.type | plane.signbits << 8
	int   surfaceFlags, contents;
	int   hitId;
	// visit counts (counting variant only)
	unsigned cNodes, cLeafs, cBrushRefs, cBrushBounds, cBrushSides, cCross, cSplits, cPatches, cPatchPlanes, cFacets;
};

template <bool COUNT> struct Cnt {
	static __device__ __forceinline__ void add(unsigned &c, unsigned v = 1) { if (COUNT) c += v; }
};

__device__ __forceinline__ float sel3(float x, float y, float z, int i) {
	return i == 0 ? x : (i == 1 ? y : z);
}

// DotProduct (q_shared.h:489): ((a0*b0 + a1*b1) + a2*b2), three roundings + two, no contraction
__device__ __forceinline__ float dotf(float ax, float ay, float az, float bx, float by, float bz) {
	float s = ax * bx;
	s = s + ay * by;
	s = s + az * bz;
	return s;
}

// DotProductDP (cm_local.h:31) on float inputs: products are exact in double
__device__ __forceinline__ double dotd(double ax, double ay, double az, float bx, float by, float bz) {
	double s = ax * (double)bx;
	s = fma(ay, (double)by, s);
	s = fma(az, (double)bz, s);
	return s;
}

// temp-box side s (CM_InitBoxHull / CM_TempBoxModel, cm_load.c:873-949):
// sides +x,-x,+y,-y,+z,-z carry dist maxs[0], -mins[0], maxs[1], -mins[1], maxs[2], -mins[2]
__device__ __forceinline__ float boxSideDist(const Work &w, int s) {
	int axis = s >> 1;
	if (s & 1) return -sel3(w.bmnx, w.bmny, w.bmnz, axis);
	return sel3(w.bmxx, w.bmxy, w.bmxz, axis);
}

// ---- CM_TraceThroughBrush, cm_trace.c:261-363 ------------------------------------------
template <bool COUNT>
__device__ void sweepBrush(const DevMap &m, Work &w, int b, int firstSide, int nsides, int contents, bool boxBrush) {
	if (nsides == 0) return;
	float enter = -1.0f, leave = 1.0f;
	int lead = -1;
	bool endOut = false, startOut = false;
	const double sx = w.sx, sy = w.sy, sz = w.sz, ex = w.ex, ey = w.ey, ez = w.ez;

	for (int s = 0; s < nsides; s++) {
		float4 pl = __ldg(&m.sidePlanes[firstSide + s]);
		if (boxBrush) pl.w = boxSideDist(w, s);
		Cnt<COUNT>::add(w.cBrushSides);
		// tw.offsets[signbits]: the box corner that touches the plane first
		const float cx = pl.x < 0.0f ? w.p0x : w.n0x;
		const float cy = pl.y < 0.0f ? w.p0y : w.n0y;
		const float cz = pl.z < 0.0f ? w.p0z : w.n0z;
		const double dist = (double)pl.w - dotd((double)cx, (double)cy, (double)cz, pl.x, pl.y, pl.z);
		const double d1 = dotd(sx, sy, sz, pl.x, pl.y, pl.z) - dist;
		const double d2 = dotd(ex, ey, ez, pl.x, pl.y, pl.z) - dist;
		if (d2 > 0) endOut = true;
		if (d1 > 0) startOut = true;
		if (d1 > 0 && (d2 >= kClipEps || d2 >= d1)) return;     // wholly in front of one face
		if (d1 <= 0 && d2 <= 0) continue;                        // face not crossed
		Cnt<COUNT>::add(w.cCross);
		if (d1 > d2) {
			float f = (float)((d1 - kClipEps) / (d1 - d2));
			if (f < 0) f = 0;
			if (f > enter) { enter = f; lead = s; }
		} else {
			float f = (float)((d1 + kClipEps) / (d1 - d2));
			if (f > 1) f = 1;
			if (f < leave) leave = f;
		}
	}

	if (!startOut) {
		w.startsolid = 1;
		if (!endOut) {
			w.allsolid = 1;
			w.fraction = 0;
			w.contents = contents;
			w.hitId = b;
		}
		return;
	}
	if (enter < leave && enter > -1 && enter < w.fraction) {
		if (enter < 0) enter = 0;
		w.fraction = enter;
		if (lead >= 0) {
			float4 pl = __ldg(&m.sidePlanes[firstSide + lead]);
			if (boxBrush) pl.w = boxSideDist(w, lead);
			const int2 meta = __ldg(&m.sideMeta[firstSide + lead]);
			w.pnx = pl.x; w.pny = pl.y; w.pnz = pl.z; w.pd = pl.w;
			w.ptypesign = meta.y;
			w.surfaceFlags = meta.x;
		}
		w.contents = contents;
		w.hitId = b;
	}
}

// ---- patch facets, cm_patch.c:1220-1531 ------------------------------------------------

// CM_CheckFacetPlane, cm_patch.c:1321-1360
__device__ __forceinline__ bool clipFacetPlane(float px, float py, float pz, float pw, const Work &w,
                                               float &enter, float &leave, bool &hit) {
	hit = false;
	const float d1 = dotf(w.sx, w.sy, w.sz, px, py, pz) - pw;
	const float d2 = dotf(w.ex, w.ey, w.ez, px, py, pz) - pw;
	if (d1 > 0 && (d2 >= 0.125f || d2 >= d1)) return false;
	if (d1 <= 0 && d2 <= 0) return true;
	if (d1 > d2) {
		float f = (float)(((double)d1 - kClipEps) / (double)(d1 - d2));
		if (f < 0) f = 0;
		if (f > enter) { enter = f; hit = true; }
	} else {
		float f = (float)(((double)d1 + kClipEps) / (double)(d1 - d2));
		if (f > 1) f = 1;
		if (f < leave) leave = f;
	}
	return true;
}

// relation of the point sweep to one patch plane (the two tables of
// CM_TracePointThroughPatchCollide, cm_patch.c:1239-1261, evaluated on demand)
__device__ __forceinline__ void pointPlaneRelation(const float4 pl, const Work &w, bool &front, float &cross) {
	const float cx = pl.x < 0.0f ? w.p0x : w.n0x;
	const float cy = pl.y < 0.0f ? w.p0y : w.n0y;
	const float cz = pl.z < 0.0f ? w.p0z : w.n0z;
	const float off = dotf(cx, cy, cz, pl.x, pl.y, pl.z);
	const float d1 = dotf(w.sx, w.sy, w.sz, pl.x, pl.y, pl.z) - pl.w + off;
	const float d2 = dotf(w.ex, w.ey, w.ez, pl.x, pl.y, pl.z) - pl.w + off;
	front = !(d1 <= 0);
	if (d1 == d2) {
		cross = 99999.0f;
	} else {
		cross = d1 / (d1 - d2);
		if (cross <= 0) cross = 99999.0f;
	}
}

template <bool COUNT>
__device__ void sweepPatchPoint(const DevMap &m, Work &w, int firstPlane, int firstFacet, int numFacets) {
	if (!m.playerCurveClip || !w.isPoint) return;
	for (int f = 0; f < numFacets; f++) {
		const int4 fc = __ldg(&m.facets[firstFacet + f]);
		Cnt<COUNT>::add(w.cFacets);
		const float4 sp = __ldg(&m.patchPlanes[firstPlane + fc.x]);
		bool front; float t;
		pointPlaneRelation(sp, w, front, t);
		Cnt<COUNT>::add(w.cPatchPlanes);
		if (!front) continue;
		if (t < 0) continue;
		if (t > w.fraction) continue;
		int j;
		for (j = 0; j < fc.y; j++) {
			const int bd = __ldg(&m.borders[fc.z + j]);
			const float4 bp = __ldg(&m.patchPlanes[firstPlane + (bd & 0x7fffffff)]);
			bool bfront; float bt;
			pointPlaneRelation(bp, w, bfront, bt);
			Cnt<COUNT>::add(w.cPatchPlanes);
			const bool inward = bd < 0;
			if (bfront != inward) {
				if (bt > t) break;
			} else {
				if (bt < t) break;
			}
		}
		if (j != fc.y) continue;
		// hit: recompute with the 1/8 push-off, mixed widths as cm_patch.c:1300-1303
		const float cx = sp.x < 0.0f ? w.p0x : w.n0x;
		const float cy = sp.y < 0.0f ? w.p0y : w.n0y;
		const float cz = sp.z < 0.0f ? w.p0z : w.n0z;
		const float off = dotf(cx, cy, cz, sp.x, sp.y, sp.z);
		const float d1 = dotf(w.sx, w.sy, w.sz, sp.x, sp.y, sp.z) - sp.w + off;
		const float d2 = dotf(w.ex, w.ey, w.ez, sp.x, sp.y, sp.z) - sp.w + off;
		float fr = (float)(((double)d1 - kClipEps) / (double)(d1 - d2));
		if (fr < 0) fr = 0;
		w.fraction = fr;
		w.pnx = sp.x; w.pny = sp.y; w.pnz = sp.z; w.pd = sp.w;   // type/signbits stay stale
	}
}

// CM_TraceThroughPatchCollide, cm_patch.c:1368-1463
template <bool COUNT>
__device__ void sweepPatch(const DevMap &m, Work &w, int p) {
	const float4 bmn = __ldg(&m.patchBox[2 * p]);
	const float4 bmx = __ldg(&m.patchBox[2 * p + 1]);
	if (w.hix < bmn.x - kBoundsEps || w.hiy < bmn.y - kBoundsEps || w.hiz < bmn.z - kBoundsEps ||
	    w.lox > bmx.x + kBoundsEps || w.loy > bmx.y + kBoundsEps || w.loz > bmx.z + kBoundsEps) return;
	Cnt<COUNT>::add(w.cPatches);
	const int4 pa = __ldg(&m.patchInfo[2 * p]);
	const int4 pb = __ldg(&m.patchInfo[2 * p + 1]);
	const int firstPlane = pa.z, firstFacet = pb.x, numFacets = pb.y;
	if (w.isPoint) {
		sweepPatchPoint<COUNT>(m, w, firstPlane, firstFacet, numFacets);
		return;
	}
	float bx = 0, by = 0, bz = 0, bw = 0;
	for (int f = 0; f < numFacets; f++) {
		const int4 fc = __ldg(&m.facets[firstFacet + f]);
		Cnt<COUNT>::add(w.cFacets);
		float enter = -1.0f, leave = 1.0f;
		int hitnum = -1;
		bool hit;
		const float4 sp = __ldg(&m.patchPlanes[firstPlane + fc.x]);
		{
			const float cx = sp.x < 0.0f ? w.p0x : w.n0x;
			const float cy = sp.y < 0.0f ? w.p0y : w.n0y;
			const float cz = sp.z < 0.0f ? w.p0z : w.n0z;
			const float pw = sp.w - dotf(cx, cy, cz, sp.x, sp.y, sp.z);
			if (!clipFacetPlane(sp.x, sp.y, sp.z, pw, w, enter, leave, hit)) continue;
			if (hit) { bx = sp.x; by = sp.y; bz = sp.z; bw = pw; }
		}
		int j;
		for (j = 0; j < fc.y; j++) {
			const int bd = __ldg(&m.borders[fc.z + j]);
			const float4 bp = __ldg(&m.patchPlanes[firstPlane + (bd & 0x7fffffff)]);
			Cnt<COUNT>::add(w.cPatchPlanes);
			// corner by the ORIGINAL plane's sign bits, dotted with the possibly flipped plane
			const float cx = bp.x < 0.0f ? w.p0x : w.n0x;
			const float cy = bp.y < 0.0f ? w.p0y : w.n0y;
			const float cz = bp.z < 0.0f ? w.p0z : w.n0z;
			float px = bp.x, py = bp.y, pz = bp.z, pw = bp.w;
			if (bd < 0) { px = -px; py = -py; pz = -pz; pw = -pw; }
			const float off = dotf(cx, cy, cz, px, py, pz);
			pw = pw + fabsf(off);        // double add of two floats rounded to float == float add
			if (!clipFacetPlane(px, py, pz, pw, w, enter, leave, hit)) break;
			if (hit) { hitnum = j; bx = px; by = py; bz = pz; bw = pw; }
		}
		if (j < fc.y) continue;
		if (hitnum == fc.y - 1) continue;    // never clip against the back side
		if (enter < leave && enter >= 0 && enter < w.fraction) {
			w.fraction = enter;
			w.pnx = bx; w.pny = by; w.pnz = bz; w.pd = bw;
		}
	}
}

// CM_PositionTestInPatchCollide, cm_patch.c:1479-1531
template <bool COUNT>
__device__ bool boxInPatch(const DevMap &m, Work &w, int p) {
	if (w.isPoint) return false;
	const int4 pa = __ldg(&m.patchInfo[2 * p]);
	const int4 pb = __ldg(&m.patchInfo[2 * p + 1]);
	const int firstPlane = pa.z, firstFacet = pb.x, numFacets = pb.y;
	for (int f = 0; f < numFacets; f++) {
		const int4 fc = __ldg(&m.facets[firstFacet + f]);
		Cnt<COUNT>::add(w.cFacets);
		const float4 sp = __ldg(&m.patchPlanes[firstPlane + fc.x]);
		{
			const float cx = sp.x < 0.0f ? w.p0x : w.n0x;
			const float cy = sp.y < 0.0f ? w.p0y : w.n0y;
			const float cz = sp.z < 0.0f ? w.p0z : w.n0z;
			const float pw = sp.w - dotf(cx, cy, cz, sp.x, sp.y, sp.z);
			if (dotf(sp.x, sp.y, sp.z, w.sx, w.sy, w.sz) - pw > 0.0f) continue;
		}
		int j;
		for (j = 0; j < fc.y; j++) {
			const int bd = __ldg(&m.borders[fc.z + j]);
			const float4 bp = __ldg(&m.patchPlanes[firstPlane + (bd & 0x7fffffff)]);
			const float cx = bp.x < 0.0f ? w.p0x : w.n0x;
			const float cy = bp.y < 0.0f ? w.p0y : w.n0y;
			const float cz = bp.z < 0.0f ? w.p0z : w.n0z;
			float px = bp.x, py = bp.y, pz = bp.z, pw = bp.w;
			if (bd < 0) { px = -px; py = -py; pz = -pz; pw = -pw; }
			const float off = dotf(cx, cy, cz, px, py, pz);
			pw = pw + fabsf(off);
			if (dotf(px, py, pz, w.sx, w.sy, w.sz) - pw > 0.0f) break;
		}
		if (j < fc.y) continue;
		return true;
	}
	return false;
}

// ---- leaf level ----------------------------------------------------------------------------

// CM_TraceThroughLeaf + CM_TraceThroughPatch, cm_trace.c:370-427,241-254
template <bool COUNT>
__device__ void sweepLeaf(const DevMap &m, Work &w, const int4 leaf) {
	Cnt<COUNT>::add(w.cLeafs);
	for (int k = 0; k < leaf.y; k++) {
		const int b = __ldg(&m.leafbrushes[leaf.x + k]);
		Cnt<COUNT>::add(w.cBrushRefs);
		float4 bmn = __ldg(&m.brushBox[2 * b]);
		const int contents = __float_as_int(bmn.w);
		if (!(contents & w.mask)) continue;
		float4 bmx = __ldg(&m.brushBox[2 * b + 1]);
		const bool boxBrush = w.isBox && b == m.boxBrush;
		if (boxBrush) {
			bmn.x = w.bmnx; bmn.y = w.bmny; bmn.z = w.bmnz;
			bmx.x = w.bmxx; bmx.y = w.bmxy; bmx.z = w.bmxz;
		}
		Cnt<COUNT>::add(w.cBrushBounds);
		// CM_BoundsIntersect, cm_test.c:481-494
		if (w.hix < bmn.x - kBoundsEps || w.hiy < bmn.y - kBoundsEps || w.hiz < bmn.z - kBoundsEps ||
		    w.lox > bmx.x + kBoundsEps || w.loy > bmx.y + kBoundsEps || w.loz > bmx.z + kBoundsEps) continue;
		sweepBrush<COUNT>(m, w, b, __float_as_int(bmx.w), __ldg(&m.brushNumSides[b]), contents, boxBrush);
		if (w.fraction == 0.0f) return;
	}
	if (m.noCurves) return;
	for (int k = 0; k < leaf.w; k++) {
		const int p = __ldg(&m.surf2patch[__ldg(&m.leafsurfaces[leaf.z + k])]);
		if (p < 0) continue;
		const int4 pa = __ldg(&m.patchInfo[2 * p]);
		if (!(pa.y & w.mask)) continue;
		const float before = w.fraction;
		sweepPatch<COUNT>(m, w, p);
		if (w.fraction < before) {
			w.surfaceFlags = pa.x;
			w.contents = pa.y;
			w.hitId = -2 - p;
		}
		if (w.fraction == 0.0f) return;
	}
}

// CM_TestBoxInBrush, cm_trace.c:83-123
template <bool COUNT>
__device__ void boxInBrush(const DevMap &m, Work &w, int b, float4 bmn, float4 bmx, int contents, bool boxBrush) {
	const int nsides = __ldg(&m.brushNumSides[b]);
	if (!nsides) return;
	Cnt<COUNT>::add(w.cBrushBounds);
	if (w.lox > bmx.x || w.loy > bmx.y || w.loz > bmx.z || w.hix < bmn.x || w.hiy < bmn.y || w.hiz < bmn.z) return;
	const int firstSide = __float_as_int(bmx.w);
	for (int s = 6; s < nsides; s++) {
		float4 pl = __ldg(&m.sidePlanes[firstSide + s]);
		if (boxBrush) pl.w = boxSideDist(w, s);
		Cnt<COUNT>::add(w.cBrushSides);
		const float cx = pl.x < 0.0f ? w.p0x : w.n0x;
		const float cy = pl.y < 0.0f ? w.p0y : w.n0y;
		const float cz = pl.z < 0.0f ? w.p0z : w.n0z;
		const float distf = pl.w - dotf(cx, cy, cz, pl.x, pl.y, pl.z);    // all-float (cm_trace.c:111)
		const double d1 = dotd((double)w.sx, (double)w.sy, (double)w.sz, pl.x, pl.y, pl.z) - (double)distf;
		if (d1 > 0) return;
	}
	w.startsolid = w.allsolid = 1;
	w.fraction = 0;
	w.contents = contents;
	w.hitId = b;
}

// CM_TestInLeaf, cm_trace.c:130-183
template <bool COUNT>
__device__ void testLeaf(const DevMap &m, Work &w, const int4 leaf) {
	Cnt<COUNT>::add(w.cLeafs);
	for (int k = 0; k < leaf.y; k++) {
		const int b = __ldg(&m.leafbrushes[leaf.x + k]);
		Cnt<COUNT>::add(w.cBrushRefs);
		float4 bmn = __ldg(&m.brushBox[2 * b]);
		const int contents = __float_as_int(bmn.w);
		if (!(contents & w.mask)) continue;
		float4 bmx = __ldg(&m.brushBox[2 * b + 1]);
		const bool boxBrush = w.isBox && b == m.boxBrush;
		if (boxBrush) {
			bmn.x = w.bmnx; bmn.y = w.bmny; bmn.z = w.bmnz;
			bmx.x = w.bmxx; bmx.y = w.bmxy; bmx.z = w.bmxz;
		}
		boxInBrush<COUNT>(m, w, b, bmn, bmx, contents, boxBrush);
		if (w.allsolid) return;
	}
	if (m.noCurves) return;
	for (int k = 0; k < leaf.w; k++) {
		const int p = __ldg(&m.surf2patch[__ldg(&m.leafsurfaces[leaf.z + k])]);
		if (p < 0) continue;
		const int4 pa = __ldg(&m.patchInfo[2 * p]);
		if (!(pa.y & w.mask)) continue;
		if (boxInPatch<COUNT>(m, w, p)) {
			w.startsolid = w.allsolid = 1;
			w.fraction = 0;
			w.contents = pa.y;
			w.hitId = -2 - p;
			return;
		}
	}
}

// ---- tree level ------------------------------------------------------------------------------

// BoxOnPlaneSide, q_math.c:724-758
__device__ __forceinline__ int boxPlaneSide(const DevMap &m, int num, const int4 nd,
                                            float lox, float loy, float loz, float hix, float hiy, float hiz) {
	const float dist = __int_as_float(nd.w);
	if (nd.z < 3) {
		if (dist <= sel3(lox, loy, loz, nd.z)) return 1;
		if (dist >= sel3(hix, hiy, hiz, nd.z)) return 2;
		return 3;
	}
	const float4 pl = __ldg(&m.nodePlanes[num]);
	// accumulate in component order from 0, as the reference's loop does
	float a = 0.0f, b = 0.0f;     // dist[0], dist[1]
	if (pl.x < 0.0f) { b = b + pl.x * hix; a = a + pl.x * lox; } else { a = a + pl.x * hix; b = b + pl.x * lox; }
	if (pl.y < 0.0f) { b = b + pl.y * hiy; a = a + pl.y * loy; } else { a = a + pl.y * hiy; b = b + pl.y * loy; }
	if (pl.z < 0.0f) { b = b + pl.z * hiz; a = a + pl.z * loz; } else { a = a + pl.z * hiz; b = b + pl.z * loz; }
	int s = 0;
	if (a >= dist) s = 1;
	if (b < dist) s |= 2;
	return s;
}

// CM_PositionTest, cm_trace.c:191-226 (+ CM_BoxLeafnums_r / CM_StoreLeafs, cm_test.c:131-156,73-88).
// The reference first lists <= 1024 leafs and then tests them in list order;
// testing each leaf when it is discovered is the same sequence of tests.
template <bool COUNT>
__device__ void positionTestWorld(const DevMap &m, Work &w, int *status) {
	const float lox = (w.sx + w.n0x) - 1.0f, loy = (w.sy + w.n0y) - 1.0f, loz = (w.sz + w.n0z) - 1.0f;
	const float hix = (w.sx + w.p0x) + 1.0f, hiy = (w.sy + w.p0y) + 1.0f, hiz = (w.sz + w.p0z) + 1.0f;
	int stack[kMaxDepth];
	int sp = 0, found = 0, num = 0;
	for (;;) {
		if (num < 0) {
			if (found >= kMaxPosLeafs) return;
			found++;
			testLeaf<COUNT>(m, w, __ldg(&m.leafs[-1 - num]));
			if (w.allsolid) return;
			if (!sp) return;
			num = stack[--sp];
			continue;
		}
		const int4 nd = __ldg(&m.nodes[num]);
		Cnt<COUNT>::add(w.cNodes);
		const int s = boxPlaneSide(m, num, nd, lox, loy, loz, hix, hiy, hiz);
		if (s == 1) {
			num = nd.x;
		} else if (s == 2) {
			num = nd.y;
		} else {
			if (sp < kMaxDepth) stack[sp++] = nd.y; else atomicOr(status, 1);
			num = nd.x;
		}
	}
}

struct Seg { int num; float f1, f2, ax, ay, az, bx, by, bz; };

// CM_TraceThroughTree, cm_trace.c:439-538, as a loop with an explicit stack of
// pending far halves.  Split points are carried (computed from the CURRENT
// segment in float), never recomputed from the root segment.
template <bool COUNT>
__device__ void sweepWorld(const DevMap &m, Work &w, int *status) {
	Seg stack[kMaxDepth];
	int sp = 0;
	Seg cur;
	cur.num = 0; cur.f1 = 0.0f; cur.f2 = 1.0f;
	cur.ax = w.sx; cur.ay = w.sy; cur.az = w.sz;
	cur.bx = w.ex; cur.by = w.ey; cur.bz = w.ez;

	for (;;) {
		bool done = false;
		if (w.fraction <= cur.f1) {
			done = true;
		} else if (cur.num < 0) {
			sweepLeaf<COUNT>(m, w, __ldg(&m.leafs[-1 - cur.num]));
			done = true;
		}
		if (done) {
			if (!sp) return;
			cur = stack[--sp];
			continue;
		}

		const int4 nd = __ldg(&m.nodes[cur.num]);
		Cnt<COUNT>::add(w.cNodes);
		double t1, t2, off;
		if (nd.z < 3) {
			const float dist = __int_as_float(nd.w);
			t1 = (double)(sel3(cur.ax, cur.ay, cur.az, nd.z) - dist);   // float subtract, then widened
			t2 = (double)(sel3(cur.bx, cur.by, cur.bz, nd.z) - dist);
			off = (double)sel3(w.extx, w.exty, w.extz, nd.z);
		} else {
			const float4 pl = __ldg(&m.nodePlanes[cur.num]);
			t1 = dotd((double)pl.x, (double)pl.y, (double)pl.z, cur.ax, cur.ay, cur.az) - (double)pl.w;
			t2 = dotd((double)pl.x, (double)pl.y, (double)pl.z, cur.bx, cur.by, cur.bz) - (double)pl.w;
			off = w.isPoint ? 0.0 : 2048.0;
		}
		if (t1 >= off + 1 && t2 >= off + 1) { cur.num = nd.x; continue; }
		if (t1 < -off - 1 && t2 < -off - 1) { cur.num = nd.y; continue; }

		int side;
		float fnear, ffar;
		if (t1 < t2) {
			const float inv = (float)(1.0 / (t1 - t2));
			side = 1;
			ffar = (float)((t1 + off + kClipEps) * (double)inv);
			fnear = (float)((t1 - off + kClipEps) * (double)inv);
		} else if (t1 > t2) {
			const float inv = (float)(1.0 / (t1 - t2));
			side = 0;
			ffar = (float)((t1 - off - kClipEps) * (double)inv);
			fnear = (float)((t1 + off + kClipEps) * (double)inv);
		} else {
			side = 0;
			fnear = 1.0f;
			ffar = 0.0f;
		}
		if (fnear < 0) fnear = 0; else if (fnear > 1) fnear = 1;
		if (ffar < 0) ffar = 0; else if (ffar > 1) ffar = 1;
		Cnt<COUNT>::add(w.cSplits);

		Seg far;
		far.num = side ? nd.x : nd.y;
		far.f1 = cur.f1 + (cur.f2 - cur.f1) * ffar;
		far.f2 = cur.f2;
		far.ax = cur.ax + ffar * (cur.bx - cur.ax);
		far.ay = cur.ay + ffar * (cur.by - cur.ay);
		far.az = cur.az + ffar * (cur.bz - cur.az);
		far.bx = cur.bx; far.by = cur.by; far.bz = cur.bz;
		const float midf = cur.f1 + (cur.f2 - cur.f1) * fnear;
		const float mx = cur.ax + fnear * (cur.bx - cur.ax);
		const float my = cur.ay + fnear * (cur.by - cur.ay);
		const float mz = cur.az + fnear * (cur.bz - cur.az);
		if (sp < kMaxDepth) stack[sp++] = far; else atomicOr(status, 1);
		cur.num = side ? nd.y : nd.x;
		cur.f2 = midf;
		cur.bx = mx; cur.by = my; cur.bz = mz;
	}
}

// ---- CM_Trace (cm_trace.c:545-687) and CM_TransformedBoxTrace (cm_trace.c:708-805) ---------------
template <bool COUNT>
__global__ void __launch_bounds__(128) traceKernel(const DevMap m, const TraceBatch q) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= q.n) return;

	Work w;
	// raw inputs
	const float s0x = q.starts[3 * (size_t)i], s0y = q.starts[3 * (size_t)i + 1], s0z = q.starts[3 * (size_t)i + 2];
	const float e0x = q.ends[3 * (size_t)i], e0y = q.ends[3 * (size_t)i + 1], e0z = q.ends[3 * (size_t)i + 2];
	float mnx = 0, mny = 0, mnz = 0, mxx = 0, mxy = 0, mxz = 0;
	if (q.mins) { const float *p = q.mins + 3 * (size_t)i * q.hullStride; mnx = p[0]; mny = p[1]; mnz = p[2]; }
	if (q.maxs) { const float *p = q.maxs + 3 * (size_t)i * q.hullStride; mxx = p[0]; mxy = p[1]; mxz = p[2]; }
	const int model = q.models ? q.models[i] : 0;
	w.mask = q.masks[(size_t)i * q.maskStride];
	w.isBox = (model == kBoxModelHandle);
	w.bmnx = w.bmny = w.bmnz = w.bmxx = w.bmxy = w.bmxz = 0.0f;
	if (w.isBox && q.boxmins) {
		w.bmnx = q.boxmins[3 * (size_t)i]; w.bmny = q.boxmins[3 * (size_t)i + 1]; w.bmnz = q.boxmins[3 * (size_t)i + 2];
		w.bmxx = q.boxmaxs[3 * (size_t)i]; w.bmxy = q.boxmaxs[3 * (size_t)i + 1]; w.bmxz = q.boxmaxs[3 * (size_t)i + 2];
	}

	// symmetrise the box (cm_trace.c:582-588 / 756-762): (mins+maxs)*0.5 is a float add and an exact halving
	const float ox = (mnx + mxx) * 0.5f, oy = (mny + mxy) * 0.5f, oz = (mnz + mxz) * 0.5f;
	w.n0x = mnx - ox; w.n0y = mny - oy; w.n0z = mnz - oz;
	w.p0x = mxx - ox; w.p0y = mxy - oy; w.p0z = mxz - oz;
	float lsx = s0x + ox, lsy = s0y + oy, lsz = s0z + oz;
	float lex = e0x + ox, ley = e0y + oy, lez = e0z + oz;

	// CM_TransformedBoxTrace: into the model's frame (cm_trace.c:764-779)
	bool rotated = false;
	float r00 = 0, r01 = 0, r02 = 0, r10 = 0, r11 = 0, r12 = 0, r20 = 0, r21 = 0, r22 = 0;
	if (q.origins) {
		const float gx = q.origins[3 * (size_t)i], gy = q.origins[3 * (size_t)i + 1], gz = q.origins[3 * (size_t)i + 2];
		lsx = lsx - gx; lsy = lsy - gy; lsz = lsz - gz;
		lex = lex - gx; ley = ley - gy; lez = lez - gz;
		const int ri = q.rotIndex ? q.rotIndex[i] : -1;
		if (ri >= 0) {
			rotated = true;
			const float *r = q.rotations + 9 * (size_t)ri;
			r00 = r[0]; r01 = r[1]; r02 = r[2]; r10 = r[3]; r11 = r[4]; r12 = r[5]; r20 = r[6]; r21 = r[7]; r22 = r[8];
			float tx = lsx, ty = lsy, tz = lsz;
			lsx = dotf(r00, r01, r02, tx, ty, tz); lsy = dotf(r10, r11, r12, tx, ty, tz); lsz = dotf(r20, r21, r22, tx, ty, tz);
			tx = lex; ty = ley; tz = lez;
			lex = dotf(r00, r01, r02, tx, ty, tz); ley = dotf(r10, r11, r12, tx, ty, tz); lez = dotf(r20, r21, r22, tx, ty, tz);
		}
	}

	// CM_Trace proper.  With origins the reference symmetrises a second time inside CM_Trace
	// (cm_trace.c:582-588); the box is already symmetric, but the float ops are repeated verbatim.
	float ts0x = lsx, ts0y = lsy, ts0z = lsz, te0x = lex, te0y = ley, te0z = lez;   // CM_Trace's start/end arguments
	if (q.origins) {
		const float o2x = (w.n0x + w.p0x) * 0.5f, o2y = (w.n0y + w.p0y) * 0.5f, o2z = (w.n0z + w.p0z) * 0.5f;
		const float a0x = w.n0x - o2x, a0y = w.n0y - o2y, a0z = w.n0z - o2z;
		const float a1x = w.p0x - o2x, a1y = w.p0y - o2y, a1z = w.p0z - o2z;
		w.n0x = a0x; w.n0y = a0y; w.n0z = a0z; w.p0x = a1x; w.p0y = a1y; w.p0z = a1z;
		lsx = lsx + o2x; lsy = lsy + o2y; lsz = lsz + o2z;
		lex = lex + o2x; ley = ley + o2y; lez = lez + o2z;
	} else {
		ts0x = s0x; ts0y = s0y; ts0z = s0z; te0x = e0x; te0y = e0y; te0z = e0z;
	}
	w.sx = lsx; w.sy = lsy; w.sz = lsz; w.ex = lex; w.ey = ley; w.ez = lez;

	w.allsolid = 0; w.startsolid = 0; w.fraction = 1.0f;
	w.pnx = w.pny = w.pnz = w.pd = 0.0f; w.ptypesign = 0;
	w.surfaceFlags = 0; w.contents = 0; w.hitId = -1;
	w.isPoint = false;
	w.extx = w.exty = w.extz = 0.0f;
	w.cNodes = w.cLeafs = w.cBrushRefs = w.cBrushBounds = w.cBrushSides = w.cCross = w.cSplits = 0;
	w.cPatches = w.cPatchPlanes = w.cFacets = 0;

	// swept bounds (cm_trace.c:628-636)
	if (w.sx < w.ex) { w.lox = w.sx + w.n0x; w.hix = w.ex + w.p0x; } else { w.lox = w.ex + w.n0x; w.hix = w.sx + w.p0x; }
	if (w.sy < w.ey) { w.loy = w.sy + w.n0y; w.hiy = w.ey + w.p0y; } else { w.loy = w.ey + w.n0y; w.hiy = w.sy + w.p0y; }
	if (w.sz < w.ez) { w.loz = w.sz + w.n0z; w.hiz = w.ez + w.p0z; } else { w.loz = w.ez + w.n0z; w.hiz = w.sz + w.p0z; }

	int4 mleaf = make_int4(0, 0, 0, 0);
	if (model) {
		mleaf = w.isBox ? make_int4(m.boxLeafBrush, 1, 0, 0) : __ldg(&m.modelLeafs[model]);
	}

	if (ts0x == te0x && ts0y == te0y && ts0z == te0z) {
		// position test: isPoint stays false and extents zero here (cm_trace.c:641-646)
		if (model) testLeaf<COUNT>(m, w, mleaf);
		else       positionTestWorld<COUNT>(m, w, q.status);
	} else {
		if (w.n0x == 0.0f && w.n0y == 0.0f && w.n0z == 0.0f) {
			w.isPoint = true;
		} else {
			w.extx = w.p0x; w.exty = w.p0y; w.extz = w.p0z;
		}
		if (model) sweepLeaf<COUNT>(m, w, mleaf);
		else       sweepWorld<COUNT>(m, w, q.status);
	}

	// plane back to world space (cm_trace.c:789-793): rows of the transpose are columns of the matrix
	if (rotated && w.fraction != 1.0f) {
		const float tx = w.pnx, ty = w.pny, tz = w.pnz;
		w.pnx = dotf(r00, r10, r20, tx, ty, tz);
		w.pny = dotf(r01, r11, r21, tx, ty, tz);
		w.pnz = dotf(r02, r12, r22, tx, ty, tz);
	}

	// endpos from the caller's start/end (cm_trace.c:672-678, 797-799)
	float px, py, pz;
	if (!q.origins && w.fraction == 1.0f) {
		px = e0x; py = e0y; pz = e0z;
	} else {
		px = s0x + w.fraction * (e0x - s0x);
		py = s0y + w.fraction * (e0y - s0y);
		pz = s0z + w.fraction * (e0z - s0z);
	}

	uint2 *o = q.out + 7 * (size_t)i;
	o[0] = make_uint2((unsigned)w.allsolid, (unsigned)w.startsolid);
	o[1] = make_uint2(__float_as_uint(w.fraction), __float_as_uint(px));
	o[2] = make_uint2(__float_as_uint(py), __float_as_uint(pz));
	o[3] = make_uint2(__float_as_uint(w.pnx), __float_as_uint(w.pny));
	o[4] = make_uint2(__float_as_uint(w.pnz), __float_as_uint(w.pd));
	o[5] = make_uint2((unsigned)w.ptypesign & 0xffffu, (unsigned)w.surfaceFlags);
	o[6] = make_uint2((unsigned)w.contents, 0u);
	if (q.hitIds) q.hitIds[i] = w.hitId;

	if (COUNT) {
		Counters *c = q.counters;
		atomicAdd(&c->traces, 1ull);
		atomicAdd(&c->nodes, (unsigned long long)w.cNodes);
		atomicAdd(&c->leafs, (unsigned long long)w.cLeafs);
		atomicAdd(&c->brushRefs, (unsigned long long)w.cBrushRefs);
		atomicAdd(&c->brushBounds, (unsigned long long)w.cBrushBounds);
		atomicAdd(&c->brushSides, (unsigned long long)w.cBrushSides);
		atomicAdd(&c->crossings, (unsigned long long)w.cCross);
		atomicAdd(&c->splits, (unsigned long long)w.cSplits);
		atomicAdd(&c->patches, (unsigned long long)w.cPatches);
		atomicAdd(&c->patchPlanes, (unsigned long long)w.cPatchPlanes);
		atomicAdd(&c->facets, (unsigned long long)w.cFacets);
	}
}

// ---- CM_PointContents / CM_TransformedPointContents, cm_test.c:217-294 -----------------------------
__global__ void __launch_bounds__(128) pointContentsKernel(const DevMap m, const PointBatch q) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= q.n) return;
	float px = q.points[3 * (size_t)i], py = q.points[3 * (size_t)i + 1], pz = q.points[3 * (size_t)i + 2];
	const int model = q.models ? q.models[i] : 0;
	if (q.origins) {
		px = px - q.origins[3 * (size_t)i]; py = py - q.origins[3 * (size_t)i + 1]; pz = pz - q.origins[3 * (size_t)i + 2];
		const int ri = q.rotIndex ? q.rotIndex[i] : -1;
		if (ri >= 0) {
			// p_l = (p.forward, -(p.right), p.up); rows of the matrix are forward, -right, up, and
			// -(a.b) == a.(-b) exactly, so the matrix rows can be used directly (cm_test.c:287-290)
			const float *r = q.rotations + 9 * (size_t)ri;
			const float tx = px, ty = py, tz = pz;
			px = dotf(tx, ty, tz, r[0], r[1], r[2]);
			py = dotf(tx, ty, tz, r[3], r[4], r[5]);
			pz = dotf(tx, ty, tz, r[6], r[7], r[8]);
		}
	}
	const bool isBox = (model == kBoxModelHandle);
	float bmnx = 0, bmny = 0, bmnz = 0, bmxx = 0, bmxy = 0, bmxz = 0;
	if (isBox && q.boxmins) {
		bmnx = q.boxmins[3 * (size_t)i]; bmny = q.boxmins[3 * (size_t)i + 1]; bmnz = q.boxmins[3 * (size_t)i + 2];
		bmxx = q.boxmaxs[3 * (size_t)i]; bmxy = q.boxmaxs[3 * (size_t)i + 1]; bmxz = q.boxmaxs[3 * (size_t)i + 2];
	}
	int4 leaf;
	if (model) {
		leaf = isBox ? make_int4(m.boxLeafBrush, 1, 0, 0) : __ldg(&m.modelLeafs[model]);
	} else {
		// CM_PointLeafnum_r, cm_test.c:31-54
		int num = 0;
		while (num >= 0) {
			const int4 nd = __ldg(&m.nodes[num]);
			float d;
			if (nd.z < 3) {
				d = sel3(px, py, pz, nd.z) - __int_as_float(nd.w);
			} else {
				const float4 pl = __ldg(&m.nodePlanes[num]);
				d = dotf(pl.x, pl.y, pl.z, px, py, pz) - pl.w;
			}
			num = (d < 0) ? nd.y : nd.x;
		}
		leaf = __ldg(&m.leafs[-1 - num]);
	}
	int contents = 0;
	for (int k = 0; k < leaf.y; k++) {
		const int b = __ldg(&m.leafbrushes[leaf.x + k]);
		float4 bmn = __ldg(&m.brushBox[2 * b]);
		float4 bmx = __ldg(&m.brushBox[2 * b + 1]);
		const bool boxBrush = isBox && b == m.boxBrush;
		if (boxBrush) {
			bmn.x = bmnx; bmn.y = bmny; bmn.z = bmnz;
			bmx.x = bmxx; bmx.y = bmxy; bmx.z = bmxz;
		}
		// CM_BoundsIntersectPoint, cm_test.c:502-515
		if (bmx.x < px - kBoundsEps || bmx.y < py - kBoundsEps || bmx.z < pz - kBoundsEps ||
		    bmn.x > px + kBoundsEps || bmn.y > py + kBoundsEps || bmn.z > pz + kBoundsEps) continue;
		const int firstSide = __float_as_int(bmx.w);
		const int nsides = __ldg(&m.brushNumSides[b]);
		int s;
		for (s = 0; s < nsides; s++) {
			float4 pl = __ldg(&m.sidePlanes[firstSide + s]);
			if (boxBrush) {
				const int axis = s >> 1;
				pl.w = (s & 1) ? -sel3(bmnx, bmny, bmnz, axis) : sel3(bmxx, bmxy, bmxz, axis);
			}
			if (dotf(px, py, pz, pl.x, pl.y, pl.z) > pl.w) break;
		}
		if (s == nsides) contents |= __float_as_int(bmn.w);
	}
	q.out[i] = contents;
}

}  // namespace cmgpu
