// cm_trace_fast.cuh -- the hot path: CM_BoxTrace against the WORLD with a hull and a mask shared by
// the whole batch (BASELINE configs 1, 2 and 5), on maps without patches (or with cm_noCurves set).
// Everything else (inline models, temp boxes, rotations, per-item hulls, patches) runs traceKernel
// in cm_kernels.cuh.
//
// This header holds the data layout (FastMap, FastHull) and the STEPS of the hot path -- beginFast, slabReject,
// clipBrushFast, finishFast -- plus traceFastKernel, the one-item-per-lane persistent scheduler over them.  The other
// two schedulers over the same steps are tracePoolKernel (cm_trace_pool.cuh: K items per lane in shared memory, the
// kernel of big launches) and traceSimpleKernel (cm_trace_simple.cuh: one thread per item, small launches).
//
// Same persistent flat-state-machine scheduling as traceKernel (see there), re-cut so that every
// step is a handful of instructions:
//   * leafs are STREAMS of 32-byte brush entries terminated by a sentinel; a node's child refers
//     straight to its leaf's stream, so there is no leaf record to load (cm_trace.c:370-393);
//   * the entries carry the brush bounds already widened by BOUNDS_CLIP_EPSILON (the float ops
//     of cm_test.c:481-494 done once at upload), so CM_BoundsIntersect is six compares;
//   * an axial node compares the FLOAT differences t1,t2 with float images of offset+1 and
//     -offset-1 computed once per batch (exact, see makeFastHull in cm_gpu.cu);
//   * the running result is (fraction, reference to the clip side, contents); the plane and surface
//     flags are fetched when the record is written;
//   * the brush clip first runs with float quotients that are only used to REJECT (enterFrac
//     certainly >= leaveFrac, or certainly >= trace.fraction); the exact double divisions of
//     cm_trace.c:314,324 are done only for brushes that survive that;
//   * a brush that has just been clipped is not clipped again in the next leaf (the reference's
//     checkcount, cm_trace.c:381-384; a repeat is a no-op anyway).
#pragma once

#include "cm_kernels.cuh"

#ifndef CMGPU_SLAB_F32X2
#define CMGPU_SLAB_F32X2 0
#endif
#ifndef CMGPU_SLAB_UNCOND
#define CMGPU_SLAB_UNCOND 0
#endif
namespace cmgpu {

struct FastMap {
	const int4   *nodes;         // child0, child1, type (0..2 axial, 3 general), dist bits; child c < 0: leaf stream at entry -1-c
	const float4 *nodePlanes;    // normal.xyz, dist (type 3 only)
	const float4 *stream;        // [2k] mins-eps.xyz + contents ; [2k+1] maxs+eps.xyz + brush index, -1 ends the leaf
	const float4 *brushRec;      // [2b] mins.xyz + firstSide ; [2b+1] maxs.xyz + (numSides | canonical << 30)
	const int    *brushContents;
	const float4 *sidePlanes;
	const int2   *sideMeta;
};

// what depends only on the batch's hull and mask; by value in the kernel parameters (constant bank)
struct FastHull {
	float n0[3], p0[3];          // tw.size[0], tw.size[1]  (cm_trace.c:582-588)
	float off[3];                // offset added to start and end
	float nodeHi[3], nodeLo[3];  // t >= offset+1  <=>  t >= nodeHi ;  t < -offset-1  <=>  t < nodeLo   (t float)
	int   isPoint;
	int   mask;
	int   slabOK;                // every brush coordinate of the map is below 2^15 in magnitude
};

enum FastMode : int { FM_IDLE = 0, FM_NODE, FM_LEAF, FM_POP, FM_SPLIT, FM_BRUSH, FM_DONE, FM_PNODE, FM_PLEAF };

struct __align__(16) FastSeg { int num; float f1, f2, ax; float ay, az, bx, by; float bz, pad0, pad1, pad2; };

struct FastLane {
	int   slot;                          // position in the (binned) batch
	int   item;                          // its number in the caller's arrays (== slot when the batch is not binned)
	float sx, sy, sz, ex, ey, ez;        // tw.start / tw.end
	float lox, loy, loz, hix, hiy, hiz;  // tw.bounds
	float fraction;
	int   flags;                         // bit0 allsolid, bit1 startsolid, bit2 slab pre-test allowed for this item
	int   planeRef;                      // brush side whose plane / surfaceFlags the result carries, -1 none
	int   hitId;                         // brush that set the result (its contents are the result's contents), -1 none
	int   num, sp;                       // tree cursor
	float f1, f2, ax, ay, az, bx, by, bz;
	int   k;                             // leaf stream cursor
	int   pend;                          // brush waiting to be clipped; position test: leafs seen so far
	int   last0, last1;                  // brushes clipped most recently
	unsigned cNodes, cLeafs, cBrushRefs, cBrushBounds, cBrushSides, cCross, cSplits;
};

// The traversal stacks are a few tens of MB that every item pushes to and pops from, next to GBs of rays and
// records that stream through L2 once: ask L2 to keep the former (evict_last) so that pops do not come from DRAM.
__device__ __forceinline__ unsigned long long l2KeepPolicy() {
	unsigned long long p;
	asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
	return p;
}
__device__ __forceinline__ void stKeep(float4 *p, const float4 v, const unsigned long long pol) {
	asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;"
	             :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(pol) : "memory");
}
__device__ __forceinline__ float4 ldKeep(const float4 *p, const unsigned long long pol) {
	float4 v;
	asm volatile("ld.global.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
	             : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol) : "memory");
	return v;
}

__device__ __forceinline__ float ldKeepF(const float *p, const unsigned long long pol) {
	float v;
	asm volatile("ld.global.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol) : "memory");
	return v;
}

__device__ __forceinline__ void stKeep32(int *p, const int v, const unsigned long long pol) {
	asm volatile("st.global.L2::cache_hint.b32 [%0], %1, %2;" :: "l"(p), "r"(v), "l"(pol) : "memory");
}

__device__ __forceinline__ float rcpApprox(float x) {
	float r;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
}

// one side of pass 1: exact distances and flags, approximate quotient.  The entering side with the largest
// quotient is remembered together with the runner-up's quotient, so that a clear winner needs ONE exact division.
struct ClipApprox {
	bool so, go, miss, anyEnter;
	float qe, q2, ql;          // largest / second largest entering quotient (clamped at 0), smallest leaving quotient (clamped at 1)
	double bd1, bd2;           // distances of the side that owns qe
	int bs;                    // its index
};

__device__ __forceinline__ void approxSide(ClipApprox &c, const double d1, const double d2, const int s) {
	c.go |= d2 > 0;
	c.so |= d1 > 0;
	c.miss |= d1 > 0 && (d2 >= kClipEps || d2 >= d1);
	const bool cross = !(d1 <= 0 && d2 <= 0);
	const bool entering = d1 > d2;
	const float numf = (float)(entering ? d1 - kClipEps : d1 + kClipEps);
	const float denf = (float)(d1 - d2);
	const float q = numf * rcpApprox(denf);
	if (cross && entering) {
		const float qc = fmaxf(q, 0.0f);
		c.anyEnter = true;
		if (qc > c.qe) { c.q2 = c.qe; c.qe = qc; c.bd1 = d1; c.bd2 = d2; c.bs = s; }
		else c.q2 = fmaxf(c.q2, qc);
	}
	if (cross && !entering) c.ql = fminf(c.ql, fminf(q, 1.0f));
}

// Conservative float slab test of the sweep against a leaf entry (the brush box widened by BOUNDS_CLIP_EPSILON),
// used to skip brushes that certainly leave the trace untouched: the start is certainly outside (so neither
// startsolid nor allsolid) and enterFrac is certainly >= leaveFrac, or certainly >= trace.fraction
// (cm_trace.c:352-353).  Why it is safe: the entry box widened by the hull contains the planes the reference
// clips against (faces pushed out by the hull, then moved by SURFACE_CLIP_EPSILON = 1/8) with 1/8 to spare on every
// side, and for |coordinates| < 2^15 all float error below (two roundings of a coordinate, rcp.approx, one
// product) is < 0.05 units.  So the reference's enter fraction on the axis that gives tEnter is > tEnter and its
// leave fraction on the axis that gives tLeave is < tLeave; tEnter > 0 means that side is really entered from
// outside, tLeave < 1 means the end point really lies beyond that face.  NaNs (0 * inf on a parallel axis) drop
// out of fminf/fmaxf and make the comparisons false: no rejection.
__device__ __forceinline__ bool slabReject(const FastHull &h, const FastLane &L, const float4 e0, const float4 e1) {
	const float ix = rcpApprox(L.ex - L.sx), iy = rcpApprox(L.ey - L.sy), iz = rcpApprox(L.ez - L.sz);
	const float x0 = ((e0.x - h.p0[0]) - L.sx) * ix, x1 = ((e1.x + h.p0[0]) - L.sx) * ix;
	const float y0 = ((e0.y - h.p0[1]) - L.sy) * iy, y1 = ((e1.y + h.p0[1]) - L.sy) * iy;
	const float z0 = ((e0.z - h.p0[2]) - L.sz) * iz, z1 = ((e1.z + h.p0[2]) - L.sz) * iz;
	const float tEnter = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fminf(z0, z1));
	const float tLeave = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1));
	return tEnter > 0.0f && ((tEnter > tLeave && tLeave < 1.0f) || tEnter > L.fraction);
}

// distances of the start and end point to the two axial sides of one axis of a canonical brush:
// side 2a   : normal -e_a, dist -mins[a], corner component size[1][a]  ->  d = -p - (size[1][a] - mins[a])
// side 2a+1 : normal +e_a, dist  maxs[a], corner component size[0][a]  ->  d =  p - (maxs[a] - size[0][a])
// (DotProductDP with a unit normal is exactly +-component; cm_trace.c:295-297)
struct AxisDist { double a0, b0, a1, b1; };
__device__ __forceinline__ AxisDist axisDist(const float s, const float e, const float mn, const float mx, const float n0, const float p0) {
	const double dn = (double)p0 - (double)mn, dp = (double)mx - (double)n0;
	AxisDist d;
	d.a0 = -(double)s - dn; d.b0 = -(double)e - dn;
	d.a1 = (double)s - dp;  d.b1 = (double)e - dp;
	return d;
}

// slabReject with everything that depends only on the item computed once per leaf step: the three reciprocals (MUFU)
// and the start point pushed out by the hull, so that an entry costs six subtractions, six products and the min/max
// tree.  (e0 - (s + p0)) instead of ((e0 - p0) - s): one rounding of a coordinate fewer, same bound on the error.
struct SlabItem { float ix, iy, iz, lx, ly, lz, hx, hy, hz, fraction; };
__device__ __forceinline__ SlabItem slabItem(const FastHull &h, const float sx, const float sy, const float sz,
                                             const float ex, const float ey, const float ez, const float fraction) {
	SlabItem t;
	t.ix = rcpApprox(ex - sx); t.iy = rcpApprox(ey - sy); t.iz = rcpApprox(ez - sz);
	t.lx = sx + h.p0[0]; t.ly = sy + h.p0[1]; t.lz = sz + h.p0[2];
	t.hx = sx - h.p0[0]; t.hy = sy - h.p0[1]; t.hz = sz - h.p0[2];
	t.fraction = fraction;
	return t;
}
__device__ __forceinline__ bool slabRejectPre(const SlabItem &t, const float4 e0, const float4 e1) {
	const float x0 = (e0.x - t.lx) * t.ix, x1 = (e1.x - t.hx) * t.ix;
	const float y0 = (e0.y - t.ly) * t.iy, y1 = (e1.y - t.hy) * t.iy;
	const float z0 = (e0.z - t.lz) * t.iz, z1 = (e1.z - t.hz) * t.iz;
	const float tEnter = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fminf(z0, z1));
	const float tLeave = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1));
	return tEnter > 0.0f && ((tEnter > tLeave && tLeave < 1.0f) || tEnter > t.fraction);
}

#if CMGPU_SLAB_F32X2
// the same test with the six subtractions and six products as three packed pairs (sub.f32x2 / mul.f32x2, sm_100: each
// half rounds like the scalar operation)
__device__ __forceinline__ unsigned long long pack2(const float a, const float b) {
	unsigned long long r;
	asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
	return r;
}
__device__ __forceinline__ void mulsub2(const float a0, const float a1, const float b0, const float b1, const float m, float &r0, float &r1) {
	unsigned long long d;
	asm("sub.f32x2 %0, %1, %2;" : "=l"(d) : "l"(pack2(a0, a1)), "l"(pack2(b0, b1)));
	asm("mul.f32x2 %0, %1, %2;" : "=l"(d) : "l"(d), "l"(pack2(m, m)));
	asm("mov.b64 {%0, %1}, %2;" : "=f"(r0), "=f"(r1) : "l"(d));
}
__device__ __forceinline__ bool slabRejectPre2(const SlabItem &t, const float4 e0, const float4 e1) {
	float x0, x1, y0, y1, z0, z1;
	mulsub2(e0.x, e1.x, t.lx, t.hx, t.ix, x0, x1);
	mulsub2(e0.y, e1.y, t.ly, t.hy, t.iy, y0, y1);
	mulsub2(e0.z, e1.z, t.lz, t.hz, t.iz, z0, z1);
	const float tEnter = fmaxf(fmaxf(fminf(x0, x1), fminf(y0, y1)), fminf(z0, z1));
	const float tLeave = fminf(fminf(fmaxf(x0, x1), fmaxf(y0, y1)), fmaxf(z0, z1));
	return tEnter > 0.0f && ((tEnter > tLeave && tLeave < 1.0f) || tEnter > t.fraction);
}
#endif

struct Sweep { float sx, sy, sz, ex, ey, ez; };
struct EnterLeave { float enter, leave; int lead; };

__device__ __forceinline__ void sideDist(const FastHull &h, const Sweep &L, const float4 pl, double &d1, double &d2) {
	// tw.offsets[signbits]: the box corner that touches the plane first
	const float cx = pl.x < 0.0f ? h.p0[0] : h.n0[0];
	const float cy = pl.y < 0.0f ? h.p0[1] : h.n0[1];
	const float cz = pl.z < 0.0f ? h.p0[2] : h.n0[2];
	const double dist = (double)pl.w - dotd((double)cx, (double)cy, (double)cz, pl.x, pl.y, pl.z);
	d1 = dotd((double)L.sx, (double)L.sy, (double)L.sz, pl.x, pl.y, pl.z) - dist;
	d2 = dotd((double)L.ex, (double)L.ey, (double)L.ez, pl.x, pl.y, pl.z) - dist;
}

// The reference's own arithmetic for every side of a brush that is known not to be missed and to be entered from
// outside (cm_trace.c:291-332): enterFrac, leaveFrac and the lead side.  Rare (ties, grazing sweeps), kept out of line.
__device__ __forceinline__ EnterLeave exactEnterLeave(const FastMap &fm, const FastHull &h, const Sweep L, const float4 r0, const float4 r1) {
	const int firstSide = __float_as_int(r0.w);
	const int ns = __float_as_int(r1.w) & 0xffffff;
	const bool canonical = (__float_as_int(r1.w) >> 30) & 1;
	float enter = -1.0f, leave = 1.0f;
	int lead = -1;
	auto cross = [&](const double d1, const double d2, const int s) {
		const bool crossing = !(d1 <= 0 && d2 <= 0);
		const bool entering = d1 > d2;
		const double num = entering ? d1 - kClipEps : d1 + kClipEps;
		const double den = crossing ? d1 - d2 : 1.0;
		const float q = (float)(num / den);
		const float fe = q < 0 ? 0.0f : q;
		const float fl = q > 1 ? 1.0f : q;
		if (crossing && entering && fe > enter) { enter = fe; lead = s; }
		if (crossing && !entering && fl < leave) leave = fl;
	};
	int side = 0;
	if (canonical) {
		const AxisDist x = axisDist(L.sx, L.ex, r0.x, r1.x, h.n0[0], h.p0[0]);
		cross(x.a0, x.b0, 0); cross(x.a1, x.b1, 1);
		const AxisDist y = axisDist(L.sy, L.ey, r0.y, r1.y, h.n0[1], h.p0[1]);
		cross(y.a0, y.b0, 2); cross(y.a1, y.b1, 3);
		const AxisDist z = axisDist(L.sz, L.ez, r0.z, r1.z, h.n0[2], h.p0[2]);
		cross(z.a0, z.b0, 4); cross(z.a1, z.b1, 5);
		side = 6;
	}
	for (; side < ns; side++) {
		double d1, d2;
		sideDist(h, L, __ldg(&fm.sidePlanes[firstSide + side]), d1, d2);
		cross(d1, d2, side);
	}
	EnterLeave r;
	r.enter = enter; r.leave = leave; r.lead = lead;
	return r;
}

// CM_TraceThroughBrush (cm_trace.c:261-363) for brush b.  Returns true when the trace is over (fraction 0).
template <bool COUNT>
__device__ __forceinline__ bool clipBrushFast(const FastMap &fm, const FastHull &h, FastLane &L, const int b) {
	const Pair4 rr = ldg32(&fm.brushRec[2 * b]);
	const float4 r0 = rr.a, r1 = rr.b;
	const int firstSide = __float_as_int(r0.w);
	const int ns = __float_as_int(r1.w) & 0xffffff;
	const bool canonical = (__float_as_int(r1.w) >> 30) & 1;
	if (ns == 0) return false;
	// ---- pass 1: flags (exact) and quotients (approximate: they only reject, or name a clear winner) ----
	ClipApprox c;
	c.so = c.go = c.miss = c.anyEnter = false;
	c.qe = -1.0f; c.q2 = -1.0f; c.ql = 1.0f;
	c.bd1 = c.bd2 = 0.0; c.bs = -1;
	int side = 0;
	if (canonical) {
		const AxisDist x = axisDist(L.sx, L.ex, r0.x, r1.x, h.n0[0], h.p0[0]);
		approxSide(c, x.a0, x.b0, 0); approxSide(c, x.a1, x.b1, 1);
		const AxisDist y = axisDist(L.sy, L.ey, r0.y, r1.y, h.n0[1], h.p0[1]);
		approxSide(c, y.a0, y.b0, 2); approxSide(c, y.a1, y.b1, 3);
		const AxisDist z = axisDist(L.sz, L.ez, r0.z, r1.z, h.n0[2], h.p0[2]);
		approxSide(c, z.a0, z.b0, 4); approxSide(c, z.a1, z.b1, 5);
		side = 6;
		if (COUNT) L.cBrushSides += 6;
	}
	Sweep sw;
	sw.sx = L.sx; sw.sy = L.sy; sw.sz = L.sz; sw.ex = L.ex; sw.ey = L.ey; sw.ez = L.ez;
	for (; side < ns && !c.miss; side++) {
		double d1, d2;
		sideDist(h, sw, __ldg(&fm.sidePlanes[firstSide + side]), d1, d2);
		if (COUNT) L.cBrushSides++;
		approxSide(c, d1, d2, side);
	}
	if (c.miss) return false;                       // wholly in front of one face
	if (!c.so) {                                    // original point was inside the brush
		L.flags |= 2;
		if (!c.go) {
			L.flags |= 1;
			L.fraction = 0.0f;
			L.hitId = b;
			return true;
		}
		return false;
	}
	if (!c.anyEnter) return false;                  // enterFrac stays -1
	float enter = -1.0f, leave = 1.0f;
	int lead = -1;
	bool decided = false;
	if (c.qe >= 0.0f) {
		// |q - f| <= 2^-21 |f| for every quotient (two conversions, rcp.approx, one product, the two roundings of
		// the reference's own f); 2^-19 leaves a factor 4.  NaN/inf never satisfy a comparison below.
		const float kRel = 1.9073486e-6f, kAbs = 1e-30f;
		const float eLo = c.qe - (c.qe * kRel + kAbs), eHi = c.qe + (c.qe * kRel + kAbs);
		const float lHi = c.ql + (fabsf(c.ql) * kRel + kAbs), lLo = c.ql - (fabsf(c.ql) * kRel + kAbs);
		if (eLo > lHi) return false;                // enterFrac >= leaveFrac
		if (eLo > L.fraction) return false;         // enterFrac >= trace.fraction
		// a clear winner among the entering sides that clearly precedes every leaving side: one exact division
		if (c.q2 < eLo - (c.qe * kRel + kAbs) && eHi < lLo) {
			float f = (float)((c.bd1 - kClipEps) / (c.bd1 - c.bd2));
			if (f < 0) f = 0.0f;
			enter = f;
			lead = c.bs;
			decided = true;
			if (COUNT) L.cCross++;
		}
	}
	if (!decided) {
		const EnterLeave r = exactEnterLeave(fm, h, sw, r0, r1);
		enter = r.enter; leave = r.leave; lead = r.lead;
	}
	if (enter < leave && enter > -1 && enter < L.fraction) {
		if (enter < 0) enter = 0;
		L.fraction = enter;
		if (lead >= 0) L.planeRef = firstSide + lead;
		L.hitId = b;
	}
	return L.fraction == 0.0f;
}

// ---- item setup: CM_Trace, cm_trace.c:545-669 ------------------------------------------------
template <bool COUNT>
__device__ __forceinline__ int beginFast(const TraceBatch &q, const FastHull &h, FastLane &L, const int slot) {
	float s0x, s0y, s0z, e0x, e0y, e0z;
	if (q.recs) {
		const Pair4 ab = ldcs32(&q.recs[2 * (size_t)slot]);
		const float4 a = ab.a, b = ab.b;
		s0x = a.x; s0y = a.y; s0z = a.z; e0x = b.x; e0y = b.y; e0z = b.z;
		L.item = __float_as_int(a.w);
	} else {
		const size_t o = 3 * (size_t)slot;
		s0x = q.starts[o]; s0y = q.starts[o + 1]; s0z = q.starts[o + 2];
		e0x = q.ends[o]; e0y = q.ends[o + 1]; e0z = q.ends[o + 2];
		L.item = slot;
	}
	L.slot = slot;
	L.sx = s0x + h.off[0]; L.sy = s0y + h.off[1]; L.sz = s0z + h.off[2];
	L.ex = e0x + h.off[0]; L.ey = e0y + h.off[1]; L.ez = e0z + h.off[2];
	L.flags = 0; L.fraction = 1.0f; L.planeRef = -1; L.hitId = -1;
	L.sp = 0; L.last0 = L.last1 = -1; L.pend = 0; L.k = 0;
	if (COUNT) L.cNodes = L.cLeafs = L.cBrushRefs = L.cBrushBounds = L.cBrushSides = L.cCross = L.cSplits = 0;
	// swept bounds (cm_trace.c:628-636)
	if (L.sx < L.ex) { L.lox = L.sx + h.n0[0]; L.hix = L.ex + h.p0[0]; } else { L.lox = L.ex + h.n0[0]; L.hix = L.sx + h.p0[0]; }
	if (L.sy < L.ey) { L.loy = L.sy + h.n0[1]; L.hiy = L.ey + h.p0[1]; } else { L.loy = L.ey + h.n0[1]; L.hiy = L.sy + h.p0[1]; }
	if (L.sz < L.ez) { L.loz = L.sz + h.n0[2]; L.hiz = L.ez + h.p0[2]; } else { L.loz = L.ez + h.n0[2]; L.hiz = L.sz + h.p0[2]; }
	// the float slab test is only trusted for coordinates below 2^15 (see slabReject)
	{
		const float big = fmaxf(fmaxf(fmaxf(fabsf(L.lox), fabsf(L.loy)), fmaxf(fabsf(L.loz), fabsf(L.hix))), fmaxf(fabsf(L.hiy), fabsf(L.hiz)));
		if (h.slabOK && big < 32768.0f) L.flags |= 4;
	}
	L.num = 0; L.f1 = 0.0f; L.f2 = 1.0f;
	L.ax = L.sx; L.ay = L.sy; L.az = L.sz; L.bx = L.ex; L.by = L.ey; L.bz = L.ez;
	if (s0x == e0x && s0y == e0y && s0z == e0z) return FM_PNODE;     // position test (cm_trace.c:641-646)
	return FM_NODE;
}

// ---- item epilogue: cm_trace.c:671-686 -------------------------------------------------------
template <bool COUNT>
__device__ __forceinline__ void finishFast(const FastMap &fm, const TraceBatch &q, FastLane &L) {
	int i = L.slot;
	if (q.compact) {
		// Two-phase results: neighbouring slots finish in the same warp at about the same time, so these 16-byte
		// stores fill their sectors while they are in L2 -- unlike 56-byte records scattered by item number.
		q.compact[i] = make_uint4(__float_as_uint(L.fraction), (unsigned)(L.flags & 3), (unsigned)L.planeRef, (unsigned)L.hitId);
		return;
	}
	float s0x, s0y, s0z, e0x, e0y, e0z;
	if (q.recs) {
		const Pair4 ab = ldcs32(&q.recs[2 * (size_t)i]);
		const float4 a = ab.a, b = ab.b;
		s0x = a.x; s0y = a.y; s0z = a.z; e0x = b.x; e0y = b.y; e0z = b.z;
		i = __float_as_int(a.w);
	} else {
		const size_t o = 3 * (size_t)i;
		s0x = q.starts[o]; s0y = q.starts[o + 1]; s0z = q.starts[o + 2];
		e0x = q.ends[o]; e0y = q.ends[o + 1]; e0z = q.ends[o + 2];
	}
	float4 pl = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
	int2 meta = make_int2(0, 0);
	if (L.planeRef >= 0) {
		pl = __ldg(&fm.sidePlanes[L.planeRef]);
		meta = __ldg(&fm.sideMeta[L.planeRef]);
	}
	const int contents = L.hitId >= 0 ? __ldg(&fm.brushContents[L.hitId]) : 0;
	float px, py, pz;
	if (L.fraction == 1.0f) {
		px = e0x; py = e0y; pz = e0z;
	} else {
		px = s0x + L.fraction * (e0x - s0x);
		py = s0y + L.fraction * (e0y - s0y);
		pz = s0z + L.fraction * (e0z - s0z);
	}
	// Seven streaming 8-byte stores.  Records are scattered by item number, so most 32-byte sectors leave L2 half
	// written and cost a read-modify-write in HBM (measured: 2.4 GB read + 2.1 GB written per 16 Mi records, against
	// 0.94 GB of payload; wider stores and other cache policies do not change that, only writing in slot order would).
	uint2 *o = q.out + 7 * (size_t)i;
	__stcs(&o[0], make_uint2((unsigned)(L.flags & 1), (unsigned)((L.flags >> 1) & 1)));
	__stcs(&o[1], make_uint2(__float_as_uint(L.fraction), __float_as_uint(px)));
	__stcs(&o[2], make_uint2(__float_as_uint(py), __float_as_uint(pz)));
	__stcs(&o[3], make_uint2(__float_as_uint(pl.x), __float_as_uint(pl.y)));
	__stcs(&o[4], make_uint2(__float_as_uint(pl.z), __float_as_uint(pl.w)));
	__stcs(&o[5], make_uint2((unsigned)meta.y & 0xffffu, (unsigned)meta.x));
	__stcs(&o[6], make_uint2((unsigned)contents, 0u));
	if (q.hitIds) q.hitIds[i] = L.hitId;
	if (COUNT) {
		Counters *c = q.counters;
		atomicAdd(&c->traces, 1ull);
		atomicAdd(&c->nodes, (unsigned long long)L.cNodes);
		atomicAdd(&c->leafs, (unsigned long long)L.cLeafs);
		atomicAdd(&c->brushRefs, (unsigned long long)L.cBrushRefs);
		atomicAdd(&c->brushBounds, (unsigned long long)L.cBrushBounds);
		atomicAdd(&c->brushSides, (unsigned long long)L.cBrushSides);
		atomicAdd(&c->crossings, (unsigned long long)L.cCross);
		atomicAdd(&c->splits, (unsigned long long)L.cSplits);
	}
}

// Second phase of a binned hot launch: one thread per ITEM.  Finds the item's 16-byte result by its slot, reads the
// item's own start and end (input arrays, coalesced), fetches the clip plane / surface flags / contents from the map
// (L2-resident) and builds the trace_t (cm_trace.c:671-686).  The block's records are put together in shared
// memory and leave as fully coalesced 16-byte stores: every sector of the output is written exactly once, whole.
constexpr int kExpandBlock = 256;
__global__ void __launch_bounds__(kExpandBlock) expandRecordsKernel(const FastMap fm, int n, const unsigned *slotOf, const uint4 *compact,
                                                                    const float *starts, const float *ends, uint2 *out, int *hitIds) {
	__shared__ __align__(32) uint2 rec[kExpandBlock * 7];       // read back in 32-byte pieces
	// one block of kExpandBlock items per CTA; a launch with fewer CTAs than blocks strides over them (the narrow launch
	// that streams records into a peer's HBM next to the walk of the following piece, cm_gpu.cu: tracePieces)
	for (int base = blockIdx.x * kExpandBlock; base < n; base += gridDim.x * kExpandBlock) {
	const int i = base + threadIdx.x;
	if (i < n) {
		const uint4 c = __ldcs(&compact[slotOf ? slotOf[i] : (unsigned)i]);
		const size_t o = 3 * (size_t)i;
		const float s0x = __ldcs(&starts[o]), s0y = __ldcs(&starts[o + 1]), s0z = __ldcs(&starts[o + 2]);
		const float e0x = __ldcs(&ends[o]), e0y = __ldcs(&ends[o + 1]), e0z = __ldcs(&ends[o + 2]);
		const float fraction = __uint_as_float(c.x);
		const int planeRef = (int)c.z, hitId = (int)c.w;
		float4 pl = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
		int2 meta = make_int2(0, 0);
		if (planeRef >= 0) {
			pl = __ldg(&fm.sidePlanes[planeRef]);
			meta = __ldg(&fm.sideMeta[planeRef]);
		}
		const int contents = hitId >= 0 ? __ldg(&fm.brushContents[hitId]) : 0;
		float px, py, pz;
		if (fraction == 1.0f) {
			px = e0x; py = e0y; pz = e0z;
		} else {
			px = s0x + fraction * (e0x - s0x);
			py = s0y + fraction * (e0y - s0y);
			pz = s0z + fraction * (e0z - s0z);
		}
		uint2 *r = rec + 7 * threadIdx.x;
		r[0] = make_uint2(c.y & 1u, (c.y >> 1) & 1u);
		r[1] = make_uint2(__float_as_uint(fraction), __float_as_uint(px));
		r[2] = make_uint2(__float_as_uint(py), __float_as_uint(pz));
		r[3] = make_uint2(__float_as_uint(pl.x), __float_as_uint(pl.y));
		r[4] = make_uint2(__float_as_uint(pl.z), __float_as_uint(pl.w));
		r[5] = make_uint2((unsigned)meta.y & 0xffffu, (unsigned)meta.x);
		r[6] = make_uint2((unsigned)contents, 0u);
		if (hitIds) hitIds[i] = hitId;
	}
	__syncthreads();
	// kExpandBlock * 56 bytes = 896 16-byte pieces; the block's first record is 16-byte aligned (base * 56)
	// kExpandBlock * 56 bytes = 448 32-byte pieces when the block is full; the block's first record is 32-byte aligned
	// when `out` is (base * 56 = blockIdx * 14336), otherwise -- and for the last, partial block -- 8-byte stores
	const int items = min(kExpandBlock, n - base);
	const int words = items * 7;                                 // 8-byte words
	uint2 *dst = out + 7 * (size_t)base;
	int done = 0;
	if ((reinterpret_cast<uintptr_t>(dst) & 31) == 0) {
		const int pieces = words >> 2;
		for (int k = threadIdx.x; k < pieces; k += kExpandBlock) {
			const uint4 a = reinterpret_cast<const uint4 *>(rec)[2 * k], b = reinterpret_cast<const uint4 *>(rec)[2 * k + 1];
			asm volatile("st.global.cs.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
			             :: "l"(dst + 4 * k), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
		}
		done = pieces << 2;
	}
	for (int k = done + threadIdx.x; k < words; k += kExpandBlock) __stcs(&dst[k], rec[k]);
	__syncthreads();                                             // rec is reused by the next block of items
	}
}

#ifndef CMGPU_PREFETCH_CHUNK
#define CMGPU_PREFETCH_CHUNK 1
#endif

// follow a child reference: a node, or the stream of a leaf
template <bool COUNT>
__device__ __forceinline__ int gotoChild(const FastMap &fm, FastLane &L, const int c) {
	if (c < 0) {
		L.k = -1 - c;
		if (COUNT) L.cLeafs++;
		return FM_LEAF;
	}
	L.num = c;
	return FM_NODE;
}

#ifndef CMGPU_FAST_MIN_BLOCKS
#define CMGPU_FAST_MIN_BLOCKS 8
#endif

template <bool COUNT>
__global__ void __launch_bounds__(kBlock, CMGPU_FAST_MIN_BLOCKS)
traceFastKernel(const FastMap fm, const TraceBatch q, const FastHull h) {
	const unsigned lane = threadIdx.x & 31u;
	const unsigned laneLt = (1u << lane) - 1u;
	// Pending far halves live in a global arena of 48-byte entries.  (A local array
	// is interleaved word by word across the lanes of a warp: with a few lanes pushing at different depths every
	// 48-byte entry cost twelve 32-byte sectors of L2 traffic; here it is two.)
	// (entry d of thread t is arena[d * threads + t]: the shallow depths that are used form a dense region)
	const size_t arenaSlots = (size_t)gridDim.x * kBlock;
	FastSeg *const stack0 = q.stackArena + (size_t)blockIdx.x * kBlock + threadIdx.x;
#define stackAt(d) (stack0 + (size_t)(d) * arenaSlots)
	FastLane L;
	int mode = FM_IDLE;
	int chunkNext = 0, chunkEnd = 0;     // warp-uniform
	bool supply = true;                   // warp-uniform: the global cursor may still have items

	for (;;) {
		// ---------------- census: one REDUX over 6-bit counters ------------------
		unsigned key;
		if (mode == FM_IDLE) key = 1u;
		else if (mode == FM_DONE) key = 1u << 6;
		else if (mode == FM_SPLIT) key = 1u << 12;
		else if (mode == FM_BRUSH) key = 1u << 18;
		else if (mode >= FM_PNODE) key = 1u << 24;
		else key = 0u;
		const unsigned census = __reduce_add_sync(0xffffffffu, key);
		const int nIdle = census & 63, nDone = (census >> 6) & 63, nSplit = (census >> 12) & 63;
		const int nBrush = (census >> 18) & 63, nPos = (census >> 24) & 63;
		const int nLight = 32 - (nIdle + nDone + nSplit + nBrush);      // position tests count as light work
		const int gang = min(q.gang, nLight);                          // 0 when only parked lanes are left

		// ---------------- write results and hand out new items --------------------------
		{
			const int nWant = nDone + (supply ? nIdle : 0);
			if (nIdle == 32 && !supply) break;
			if (nWant > 0 && nWant >= min(q.gangFetch, nLight)) {
				if (mode == FM_DONE) {
					finishFast<COUNT>(fm, q, L);
					mode = FM_IDLE;
				}
				const unsigned need = __ballot_sync(0xffffffffu, mode == FM_IDLE);
				const int nNeed = __popc(need);
				if (chunkNext >= chunkEnd && supply) {
					int base = 0;
					if (lane == 0) base = atomicAdd(q.cursor, q.chunk);
					base = __shfl_sync(0xffffffffu, base, 0);
					if (base >= q.n) { supply = false; chunkNext = chunkEnd = 0; }
					else {
						chunkNext = base; chunkEnd = min(base + q.chunk, q.n);
#if CMGPU_PREFETCH_CHUNK
						// pull the chunk's ray records towards L2 while the lanes are still busy with the previous one
						if (q.recs) {
							for (int i = base + (int)lane * 4; i < chunkEnd; i += 128)
								asm volatile("prefetch.global.L2 [%0];" :: "l"(&q.recs[2 * (size_t)i]));
						}
#endif
					}
				}
				if (chunkNext < chunkEnd) {
					if (mode == FM_IDLE) {
						const int my = chunkNext + __popc(need & laneLt);
						if (my < chunkEnd) mode = beginFast<COUNT>(q, h, L, my);
					}
					chunkNext = min(chunkNext + nNeed, chunkEnd);
				} else if (need == 0xffffffffu) {
					break;                                           // nothing in flight, nothing left
				}
			}
		}

		// ---------------- light steps -----------------
		#pragma unroll 1
		for (int rep = 0; rep < q.lightReps; rep++) {
			// Every lane first ASKS for the record its step needs (node, leaf entry or pending segment -- a lane is in
			// one mode, so the three share registers), then the steps run: the three latencies overlap instead of
			// following each other.
			float4 t0 = make_float4(0.0f, 0.0f, 0.0f, 0.0f), t1 = t0, t2 = t0;
			if (mode == FM_NODE) {
				t0 = __ldg(reinterpret_cast<const float4 *>(&fm.nodes[L.num]));
			} else if (mode == FM_LEAF) {
				{ const Pair4 ee = ldg32(&fm.stream[2 * L.k]); t0 = ee.a; t1 = ee.b; }
			} else if (mode == FM_POP && L.sp > 0) {
				const float4 *sp4 = reinterpret_cast<const float4 *>(stackAt(L.sp - 1));
				t0 = __ldcg(&sp4[0]); t1 = __ldcg(&sp4[1]); t2 = __ldcg(&sp4[2]);
			}
			// ---------------- one brush entry of CM_TraceThroughLeaf, cm_trace.c:377-393
			if (mode == FM_LEAF) {
				const float4 e0 = t0, e1 = t1;
				L.k++;
				const int b = __float_as_int(e1.w);
				if (b < 0) {
					mode = FM_POP;
				} else {
					if (COUNT) L.cBrushRefs++;
					const bool want = (__float_as_int(e0.w) & h.mask) != 0 && b != L.last0 && b != L.last1;
					const bool apart = L.hix < e0.x || L.hiy < e0.y || L.hiz < e0.z || L.lox > e1.x || L.loy > e1.y || L.loz > e1.z;
					if (COUNT) L.cBrushBounds += want ? 1u : 0u;
					if (want && !apart) {
						if (!((L.flags & 4) && slabReject(h, L, e0, e1))) { L.pend = b; mode = FM_BRUSH; }
					}
				}
			}
			// ---------------- return to the pending far half of a split (the returns of the reference's recursion)
			else if (mode == FM_POP) {
				if (L.sp == 0) {
					mode = FM_DONE;
				} else {
					L.sp--;
					if (!(L.fraction <= t0.y)) {                // cm_trace.c:449: already hit something nearer
						L.f1 = t0.y; L.f2 = t0.z;
						L.ax = t0.w; L.ay = t1.x; L.az = t1.y; L.bx = t1.z; L.by = t1.w; L.bz = t2.x;
						mode = gotoChild<COUNT>(fm, L, __float_as_int(t0.x));
					}
				}
			}
			// ---------------- one axial node of CM_TraceThroughTree, cm_trace.c:456-486
			else if (mode == FM_NODE) {
				const int4 nd = make_int4(__float_as_int(t0.x), __float_as_int(t0.y), __float_as_int(t0.z), __float_as_int(t0.w));
				if (COUNT) L.cNodes++;
				const bool ax0 = nd.z == 0, ax1 = nd.z == 1;
				const float dist = __int_as_float(nd.w);
				const float t1v = (ax0 ? L.ax : (ax1 ? L.ay : L.az)) - dist;
				const float t2v = (ax0 ? L.bx : (ax1 ? L.by : L.bz)) - dist;
				const float hi = ax0 ? h.nodeHi[0] : (ax1 ? h.nodeHi[1] : h.nodeHi[2]);
				const float lo = ax0 ? h.nodeLo[0] : (ax1 ? h.nodeLo[1] : h.nodeLo[2]);
				const bool front = t1v >= hi && t2v >= hi;
				const bool back = t1v < lo && t2v < lo;
				if (nd.z < 3 && (front || back)) mode = gotoChild<COUNT>(fm, L, front ? nd.x : nd.y);
				else mode = FM_SPLIT;                       // park: split (or a non-axial node) is done by a gang
			}
		}

		// ---------------- a node that is not a plain axial descent, cm_trace.c:456-537 (heavy) -----------------
		if (nSplit > 0 && nSplit >= gang) {
			if (mode == FM_SPLIT) {
				const int4 nd = __ldg(&fm.nodes[L.num]);
				double t1, t2, off;
				if (nd.z < 3) {
					const float dist = __int_as_float(nd.w);
					t1 = (double)(sel3(L.ax, L.ay, L.az, nd.z) - dist);      // float subtract, then widened
					t2 = (double)(sel3(L.bx, L.by, L.bz, nd.z) - dist);
					off = h.isPoint ? 0.0 : (double)sel3(h.p0[0], h.p0[1], h.p0[2], nd.z);
				} else {
					const float4 pl = __ldg(&fm.nodePlanes[L.num]);
					t1 = dotd((double)pl.x, (double)pl.y, (double)pl.z, L.ax, L.ay, L.az) - (double)pl.w;
					t2 = dotd((double)pl.x, (double)pl.y, (double)pl.z, L.bx, L.by, L.bz) - (double)pl.w;
					off = h.isPoint ? 0.0 : 2048.0;
				}
				if (t1 >= off + 1 && t2 >= off + 1) {
					mode = gotoChild<COUNT>(fm, L, nd.x);
				} else if (t1 < -off - 1 && t2 < -off - 1) {
					mode = gotoChild<COUNT>(fm, L, nd.y);
				} else {
					// branch-free form of the three cases t1<t2, t1>t2, t1==t2
					const bool fwd = t1 < t2, eq = !(t1 < t2) && !(t1 > t2);
					const float inv = (float)(1.0 / (eq ? 1.0 : (t1 - t2)));
					const double nfar = fwd ? (t1 + off + kClipEps) : (t1 - off - kClipEps);
					const double nnear = fwd ? (t1 - off + kClipEps) : (t1 + off + kClipEps);
					float ffar = (float)(nfar * (double)inv);
					float fnear = (float)(nnear * (double)inv);
					if (eq) { fnear = 1.0f; ffar = 0.0f; }
					if (fnear < 0) fnear = 0; else if (fnear > 1) fnear = 1;
					if (ffar < 0) ffar = 0; else if (ffar > 1) ffar = 1;
					if (COUNT) L.cSplits++;
					// the far half [mid(ffar), b] waits on the stack; the near half [a, mid(fnear)] goes on.
					// split points are float lerps of the CURRENT segment (cm_trace.c:516-535) and are carried.
					FastSeg far;
					far.num = fwd ? nd.x : nd.y;
					far.f1 = L.f1 + (L.f2 - L.f1) * ffar;
					far.f2 = L.f2;
					far.ax = L.ax + ffar * (L.bx - L.ax);
					far.ay = L.ay + ffar * (L.by - L.ay);
					far.az = L.az + ffar * (L.bz - L.az);
					far.bx = L.bx; far.by = L.by; far.bz = L.bz;
					const float midf = L.f1 + (L.f2 - L.f1) * fnear;
					const float mx = L.ax + fnear * (L.bx - L.ax);
					const float my = L.ay + fnear * (L.by - L.ay);
					const float mz = L.az + fnear * (L.bz - L.az);
					if (L.sp < q.stackDepth) {
						float4 *sp4 = reinterpret_cast<float4 *>(stackAt(L.sp++));
						__stcg(&sp4[0], make_float4(__int_as_float(far.num), far.f1, far.f2, far.ax));
						__stcg(&sp4[1], make_float4(far.ay, far.az, far.bx, far.by));
						__stcg(&sp4[2], make_float4(far.bz, 0.0f, 0.0f, 0.0f));
					} else {
						atomicOr(q.status, 1);
					}
					L.f2 = midf;
					L.bx = mx; L.by = my; L.bz = mz;
					mode = gotoChild<COUNT>(fm, L, fwd ? nd.y : nd.x);
				}
			}
		}

		// ---------------- CM_TraceThroughBrush for the pending brush (heavy) -----
		if (nBrush > 0 && nBrush >= gang) {
			if (mode == FM_BRUSH) {
				const int b = L.pend;
				const bool over = clipBrushFast<COUNT>(fm, h, L, b);
				L.last1 = L.last0; L.last0 = b;
				mode = over ? FM_DONE : FM_LEAF;
			}
		}

		// ---------------- position test: CM_PositionTest / CM_BoxLeafnums_r / CM_TestInLeaf / CM_TestBoxInBrush,
		// cm_trace.c:83-226, cm_test.c:131-156.  The reference lists <= 1024 leafs and then tests them in list order;
		// testing each leaf when it is found is the same sequence of tests.  Binned batches keep these items
		// together (class 2), so whole warps run this block.
		if (nPos > 0) {
			if (mode == FM_PNODE) {
				const int c = L.num;
				if (c < 0) {
					if (L.pend >= kMaxPosLeafs) {
						mode = FM_DONE;
					} else {
						L.pend++;
						L.k = -1 - c;
						if (COUNT) L.cLeafs++;
						mode = FM_PLEAF;
					}
				} else {
					const int4 nd = __ldg(&fm.nodes[c]);
					if (COUNT) L.cNodes++;
					const float lox = L.lox - 1.0f, loy = L.loy - 1.0f, loz = L.loz - 1.0f;
					const float hix = L.hix + 1.0f, hiy = L.hiy + 1.0f, hiz = L.hiz + 1.0f;
					int s;
					const float dist = __int_as_float(nd.w);
					if (nd.z < 3) {                                    // BoxOnPlaneSide, q_math.c:724-758
						if (dist <= sel3(lox, loy, loz, nd.z)) s = 1;
						else if (dist >= sel3(hix, hiy, hiz, nd.z)) s = 2;
						else s = 3;
					} else {
						const float4 pl = __ldg(&fm.nodePlanes[c]);
						float a = 0.0f, bb = 0.0f;
						if (pl.x < 0.0f) { bb = bb + pl.x * hix; a = a + pl.x * lox; } else { a = a + pl.x * hix; bb = bb + pl.x * lox; }
						if (pl.y < 0.0f) { bb = bb + pl.y * hiy; a = a + pl.y * loy; } else { a = a + pl.y * hiy; bb = bb + pl.y * loy; }
						if (pl.z < 0.0f) { bb = bb + pl.z * hiz; a = a + pl.z * loz; } else { a = a + pl.z * hiz; bb = bb + pl.z * loz; }
						s = 0;
						if (a >= dist) s = 1;
						if (bb < dist) s |= 2;
					}
					if (s == 1) {
						L.num = nd.x;
					} else if (s == 2) {
						L.num = nd.y;
					} else {
						if (L.sp < q.stackDepth) __stcg(&stackAt(L.sp++)->num, nd.y); else atomicOr(q.status, 1);
						L.num = nd.x;
					}
				}
			} else if (mode == FM_PLEAF) {
				const Pair4 ee = ldg32(&fm.stream[2 * L.k]);
				const float4 e0 = ee.a, e1 = ee.b;
				L.k++;
				const int b = __float_as_int(e1.w);
				if (b < 0) {
					if (L.sp == 0) mode = FM_DONE;
					else { L.num = __ldcg(&stackAt(--L.sp)->num); mode = FM_PNODE; }
				} else {
					if (COUNT) L.cBrushRefs++;
					const int contents = __float_as_int(e0.w);
					if (contents & h.mask) {
						const Pair4 rr = ldg32(&fm.brushRec[2 * b]);
						const float4 r0 = rr.a, r1 = rr.b;
						const int firstSide = __float_as_int(r0.w);
						const int ns = __float_as_int(r1.w) & 0xffffff;
						if (ns > 0) {
							if (COUNT) L.cBrushBounds++;
							if (!(L.lox > r1.x || L.loy > r1.y || L.loz > r1.z || L.hix < r0.x || L.hiy < r0.y || L.hiz < r0.z)) {
								bool inside = true;
								for (int s = 6; s < ns; s++) {
									const float4 pl = __ldg(&fm.sidePlanes[firstSide + s]);
									if (COUNT) L.cBrushSides++;
									const float cx = pl.x < 0.0f ? h.p0[0] : h.n0[0];
									const float cy = pl.y < 0.0f ? h.p0[1] : h.n0[1];
									const float cz = pl.z < 0.0f ? h.p0[2] : h.n0[2];
									const float distf = pl.w - dotf(cx, cy, cz, pl.x, pl.y, pl.z);    // all-float (cm_trace.c:111)
									const double d1 = dotd((double)L.sx, (double)L.sy, (double)L.sz, pl.x, pl.y, pl.z) - (double)distf;
									if (d1 > 0) { inside = false; break; }
								}
								if (inside) {
									L.flags |= 3;
									L.fraction = 0.0f;
									L.hitId = b;
									mode = FM_DONE;
								}
							}
						}
					}
				}
			}
		}
	}
}

#undef stackAt

}  // namespace cmgpu
