// cm_trace_simple.cuh -- the hot path for SMALL launches: one thread per item, the reference's own loop nest.
//
// The persistent kernels (cm_trace_pool.cuh, cm_trace_fast.cuh) buy throughput with a scheduler round per step; a
// round costs ~1 us, so a launch of a few thousand items -- a game tick, one phase of a rollout -- takes as long as
// its slowest item's ~150 rounds.  Below ~256 Ki items the GPU is not full anyway, and latency is what matters: this
// kernel walks one item per thread with plain loops (same step arithmetic: beginFast, slabReject, clipBrushFast,
// finishFast), no voting, no parking.
#pragma once

#include "cm_trace_fast.cuh"

namespace cmgpu {

#ifndef CMGPU_SIMPLE_MIN_BLOCKS
#define CMGPU_SIMPLE_MIN_BLOCKS 6
#endif
template <bool COUNT, int DEPTH = kMaxDepth>
__global__ void __launch_bounds__(kBlock, CMGPU_SIMPLE_MIN_BLOCKS)
traceSimpleKernel(const FastMap fm, const TraceBatch q, const FastHull h) {
	// Small launches leave most of the GPU's warp slots empty, and a warp is only done when the slowest of its 32 walks
	// is: while the launch fits, only every (1 << spreadLog2)-th lane takes an item -- more warps, each with fewer
	// walks to wait for (1024 items: 144 -> 52 us with a warp per item; from 64 Ki items up there is nothing to gain).
	const int t = blockIdx.x * kBlock + threadIdx.x;
	if (t & ((1 << q.spreadLog2) - 1)) return;
	const int slot = t >> q.spreadLog2;
	if (slot >= q.n) return;
	FastSeg stack[DEPTH];            // local memory: per resident thread, not per item (small launches)
#define stackAt(d) (&stack[d])
	FastLane L;
	const int kind = beginFast<COUNT>(q, h, L, slot);

	if (kind == FM_PNODE) {
		// ---- position test: CM_PositionTest / CM_BoxLeafnums_r / CM_TestInLeaf / CM_TestBoxInBrush (cm_trace.c:83-226,
		// cm_test.c:131-156); leafs are tested as they are found, which is the order of the reference's list
		const float lox = L.lox - 1.0f, loy = L.loy - 1.0f, loz = L.loz - 1.0f;
		const float hix = L.hix + 1.0f, hiy = L.hiy + 1.0f, hiz = L.hiz + 1.0f;
		int c = 0, found = 0;
		bool done = false;
		while (!done) {
			while (c >= 0) {
				const int4 nd = __ldg(&fm.nodes[c]);
				if (COUNT) L.cNodes++;
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
				if (s == 1) c = nd.x;
				else if (s == 2) c = nd.y;
				else {
					if (L.sp < DEPTH) stackAt(L.sp++)->num = nd.y; else atomicOr(q.status, 1);
					c = nd.x;
				}
			}
			if (found >= kMaxPosLeafs) break;
			found++;
			if (COUNT) L.cLeafs++;
			for (int k = -1 - c; ; k++) {
				const Pair4 ee = ldg32(&fm.stream[2 * k]);
				const float4 e0 = ee.a, e1 = ee.b;
				const int b = __float_as_int(e1.w);
				if (b < 0) break;
				if (COUNT) L.cBrushRefs++;
				if (!(__float_as_int(e0.w) & h.mask)) continue;
				const Pair4 rr = ldg32(&fm.brushRec[2 * b]);
				const float4 r0 = rr.a, r1 = rr.b;
				const int firstSide = __float_as_int(r0.w);
				const int ns = __float_as_int(r1.w) & 0xffffff;
				if (ns == 0) continue;
				if (COUNT) L.cBrushBounds++;
				if (L.lox > r1.x || L.loy > r1.y || L.loz > r1.z || L.hix < r0.x || L.hiy < r0.y || L.hiz < r0.z) continue;
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
					done = true;
					break;
				}
			}
			if (done || L.sp == 0) break;
			c = stackAt(--L.sp)->num;
		}
	} else {
		// ---- sweep: CM_TraceThroughTree / CM_TraceThroughLeaf / CM_TraceThroughBrush (cm_trace.c:261-538)
		int c = 0;
		for (;;) {
			// down to a leaf, splitting on the way (cm_trace.c:456-537)
			while (c >= 0) {
				const int4 nd = __ldg(&fm.nodes[c]);
				if (COUNT) L.cNodes++;
				if (nd.z < 3) {
					const float dist = __int_as_float(nd.w);
					const float t1 = sel3(L.ax, L.ay, L.az, nd.z) - dist, t2 = sel3(L.bx, L.by, L.bz, nd.z) - dist;
					const float hi = sel3(h.nodeHi[0], h.nodeHi[1], h.nodeHi[2], nd.z), lo = sel3(h.nodeLo[0], h.nodeLo[1], h.nodeLo[2], nd.z);
					if (t1 >= hi && t2 >= hi) { c = nd.x; continue; }
					if (t1 < lo && t2 < lo) { c = nd.y; continue; }
				}
				double t1, t2, off;
				if (nd.z < 3) {
					const float dist = __int_as_float(nd.w);
					t1 = (double)(sel3(L.ax, L.ay, L.az, nd.z) - dist);      // float subtract, then widened
					t2 = (double)(sel3(L.bx, L.by, L.bz, nd.z) - dist);
					off = h.isPoint ? 0.0 : (double)sel3(h.p0[0], h.p0[1], h.p0[2], nd.z);
				} else {
					const float4 pl = __ldg(&fm.nodePlanes[c]);
					t1 = dotd((double)pl.x, (double)pl.y, (double)pl.z, L.ax, L.ay, L.az) - (double)pl.w;
					t2 = dotd((double)pl.x, (double)pl.y, (double)pl.z, L.bx, L.by, L.bz) - (double)pl.w;
					off = h.isPoint ? 0.0 : 2048.0;
				}
				if (t1 >= off + 1 && t2 >= off + 1) { c = nd.x; continue; }
				if (t1 < -off - 1 && t2 < -off - 1) { c = nd.y; continue; }
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
				if (L.sp < DEPTH) {
					float4 *sp4 = reinterpret_cast<float4 *>(stackAt(L.sp++));
					sp4[0] = make_float4(__int_as_float(fwd ? nd.x : nd.y), L.f1 + (L.f2 - L.f1) * ffar, L.f2, L.ax + ffar * (L.bx - L.ax));
					sp4[1] = make_float4(L.ay + ffar * (L.by - L.ay), L.az + ffar * (L.bz - L.az), L.bx, L.by);
					sp4[2] = make_float4(L.bz, 0.0f, 0.0f, 0.0f);
				} else {
					atomicOr(q.status, 1);
				}
				const float midf = L.f1 + (L.f2 - L.f1) * fnear;
				const float mx = L.ax + fnear * (L.bx - L.ax), my = L.ay + fnear * (L.by - L.ay), mz = L.az + fnear * (L.bz - L.az);
				L.f2 = midf;
				L.bx = mx; L.by = my; L.bz = mz;
				c = fwd ? nd.y : nd.x;
			}
			// the leaf's brush stream (cm_trace.c:377-393)
			if (COUNT) L.cLeafs++;
			bool over = false;
			for (int k = -1 - c; ; k++) {
				const Pair4 ee = ldg32(&fm.stream[2 * k]);
				const float4 e0 = ee.a, e1 = ee.b;
				const int b = __float_as_int(e1.w);
				if (b < 0) break;
				if (COUNT) L.cBrushRefs++;
				if (!(__float_as_int(e0.w) & h.mask) || b == L.last0 || b == L.last1) continue;
				if (COUNT) L.cBrushBounds++;
				if (L.hix < e0.x || L.hiy < e0.y || L.hiz < e0.z || L.lox > e1.x || L.loy > e1.y || L.loz > e1.z) continue;
				if ((L.flags & 4) && slabReject(h, L, e0, e1)) continue;
				over = clipBrushFast<COUNT>(fm, h, L, b);
				L.last1 = L.last0; L.last0 = b;
				if (over) break;
			}
			if (over) break;
			// back to the pending far halves (cm_trace.c:449: skip those that start beyond the hit)
			bool got = false;
			while (L.sp > 0) {
				const float4 *sp4 = reinterpret_cast<const float4 *>(stackAt(--L.sp));
				const float4 t0 = sp4[0];
				if (L.fraction <= t0.y) continue;
				const float4 t1 = sp4[1], t2 = sp4[2];
				L.f1 = t0.y; L.f2 = t0.z;
				L.ax = t0.w; L.ay = t1.x; L.az = t1.y; L.bx = t1.z; L.by = t1.w; L.bz = t2.x;
				c = __float_as_int(t0.x);
				got = true;
				break;
			}
			if (!got) break;
		}
	}
	finishFast<COUNT>(fm, q, L);
#undef stackAt
}

}  // namespace cmgpu
