// cm_trace_pool.cuh -- the hot kernel, second cut: same steps as traceFastKernel (cm_trace_fast.cuh), other scheduling.
//
// traceFastKernel keeps ONE item per lane in registers; with six kinds of step in flight a block of code runs with
// ~11 of 32 lanes on average, whatever the gang sizes (measured, and reproduced by a scheduling model over the
// oracle's per-item step sequences).  Here every lane owns K items whose state lives in shared memory
// (structure of arrays, [field][k][lane]: conflict-free, no cross-lane traffic, nothing shared between warps).  Each
// round the warp counts how many lanes own an item waiting for each kind of step and runs every kind that enough
// lanes can run; a lane runs the step on its first item of that kind.  Light steps are batched: a node step descends
// up to kNodeLevels levels through a per-hull table of thresholds on the coordinates (nodeThresholdKernel), a leaf step
// looks at kLeafEntries stream entries and runs the float slab pre-test on the ones that pass the bounds test, so that
// only brushes that need the exact clip cost a round of their own.  The model gives 17 / 21 / 23 active
// lanes per instruction for K = 2 / 3 / 4; K = 3 with 7 CTAs per SM is the measured optimum (more items per lane
// leave too few warps and too little L1 for the same shared memory).
//
// Results are bit-identical to traceFastKernel: the step arithmetic is the same code.
#pragma once

#include "cm_trace_fast.cuh"

namespace cmgpu {

#ifndef CMGPU_POOL_LAST2
#define CMGPU_POOL_LAST2 0
#endif
#ifndef CMGPU_NODE_LEVELS
#define CMGPU_NODE_LEVELS 3
#endif
#ifndef CMGPU_INLINE_SLAB
#define CMGPU_INLINE_SLAB 1
#endif
#ifndef CMGPU_LEAF_ENTRIES
#define CMGPU_LEAF_ENTRIES 4
#endif
#ifndef CMGPU_SLAB_HOIST
#define CMGPU_SLAB_HOIST 1
#endif
constexpr int kLeafEntries = CMGPU_LEAF_ENTRIES;  // stream entries a lane looks at per repeat of the leaf step
constexpr int kNodeLevels = CMGPU_NODE_LEVELS;   // tree levels a lane may descend per repeat of the node step

// step kinds an item can wait for (one-hot bit m-1 of its mode byte).  PM_POS is the position test; bit 3 of the item's
// flags says whether it is in the tree part (CM_BoxLeafnums_r) or the leaf part (CM_TestInLeaf).
enum PoolMode : int { PM_EMPTY = 0, PM_NODE, PM_LEAF, PM_SLAB, PM_POP, PM_SPLIT, PM_BRUSH, PM_DONE, PM_POS, PM_COUNT };

// shared-memory fields of one item (32-bit words)
enum PoolField : int {
	PF_SX = 0, PF_SY, PF_SZ, PF_EX, PF_EY, PF_EZ, PF_FRACTION, PF_FLAGS /* flags | sp << 8 */,
	PF_CUR /* node number or leaf-stream cursor */,
	PF_F1, PF_F2,
	PF_AX, PF_AY, PF_AZ, PF_BX, PF_BY, PF_BZ,
	PF_LAST0,
#if CMGPU_POOL_LAST2
	PF_LAST1,
#endif
	PF_COUNT
	// (the brush waiting to be clipped is the one of stream entry CUR-1; the position test's leaf count lives in
	// bits 16.. of FLAGS, the stack depth in bits 8..15)
};

// what an item only needs when its record is written lives in global memory (one 16-byte header per item slot)
struct __align__(16) PoolHdr { int slot, planeRef, hitId, pad; };

// The step kinds of a lane's K (<= 4) items live in one register, kind-major: nibble m-1 holds one bit per item that
// waits for kind m.  "Do I own an item of kind m, and which" is a shift, a mask and a find-first-set; moving item j
// to another kind clears bit j of every nibble and sets one bit.
template <int K>
__device__ __forceinline__ int firstInMode(const unsigned modes, const int m) {
	return __ffs((int)((modes >> (4 * (m - 1))) & 0xfu)) - 1;
}

__device__ __forceinline__ unsigned withMode(const unsigned modes, const int j, const int m) {
	const unsigned clr = 0x11111111u << j;
	const unsigned set = m ? ((1u << (4 * (m - 1))) << j) : 0u;
	return (modes & ~clr) | set;
}

#ifndef CMGPU_POOL_K
#define CMGPU_POOL_K 3
#endif
#ifndef CMGPU_POOL_MIN_BLOCKS
#define CMGPU_POOL_MIN_BLOCKS 7
#endif

// ---- the batch's hull folded into the nodes --------------------------------------------------------------------
// An axial node asks whether fl(p - dist) >= hi (both ends: in front) or fl(p - dist) < lo (both ends: behind), with
// hi / lo the float images of offset+1 / -offset-1 for the node's axis (makeFastHull).  Float subtraction is monotone
// in p, so each question is a threshold on p itself: p >= tHi with tHi the SMALLEST float whose difference reaches
// hi, and p < tLo with tLo the smallest float whose difference reaches lo.  The table holds, per node,
// {child0 << 2 | type, child1, tHi, tLo}; a non-axial node gets (+inf, -inf) and so never passes as a plain descent.
// Built once per (map, hull) -- 16 bytes per node -- and saves the node step its two subtractions and the per-axis
// selection of hi and lo.
// (bisection over the floats in their numeric order: near dist = -want the answer lies astronomically many
// representable values away from want + dist, so stepping with nextafterf does not get there)
__device__ __forceinline__ unsigned floatOrderKey(const float f) {
	const unsigned u = __float_as_uint(f);
	return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float floatFromOrderKey(const unsigned k) {
	return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}
__device__ __forceinline__ float smallestReaching(const float dist, const float want) {
	unsigned lo = floatOrderKey(-3.0e38f), hi = floatOrderKey(3.0e38f);   // predicate false at lo, true at hi
	while (hi - lo > 1u) {
		const unsigned mid = lo + ((hi - lo) >> 1);
		if ((floatFromOrderKey(mid) - dist) >= want) hi = mid; else lo = mid;
	}
	return floatFromOrderKey(hi);
}

struct NodeHull { float hi[3], lo[3]; };

__global__ void __launch_bounds__(256) nodeThresholdKernel(const int4 *nodes, const int n, const NodeHull h, int4 *out, int *status) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const int4 nd = nodes[i];
	float tHi = INFINITY, tLo = -INFINITY;
	if (nd.z < 3) {
		const float dist = __int_as_float(nd.w);
		const float hi = nd.z == 0 ? h.hi[0] : (nd.z == 1 ? h.hi[1] : h.hi[2]);
		const float lo = nd.z == 0 ? h.lo[0] : (nd.z == 1 ? h.lo[1] : h.lo[2]);
		tHi = smallestReaching(dist, hi);
		tLo = smallestReaching(dist, lo);
		// the boundary must be exact: the float just below fails, the threshold itself passes (else: no table for this map)
		const bool okHi = (tHi - dist) >= hi && !((nextafterf(tHi, -INFINITY) - dist) >= hi) && fabsf(tHi) < 1.0e38f;
		const bool okLo = (tLo - dist) >= lo && !((nextafterf(tLo, -INFINITY) - dist) >= lo) && fabsf(tLo) < 1.0e38f;
		if (!okHi || !okLo) atomicOr(status, 2);
	}
	if (nd.x >= (1 << 29) || nd.x < -(1 << 29)) atomicOr(status, 2);
	out[i] = make_int4((int)(((unsigned)nd.x << 2) | (unsigned)(nd.z & 3)), nd.y, __float_as_int(tHi), __float_as_int(tLo));
}

// ---- entry grid (EntryGrid, cm_kernels.cuh) ----------------------------------------------------------------------
__device__ __forceinline__ int entryCellAxis(const float p, const float lo, const float inv, const int cells) {
	const float t = (p - lo) * inv;                    // two separately rounded operations: monotone in p
	const int c = (int)t;                              // truncation: monotone too
	return min(max(c, 0), cells - 1);
}

// smallest float in [-kEntryMaxCoord, kEntryMaxCoord] whose cell along the axis is >= want (kEntryMaxCoord's
// successor when there is none): bisection over the floats in numeric order, like smallestReaching
__device__ __forceinline__ float entryCellStart(const float glo, const float ginv, const int cells, const int want) {
	if (entryCellAxis(-kEntryMaxCoord, glo, ginv, cells) >= want) return -kEntryMaxCoord;
	if (entryCellAxis(kEntryMaxCoord, glo, ginv, cells) < want) return nextafterf(kEntryMaxCoord, INFINITY);
	unsigned lo = floatOrderKey(-kEntryMaxCoord), hi = floatOrderKey(kEntryMaxCoord);   // predicate false at lo, true at hi
	while (hi - lo > 1u) {
		const unsigned mid = lo + ((hi - lo) >> 1);
		if (entryCellAxis(floatFromOrderKey(mid), glo, ginv, cells) >= want) hi = mid; else lo = mid;
	}
	return floatFromOrderKey(hi);
}

// One thread per cell.  The cell's points along an axis are the floats [first, last] (within +-kEntryMaxCoord; items
// beyond that do not use the grid); a node is passed by a plain descent of EVERY point of the cell when first >= tHi
// (front) or last < tLo (back) -- the very comparisons the node step makes.  The walk stops at the first node that
// some point of the cell does not pass that way, at a leaf, or after kEntryLevels levels.
__global__ void __launch_bounds__(256) entryGridKernel(const int4 *tnodes, const EntryGrid g, uint2 *out, int *status) {
	const int cell = blockIdx.x * blockDim.x + threadIdx.x;
	const int total = g.cells[0] * g.cells[1] * g.cells[2];
	if (cell >= total) return;
	const int ix = cell % g.cells[0], iy = (cell / g.cells[0]) % g.cells[1], iz = cell / (g.cells[0] * g.cells[1]);
	float first[3], last[3];
	const int idx[3] = {ix, iy, iz};
	bool ok = true;
	#pragma unroll
	for (int a = 0; a < 3; a++) {
		first[a] = entryCellStart(g.lo[a], g.inv[a], g.cells[a], idx[a]);
		const float next = entryCellStart(g.lo[a], g.inv[a], g.cells[a], idx[a] + 1);
		last[a] = nextafterf(next, -INFINITY);
		// the boundaries must be exact (else: no grid for this map): both ends map to this cell, their outer neighbours do not
		if (first[a] <= last[a]) {
			ok = ok && entryCellAxis(first[a], g.lo[a], g.inv[a], g.cells[a]) == idx[a] && entryCellAxis(last[a], g.lo[a], g.inv[a], g.cells[a]) == idx[a];
			if (first[a] > -kEntryMaxCoord) ok = ok && entryCellAxis(nextafterf(first[a], -INFINITY), g.lo[a], g.inv[a], g.cells[a]) < idx[a];
			if (last[a] < kEntryMaxCoord) ok = ok && entryCellAxis(next, g.lo[a], g.inv[a], g.cells[a]) > idx[a];
		}
	}
	if (!ok) atomicOr(status, 2);
	int node = 0, depth = 0;
	unsigned path = 0u;
	const bool empty = first[0] > last[0] || first[1] > last[1] || first[2] > last[2];    // no float maps here: never looked up
	while (!empty && node >= 0 && depth < kEntryLevels) {
		const int4 nd = tnodes[node];
		const int ax = nd.x & 3;
		if (ax == 3) break;
		const float tHi = __int_as_float(nd.z), tLo = __int_as_float(nd.w);
		const float f = ax == 0 ? first[0] : (ax == 1 ? first[1] : first[2]);
		const float l = ax == 0 ? last[0] : (ax == 1 ? last[1] : last[2]);
		int next;
		if (f >= tHi) { next = nd.x >> 2; path = (path << 1) | 1u; }
		else if (l < tLo) { next = nd.y; path = path << 1; }
		else break;
		node = next;
		depth++;
	}
	out[cell] = make_uint2(path | ((unsigned)depth << 16), (unsigned)node);
}

// where the walk of the sweep a -> b starts (a node, or a leaf-stream reference < 0)
__device__ __forceinline__ int entryLookup(const EntryGrid &g, const float ax, const float ay, const float az,
                                           const float bx, const float by, const float bz) {
	const int ca = (entryCellAxis(az, g.lo[2], g.inv[2], g.cells[2]) * g.cells[1] + entryCellAxis(ay, g.lo[1], g.inv[1], g.cells[1])) * g.cells[0]
	               + entryCellAxis(ax, g.lo[0], g.inv[0], g.cells[0]);
	const int cb = (entryCellAxis(bz, g.lo[2], g.inv[2], g.cells[2]) * g.cells[1] + entryCellAxis(by, g.lo[1], g.inv[1], g.cells[1])) * g.cells[0]
	               + entryCellAxis(bx, g.lo[0], g.inv[0], g.cells[0]);
	const uint2 ra = __ldg(&g.cellRec[ca]);
	const uint2 rb = __ldg(&g.cellRec[cb]);
	const int da = (int)(ra.x >> 16), db = (int)(rb.x >> 16);
	const unsigned pa = ra.x & 0xffffu, pb = rb.x & 0xffffu;
	const int m = min(da, db);
	const unsigned xa = pa >> (da - m), xb = pb >> (db - m);
	const unsigned x = xa ^ xb;
	const int L = x ? m - (32 - __clz((int)x)) : m;      // levels the two paths share
	if (L == da) return (int)ra.y;                        // a's node is b's node or one of its ancestors
	if (L == db) return (int)rb.y;
	return __ldg(&g.top[(1 << L) - 1 + (int)(xa >> (m - L))]);
}

template <bool COUNT, int K, bool THR>
__global__ void __launch_bounds__(kBlock, CMGPU_POOL_MIN_BLOCKS)
tracePoolKernel(const FastMap fm, const TraceBatch q, const FastHull h) {
	extern __shared__ float poolMem[];
	const unsigned lane = threadIdx.x & 31u;
	const unsigned laneLt = (1u << lane) - 1u;
	const int warpInBlock = threadIdx.x >> 5;
	float *const P = poolMem + (size_t)warpInBlock * (PF_COUNT * K * 32) + lane;
#define FLD(f, j) P[((f) * K + (j)) * 32]
#define FLDI(f, j) __float_as_int(FLD(f, j))
#define SETI(f, j, v) FLD(f, j) = __int_as_float(v)
	const size_t globalWarp = (size_t)blockIdx.x * (kBlock / 32) + warpInBlock;
	// traversal stacks, depth-major: the entries of neighbouring slots at one depth share cache lines, so the few shallow
	// depths that are ever used are a dense, L2-sized region.  An entry is 36 bytes; its first eight words are one whole
	// 32-byte sector of arena32[d][slot] (child, f1, f2, a.xyz, b.xy -- never a partial-sector write), the ninth (b.z)
	// lives in arenaZ[d][slot], where the lanes of a warp are neighbours.
	const size_t arenaSlots = (size_t)gridDim.x * kBlock * K;
	float4 *const arena32 = reinterpret_cast<float4 *>(q.stackArena) + 2 * ((globalWarp * K) * 32 + lane);      // item j: + 2 * j * 32
	float *const arenaZ = reinterpret_cast<float *>(reinterpret_cast<float4 *>(q.stackArena) + 2 * (size_t)q.stackDepth * arenaSlots)
	                      + (globalWarp * K) * 32 + lane;
#define STACKAT(j, d) (arena32 + 2 * ((size_t)(d) * arenaSlots + (size_t)(j) * 32))
#define STACKZ(j, d) (arenaZ + (size_t)(d) * arenaSlots + (size_t)(j) * 32)
	PoolHdr *const hdrs = reinterpret_cast<PoolHdr *>(q.poolHdr) + (globalWarp * K) * 32 + lane;
#define HDR(j) hdrs[(j) * 32]

	unsigned modes = 0u;
	#pragma unroll
	for (int j = 0; j < K; j++) { modes = withMode(modes, j, PM_DONE); HDR(j).slot = -1; }
	int chunkNext = 0, chunkEnd = 0;     // warp-uniform
	bool supply = true;                   // warp-uniform
	unsigned cNodes = 0, cLeafs = 0, cBrushRefs = 0, cBrushBounds = 0, cBrushSides = 0, cCross = 0, cSplits = 0, cTraces = 0;

	// Return to the pending far half of a split (cm_trace.c:449, 516-537): the new step kind of item j.  PM_POP again
	// when the half is already behind the nearest hit -- the next one is tried in a later round.
	auto popHalf = [&](const int j) -> int {
		const int fl = FLDI(PF_FLAGS, j);
		const int sp = (fl >> 8) & 0xff;
		if (sp == 0) return PM_DONE;
		const float4 *sp4 = STACKAT(j, sp - 1);
		const unsigned long long keep = l2KeepPolicy();
		const Pair4 tt = ldKeep32B(sp4, keep);
		const float4 t0 = tt.a, t1 = tt.b;
		const float t2x = ldKeepF(STACKZ(j, sp - 1), keep);
		SETI(PF_FLAGS, j, fl - 256);
		if (FLD(PF_FRACTION, j) <= t0.y) return PM_POP;            // cm_trace.c:449: already hit something nearer
		FLD(PF_F1, j) = t0.y; FLD(PF_F2, j) = t0.z;
		FLD(PF_AX, j) = t0.w; FLD(PF_AY, j) = t1.x; FLD(PF_AZ, j) = t1.y;
		FLD(PF_BX, j) = t1.z; FLD(PF_BY, j) = t1.w; FLD(PF_BZ, j) = t2x;
		const int c = __float_as_int(t0.x);
		if (c < 0) { SETI(PF_CUR, j, -1 - c); if (COUNT) cLeafs++; return PM_LEAF; }
		SETI(PF_CUR, j, c);
		return PM_NODE;
	};

	for (;;) {
		// ---------------- vote: which kind of step can the most lanes run? ------------------
		// How many lanes own an item waiting for each kind: the presence bits of a lane are spread into byte-wide
		// counters (4 per word) and summed over the warp with two REDUX.  Every kind that at least `thr` lanes can
		// run (a share of the best count) is run this round.
		int cNode, cLeaf, cSlab, cPop, cSplit, cBrush, cDone, cPos, thr, thrBrush, thrDone;
		{
			// bit 4k of p: nibble k is not empty.  (x | x << 12) & 0x01010101 puts kinds 0, 2, 1, 3 of a half into bytes 0..3
			unsigned p = modes | (modes >> 1);
			p = (p | (p >> 2)) & 0x11111111u;
			const unsigned xa = p & 0xffffu, xb = p >> 16;
			const unsigned ka = (xa | (xa << 12)) & 0x01010101u;
			const unsigned kb = (xb | (xb << 12)) & 0x01010101u;
			const unsigned ca = __reduce_add_sync(0xffffffffu, ka), cb = __reduce_add_sync(0xffffffffu, kb);
			cNode = ca & 0xff; cSlab = (ca >> 8) & 0xff; cLeaf = (ca >> 16) & 0xff; cPop = ca >> 24;
			cSplit = cb & 0xff; cDone = (cb >> 8) & 0xff; cBrush = (cb >> 16) & 0xff; cPos = cb >> 24;
			const int best = max(max(max(cNode, cLeaf), max(cSlab, cPop)), max(max(cSplit, cBrush), max(cDone, cPos)));
			if (best == 0) break;                            // every slot empty
			thr = max(1, (best * q.gang) >> 3);
			thrBrush = max(1, (best * q.gangBrush) >> 3);
			thrDone = max(1, (best * q.gangFetch) >> 3);
		}

		if (cDone >= thrDone) {
			// ---------------- write results, take new items (cm_trace.c:545-686) -----------------
			const int j = firstInMode<K>(modes, PM_DONE);
			PoolHdr hd;
			hd.slot = -1;
			if (j >= 0) {
				const float4 v = ldKeep(reinterpret_cast<const float4 *>(&HDR(j)), l2KeepPolicy());
				hd.slot = __float_as_int(v.x); hd.planeRef = __float_as_int(v.y); hd.hitId = __float_as_int(v.z); hd.pad = __float_as_int(v.w);
			}
			if (hd.slot >= 0) {
				FastLane L;
				L.slot = q.compactByItem ? hd.pad : hd.slot; L.fraction = FLD(PF_FRACTION, j); L.flags = FLDI(PF_FLAGS, j) & 0xff;
				L.planeRef = hd.planeRef; L.hitId = hd.hitId;
				L.cNodes = L.cLeafs = L.cBrushRefs = L.cBrushBounds = L.cBrushSides = L.cCross = L.cSplits = 0;
				finishFast<false>(fm, q, L);
				if (COUNT) cTraces++;
			}
			const unsigned need = __ballot_sync(0xffffffffu, j >= 0);
			const int nNeed = __popc(need);
			if (chunkNext >= chunkEnd && supply) {
				int base = 0;
				if (lane == 0) base = atomicAdd(q.cursor, q.chunk);
				base = __shfl_sync(0xffffffffu, base, 0);
				if (base >= q.n) { supply = false; chunkNext = chunkEnd = 0; }
				else {
					chunkNext = base; chunkEnd = min(base + q.chunk, q.n);
					if (q.recs) {
						for (int i = base + (int)lane * 4; i < chunkEnd; i += 128)
							asm volatile("prefetch.global.L2 [%0];" :: "l"(&q.recs[2 * (size_t)i]));
					}
				}
			}
			if (j >= 0) {
				const int my = chunkNext + __popc(need & laneLt);
				if (my < chunkEnd) {
					FastLane L;
					const int m = beginFast<false>(q, h, L, my);
					hd.slot = L.slot; hd.planeRef = -1; hd.hitId = -1; hd.pad = 0;
					stKeep(reinterpret_cast<float4 *>(&HDR(j)), make_float4(__int_as_float(hd.slot), __int_as_float(-1), __int_as_float(-1), __int_as_float(L.item)), l2KeepPolicy());
					FLD(PF_SX, j) = L.sx; FLD(PF_SY, j) = L.sy; FLD(PF_SZ, j) = L.sz;
					FLD(PF_EX, j) = L.ex; FLD(PF_EY, j) = L.ey; FLD(PF_EZ, j) = L.ez;
					FLD(PF_FRACTION, j) = 1.0f; SETI(PF_FLAGS, j, L.flags);
					int cur0 = 0;
					int m0 = m == FM_PNODE ? PM_POS : PM_NODE;
					if (THR && !COUNT && q.entry.cellRec && m != FM_PNODE) {
						// skip the plain descents at the top of the tree (EntryGrid); NaNs fail the comparisons
						const bool inGrid = fabsf(L.sx) <= kEntryMaxCoord && fabsf(L.sy) <= kEntryMaxCoord && fabsf(L.sz) <= kEntryMaxCoord &&
						                    fabsf(L.ex) <= kEntryMaxCoord && fabsf(L.ey) <= kEntryMaxCoord && fabsf(L.ez) <= kEntryMaxCoord;
						if (inGrid) {
							const int c = entryLookup(q.entry, L.sx, L.sy, L.sz, L.ex, L.ey, L.ez);
							if (c < 0) { cur0 = -1 - c; m0 = PM_LEAF; } else cur0 = c;
						}
					}
					SETI(PF_CUR, j, cur0);
					FLD(PF_F1, j) = 0.0f; FLD(PF_F2, j) = 1.0f;
					FLD(PF_AX, j) = L.sx; FLD(PF_AY, j) = L.sy; FLD(PF_AZ, j) = L.sz;
					FLD(PF_BX, j) = L.ex; FLD(PF_BY, j) = L.ey; FLD(PF_BZ, j) = L.ez;
					SETI(PF_LAST0, j, -1);
#if CMGPU_POOL_LAST2
					SETI(PF_LAST1, j, -1);
#endif
					modes = withMode(modes, j, m0);
				} else {
					stKeep32(&HDR(j).slot, -1, l2KeepPolicy());        // nothing to write back next time
					if (!supply) modes = withMode(modes, j, PM_EMPTY); // else: stays PM_DONE and asks again
				}
			}
			chunkNext = min(chunkNext + nNeed, chunkEnd);
		}
		if (cNode >= thr) {
			// ---------------- axial nodes of CM_TraceThroughTree, cm_trace.c:456-486 -----------------
			// up to kNodeLevels tree levels per repeat: a plain descent to another node goes straight on
			#pragma unroll 1
			for (int rep = 0; rep < q.lightReps; rep++) {
				const int j = firstInMode<K>(modes, PM_NODE);
				if (!__any_sync(0xffffffffu, j >= 0)) break;
				if (j >= 0) {
					int num = FLDI(PF_CUR, j);
					int next = PM_NODE;
					#pragma unroll
					for (int lvl = 0; lvl < kNodeLevels; lvl++) {
						if (next == PM_NODE) {
							bool go, front;
							int c0, c1;
							if (THR) {
								// thresholds on the coordinates themselves (nodeThresholdKernel); type 3 reads the field
								// after the axis triple and compares it with +-inf: never a plain descent
								const int4 nd = __ldg(&q.tnodes[num]);
								const int ax = nd.x & 3;
								const float tHi = __int_as_float(nd.z), tLo = __int_as_float(nd.w);
								const float a = FLD(PF_AX + ax, j), b = FLD(PF_BX + ax, j);
								front = a >= tHi && b >= tHi;
								go = front || (a < tLo && b < tLo);
								c0 = nd.x >> 2; c1 = nd.y;
							} else {
								const int4 nd = __ldg(&fm.nodes[num]);
								const int ax = nd.z < 3 ? nd.z : 0;
								const float dist = __int_as_float(nd.w);
								const float t1 = FLD(PF_AX + ax, j) - dist;
								const float t2 = FLD(PF_BX + ax, j) - dist;
								const float hi = ax == 0 ? h.nodeHi[0] : (ax == 1 ? h.nodeHi[1] : h.nodeHi[2]);
								const float lo = ax == 0 ? h.nodeLo[0] : (ax == 1 ? h.nodeLo[1] : h.nodeLo[2]);
								front = t1 >= hi && t2 >= hi;
								const bool back = t1 < lo && t2 < lo;
								go = nd.z < 3 && (front || back);      // else: split or non-axial node
								c0 = nd.x; c1 = nd.y;
							}
							if (COUNT) cNodes++;
							const int c = front ? c0 : c1;
							const bool leaf = go && c < 0;
							if (go) num = c < 0 ? -1 - c : c;
							if (COUNT) cLeafs += leaf ? 1u : 0u;
							next = !go ? PM_SPLIT : (leaf ? PM_LEAF : PM_NODE);
						}
					}
					// one update of the cursor and of the step kind
					SETI(PF_CUR, j, num);
					if (next != PM_NODE) modes = withMode(modes, j, next);
				}
			}
		}
		if (cLeaf >= thr) {
			// ---------------- brush entries of CM_TraceThroughLeaf, cm_trace.c:377-393 -----------------
			#pragma unroll 1
			for (int rep = 0; rep < q.lightReps; rep++) {
				const int j = firstInMode<K>(modes, PM_LEAF);
				if (!__any_sync(0xffffffffu, j >= 0)) break;
				if (j >= 0) {
					const int k = FLDI(PF_CUR, j);
					// kLeafEntries entries per repeat (the stream has spare sentinels at its end, so the loads are always
					// in bounds); an entry is only looked at when the ones before it neither end the leaf nor need clipping
					float4 e0[kLeafEntries], e1[kLeafEntries];
					#pragma unroll
					for (int i = 0; i < kLeafEntries; i++) {
						const Pair4 ee = ldg32(&fm.stream[2 * (k + i)]);
						e0[i] = ee.a; e1[i] = ee.b;
					}
					const int last0 = FLDI(PF_LAST0, j);
#if CMGPU_POOL_LAST2
					const int last1 = FLDI(PF_LAST1, j);
#else
					const int last1 = last0;
#endif
					const float sx = FLD(PF_SX, j), sy = FLD(PF_SY, j), sz = FLD(PF_SZ, j);
					const float ex = FLD(PF_EX, j), ey = FLD(PF_EY, j), ez = FLD(PF_EZ, j);
					// tw.bounds (cm_trace.c:628-636), recomputed instead of stored
					const float lox = (sx < ex ? sx : ex) + h.n0[0], hix = (sx < ex ? ex : sx) + h.p0[0];
					const float loy = (sy < ey ? sy : ey) + h.n0[1], hiy = (sy < ey ? ey : sy) + h.p0[1];
					const float loz = (sz < ez ? sz : ez) + h.n0[2], hiz = (sz < ez ? ez : sz) + h.p0[2];
					int adv = 0, b = 0;
					bool stopped = false, clip = false;
#if CMGPU_INLINE_SLAB
					// the float slab pre-test (slabReject) right here, on the entry that is in registers anyway: an entry it
					// rejects is a certain no-op and the scan goes on, instead of two more rounds (SLAB, then LEAF again)
#if CMGPU_SLAB_HOIST
					const SlabItem S = slabItem(h, sx, sy, sz, ex, ey, ez, FLD(PF_FRACTION, j));
#else
					FastLane S;
					S.sx = sx; S.sy = sy; S.sz = sz; S.ex = ex; S.ey = ey; S.ez = ez;
					S.fraction = FLD(PF_FRACTION, j);
#endif
					const bool slabOn = (FLDI(PF_FLAGS, j) & 4) != 0;
#endif
					#pragma unroll
					for (int i = 0; i < kLeafEntries; i++) {
						const int bi = __float_as_int(e1[i].w);
						const bool want = bi >= 0 && (__float_as_int(e0[i].w) & h.mask) != 0 && bi != last0 && bi != last1;
						const bool apart = hix < e0[i].x || hiy < e0[i].y || hiz < e0[i].z || lox > e1[i].x || loy > e1[i].y || loz > e1[i].z;
#if CMGPU_INLINE_SLAB
						bool clipi = want && !apart;                                   // CM_BoundsIntersect passed
#if CMGPU_SLAB_HOIST && CMGPU_SLAB_UNCOND
						// every lane runs the test on every entry (a warp skips it only when no lane has a candidate, which
						// hardly ever happens): no divergent region per entry
#if CMGPU_SLAB_F32X2
						if (slabRejectPre2(S, e0[i], e1[i]) && slabOn) clipi = false;
#else
						if (slabRejectPre(S, e0[i], e1[i]) && slabOn) clipi = false;
#endif
#elif CMGPU_SLAB_HOIST && CMGPU_SLAB_F32X2
						if (!stopped && clipi && slabOn && slabRejectPre2(S, e0[i], e1[i])) clipi = false;
#elif CMGPU_SLAB_HOIST
						if (!stopped && clipi && slabOn && slabRejectPre(S, e0[i], e1[i])) clipi = false;
#else
						if (!stopped && clipi && slabOn && slabReject(h, S, e0[i], e1[i])) clipi = false;
#endif
#else
						const bool clipi = want && !apart;                             // CM_BoundsIntersect passed
#endif
						if (!stopped) {
							adv = i + 1;
							b = bi;
							clip = clipi;
							if (COUNT) { cBrushRefs += bi >= 0 ? 1u : 0u; cBrushBounds += want ? 1u : 0u; }
							stopped = bi < 0 || clipi;
						}
					}
					SETI(PF_CUR, j, k + adv);
					if (stopped)
#if CMGPU_INLINE_SLAB
						modes = withMode(modes, j, b < 0 ? PM_POP : PM_BRUSH);
#else
						modes = withMode(modes, j, b < 0 ? PM_POP : ((FLDI(PF_FLAGS, j) & 4) ? PM_SLAB : PM_BRUSH));
#endif
				}
			}
		}
#if !CMGPU_INLINE_SLAB
		if (cSlab >= thr) {
			// ---------------- float slab pre-test of the entry that passed CM_BoundsIntersect (see slabReject) -----------------
			const int j = firstInMode<K>(modes, PM_SLAB);
			if (j >= 0) {
				const int k = FLDI(PF_CUR, j) - 1;
				const Pair4 ee = ldg32(&fm.stream[2 * k]);
				const float4 e0 = ee.a, e1 = ee.b;
				FastLane L;
				L.sx = FLD(PF_SX, j); L.sy = FLD(PF_SY, j); L.sz = FLD(PF_SZ, j);
				L.ex = FLD(PF_EX, j); L.ey = FLD(PF_EY, j); L.ez = FLD(PF_EZ, j);
				L.fraction = FLD(PF_FRACTION, j);
				modes = withMode(modes, j, slabReject(h, L, e0, e1) ? PM_LEAF : PM_BRUSH);
			}
		}
#endif
		if (cPop >= thr) {
			// ---------------- return to the pending far half of a split -----------------
			const int j = firstInMode<K>(modes, PM_POP);
			if (j >= 0) {
				const int m = popHalf(j);
				if (m != PM_POP) modes = withMode(modes, j, m);
			}
		}
		if (cSplit >= thr) {
			// ---------------- a node that is not a plain axial descent, cm_trace.c:456-537 -----------------
			const int j = firstInMode<K>(modes, PM_SPLIT);
			if (j >= 0) {
				const int num = FLDI(PF_CUR, j);
				const int4 nd = __ldg(&fm.nodes[num]);
				const float ax_ = FLD(PF_AX, j), ay_ = FLD(PF_AY, j), az_ = FLD(PF_AZ, j);
				const float bx_ = FLD(PF_BX, j), by_ = FLD(PF_BY, j), bz_ = FLD(PF_BZ, j);
				const float f1 = FLD(PF_F1, j), f2 = FLD(PF_F2, j);
				double t1, t2, off;
				if (nd.z < 3) {
					const float dist = __int_as_float(nd.w);
					t1 = (double)(sel3(ax_, ay_, az_, nd.z) - dist);      // float subtract, then widened
					t2 = (double)(sel3(bx_, by_, bz_, nd.z) - dist);
					off = h.isPoint ? 0.0 : (double)sel3(h.p0[0], h.p0[1], h.p0[2], nd.z);
				} else {
					const float4 pl = __ldg(&fm.nodePlanes[num]);
					t1 = dotd((double)pl.x, (double)pl.y, (double)pl.z, ax_, ay_, az_) - (double)pl.w;
					t2 = dotd((double)pl.x, (double)pl.y, (double)pl.z, bx_, by_, bz_) - (double)pl.w;
					off = h.isPoint ? 0.0 : 2048.0;
				}
				int c;
				if (t1 >= off + 1 && t2 >= off + 1) {
					c = nd.x;
				} else if (t1 < -off - 1 && t2 < -off - 1) {
					c = nd.y;
				} else {
					const bool fwd = t1 < t2, eq = !(t1 < t2) && !(t1 > t2);
					const float inv = (float)(1.0 / (eq ? 1.0 : (t1 - t2)));
					const double nfar = fwd ? (t1 + off + kClipEps) : (t1 - off - kClipEps);
					const double nnear = fwd ? (t1 - off + kClipEps) : (t1 + off + kClipEps);
					float ffar = (float)(nfar * (double)inv);
					float fnear = (float)(nnear * (double)inv);
					if (eq) { fnear = 1.0f; ffar = 0.0f; }
					if (fnear < 0) fnear = 0; else if (fnear > 1) fnear = 1;
					if (ffar < 0) ffar = 0; else if (ffar > 1) ffar = 1;
					if (COUNT) cSplits++;
					const int fl = FLDI(PF_FLAGS, j);
					const int sp = (fl >> 8) & 0xff;
					if (sp < q.stackDepth) {
						float4 *sp4 = STACKAT(j, sp);
						const unsigned long long keep = l2KeepPolicy();
						stKeep32B(sp4, make_float4(__int_as_float(fwd ? nd.x : nd.y), f1 + (f2 - f1) * ffar, f2, ax_ + ffar * (bx_ - ax_)),
						          make_float4(ay_ + ffar * (by_ - ay_), az_ + ffar * (bz_ - az_), bx_, by_), keep);
						stKeep32(reinterpret_cast<int *>(STACKZ(j, sp)), __float_as_int(bz_), keep);
						SETI(PF_FLAGS, j, fl + 256);
					} else {
						atomicOr(q.status, 1);
					}
					FLD(PF_F2, j) = f1 + (f2 - f1) * fnear;
					FLD(PF_BX, j) = ax_ + fnear * (bx_ - ax_);
					FLD(PF_BY, j) = ay_ + fnear * (by_ - ay_);
					FLD(PF_BZ, j) = az_ + fnear * (bz_ - az_);
					c = fwd ? nd.y : nd.x;
				}
				if (c < 0) { SETI(PF_CUR, j, -1 - c); modes = withMode(modes, j, PM_LEAF); if (COUNT) cLeafs++; }
				else { SETI(PF_CUR, j, c); modes = withMode(modes, j, PM_NODE); }
			}
		}
		if (cBrush >= thrBrush) {
			// ---------------- CM_TraceThroughBrush for the pending brush, cm_trace.c:261-363 -----
			const int j = firstInMode<K>(modes, PM_BRUSH);
			if (j >= 0) {
				FastLane L;
				L.sx = FLD(PF_SX, j); L.sy = FLD(PF_SY, j); L.sz = FLD(PF_SZ, j);
				L.ex = FLD(PF_EX, j); L.ey = FLD(PF_EY, j); L.ez = FLD(PF_EZ, j);
				L.fraction = FLD(PF_FRACTION, j);
				const int fl = FLDI(PF_FLAGS, j);
				L.flags = fl & 0xff;
				constexpr int kUnset = (int)0x80000000;           // the header is only written when the clip changed it
				L.planeRef = kUnset; L.hitId = kUnset;
				L.cBrushSides = L.cCross = 0;
				const int b = __float_as_int(__ldg(&fm.stream[2 * (FLDI(PF_CUR, j) - 1) + 1]).w);   // the entry that passed
				const bool over = clipBrushFast<COUNT>(fm, h, L, b);
				if (COUNT) { cBrushSides += L.cBrushSides; cCross += L.cCross; }
				FLD(PF_FRACTION, j) = L.fraction;
				SETI(PF_FLAGS, j, (fl & ~0xff) | L.flags);
				if (L.planeRef != kUnset) stKeep32(&HDR(j).planeRef, L.planeRef, l2KeepPolicy());
				if (L.hitId != kUnset) stKeep32(&HDR(j).hitId, L.hitId, l2KeepPolicy());
#if CMGPU_POOL_LAST2
				SETI(PF_LAST1, j, FLDI(PF_LAST0, j));
#endif
				SETI(PF_LAST0, j, b);
				modes = withMode(modes, j, over ? PM_DONE : PM_LEAF);
			}
		}
		if (cPos >= thr) {
			// ---------------- position test: CM_PositionTest / CM_BoxLeafnums_r / CM_TestInLeaf / CM_TestBoxInBrush,
			// cm_trace.c:83-226, cm_test.c:131-156.  The reference lists <= 1024 leafs and then tests them in list order;
			// testing each leaf when it is found is the same sequence of tests.
			#pragma unroll 1
			for (int rep = 0; rep < q.lightReps; rep++) {
				const int j = firstInMode<K>(modes, PM_POS);
				if (!__any_sync(0xffffffffu, j >= 0)) break;
				if (j >= 0 && !(FLDI(PF_FLAGS, j) & 8)) {
					const int c = FLDI(PF_CUR, j);
					if (c < 0) {
						const int found = FLDI(PF_FLAGS, j) >> 16;
						if (found >= kMaxPosLeafs) {
							modes = withMode(modes, j, PM_DONE);
						} else {
							SETI(PF_FLAGS, j, FLDI(PF_FLAGS, j) + (1 << 16));
							SETI(PF_CUR, j, -1 - c);
							if (COUNT) cLeafs++;
							SETI(PF_FLAGS, j, FLDI(PF_FLAGS, j) | 8);
						}
					} else {
						const int4 nd = __ldg(&fm.nodes[c]);
						if (COUNT) cNodes++;
						// start == end: tw.bounds = start + size; the listing box is one unit wider (cm_trace.c:199-205)
						const float lox = (FLD(PF_SX, j) + h.n0[0]) - 1.0f, loy = (FLD(PF_SY, j) + h.n0[1]) - 1.0f, loz = (FLD(PF_SZ, j) + h.n0[2]) - 1.0f;
						const float hix = (FLD(PF_SX, j) + h.p0[0]) + 1.0f, hiy = (FLD(PF_SY, j) + h.p0[1]) + 1.0f, hiz = (FLD(PF_SZ, j) + h.p0[2]) + 1.0f;
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
							SETI(PF_CUR, j, nd.x);
						} else if (s == 2) {
							SETI(PF_CUR, j, nd.y);
						} else {
							const int fl = FLDI(PF_FLAGS, j);
							const int sp = (fl >> 8) & 0xff;
							if (sp < q.stackDepth) { __stcg(reinterpret_cast<int *>(STACKAT(j, sp)), nd.y); SETI(PF_FLAGS, j, fl + 256); }
							else atomicOr(q.status, 1);
							SETI(PF_CUR, j, nd.x);
						}
					}
				} else if (j >= 0) {
					const int k = FLDI(PF_CUR, j);
					const Pair4 ee = ldg32(&fm.stream[2 * k]);
					const float4 e0 = ee.a, e1 = ee.b;
					SETI(PF_CUR, j, k + 1);
					const int b = __float_as_int(e1.w);
					if (b < 0) {
						const int fl = FLDI(PF_FLAGS, j);
						const int sp = (fl >> 8) & 0xff;
						if (sp == 0) {
							modes = withMode(modes, j, PM_DONE);
						} else {
							SETI(PF_CUR, j, __ldcg(reinterpret_cast<const int *>(STACKAT(j, sp - 1))));
							SETI(PF_FLAGS, j, (fl - 256) & ~8);
						}
					} else {
						if (COUNT) cBrushRefs++;
						const int contents = __float_as_int(e0.w);
						if (contents & h.mask) {
							const Pair4 rr = ldg32(&fm.brushRec[2 * b]);
							const float4 r0 = rr.a, r1 = rr.b;
							const int firstSide = __float_as_int(r0.w);
							const int ns = __float_as_int(r1.w) & 0xffffff;
							if (ns > 0) {
								if (COUNT) cBrushBounds++;
								const float sx = FLD(PF_SX, j), sy = FLD(PF_SY, j), sz = FLD(PF_SZ, j);
								const float lox = sx + h.n0[0], loy = sy + h.n0[1], loz = sz + h.n0[2];
								const float hix = sx + h.p0[0], hiy = sy + h.p0[1], hiz = sz + h.p0[2];
								if (!(lox > r1.x || loy > r1.y || loz > r1.z || hix < r0.x || hiy < r0.y || hiz < r0.z)) {
									bool inside = true;
									for (int s = 6; s < ns; s++) {
										const float4 pl = __ldg(&fm.sidePlanes[firstSide + s]);
										if (COUNT) cBrushSides++;
										const float cx = pl.x < 0.0f ? h.p0[0] : h.n0[0];
										const float cy = pl.y < 0.0f ? h.p0[1] : h.n0[1];
										const float cz = pl.z < 0.0f ? h.p0[2] : h.n0[2];
										const float distf = pl.w - dotf(cx, cy, cz, pl.x, pl.y, pl.z);    // all-float (cm_trace.c:111)
										const double d1 = dotd((double)sx, (double)sy, (double)sz, pl.x, pl.y, pl.z) - (double)distf;
										if (d1 > 0) { inside = false; break; }
									}
									if (inside) {
										SETI(PF_FLAGS, j, FLDI(PF_FLAGS, j) | 3);
										FLD(PF_FRACTION, j) = 0.0f;
										stKeep32(&HDR(j).hitId, b, l2KeepPolicy());
										modes = withMode(modes, j, PM_DONE);
									}
								}
							}
						}
					}
				}
			}
		}
	}
	if (COUNT) {
		Counters *c = q.counters;
		atomicAdd(&c->traces, (unsigned long long)cTraces);
		atomicAdd(&c->nodes, (unsigned long long)cNodes);
		atomicAdd(&c->leafs, (unsigned long long)cLeafs);
		atomicAdd(&c->brushRefs, (unsigned long long)cBrushRefs);
		atomicAdd(&c->brushBounds, (unsigned long long)cBrushBounds);
		atomicAdd(&c->brushSides, (unsigned long long)cBrushSides);
		atomicAdd(&c->crossings, (unsigned long long)cCross);
		atomicAdd(&c->splits, (unsigned long long)cSplits);
	}
#undef FLD
#undef FLDI
#undef SETI
#undef STACKAT
#undef STACKZ
#undef HDR
}

}  // namespace cmgpu
