// cm_sv_world.cuh -- SV_Trace / SV_PointContents as device work (server/sv_world.c:330-647).
//
// The reference clips a move against the world, then asks its sector tree for the linked entities whose absolute
// box touches the box of the move (SV_AreaEntities), drops the pass entity, its missiles and the entities whose
// contents the mask does not want (sv_world.c:486-520), clips the move against each survivor with
// CM_TransformedBoxTrace and merges the results in list order (sv_world.c:532-549).  Here
//
//   world pass    the batch goes through the trace kernels as it is (model 0)
//   svGather<0>   one thread per move: entityNum of the world result, the box of the move, a walk over the sector
//                 tree that COUNTS the surviving candidates
//   scan          exclusive prefix sum of the counts (the binning scan kernels)
//   svGather<1>   the same walk, now WRITING one transformed-trace item per candidate
//   entity pass   all candidates of all moves through the trace kernels in one launch
//   svMerge       one thread per move: the reference's merge rules over its candidates, in list order
//
// Order is what makes the result bit-identical: SV_AreaEntities_r lists a node's entities, then child 0, then
// child 1, so every query sees a subsequence of ONE global order (the order a query with an infinite box returns).
// The host stores the entities in that order, node by node, and the nodes in depth-first order; a node is entered
// if its parent was and the parent's plane test passes (sv_world.c:370-375), else its whole subtree is skipped.
// An entity in a subtree the query does not enter fails the box test anyway (it was linked there because its box
// lies beyond that plane), so the candidates are simply "the entities whose box touches, in global order" -- which
// lets the common case skip the tree: two uniform grids hold, per cell, the ascending indices of the entities that
// reach the cell, and a move merges the lists of the few cells its box covers.
#pragma once

#include "cm_kernels.cuh"

namespace cmgpu {

constexpr int kEntityNumNone = 4095;     // ENTITYNUM_NONE = MAX_GENTITIES - 1, GENTITYNUM_BITS 12 (q_shared.h:1054-1061)
constexpr int kEntityNumWorld = 4094;    // ENTITYNUM_WORLD

struct SvWorld {
	const int4   *nodes;       // [2i]   parent axis, float bits of parent dist, side (0: child 0, 1: child 1, -1: root), end of subtree
	                           // [2i+1] first entity, end entity, 0, 0
	const float4 *entLo;       // absmin.xyz, ownerNum bits
	const float4 *entHi;       // absmax.xyz, contents bits
	const int4   *entInfo;     // number, clip handle (inline model or 1023), rotation index (-1: angles are zero),
	                           // first cell of the entity's box: x | y << 8 at the fine level, the same << 16 at the coarse level
	const float  *entOrigin;   // [e][3]
	const float  *entMins;     // [e][3]  r.mins / r.maxs: the CM_TempBoxModel box
	const float  *entMaxs;
	const float  *entRot;      // [e][9]
	const int    *ownerByNum;  // r.ownerNum of every gentity
	int numNodes, numEnts, numGentities;
	// two uniform x/y grids over the world (fine, coarse): per cell the entities whose box reaches it, ascending
	const int    *cellStart[2];
	const int    *cellEnts[2];
	float gridLo[2];
	float gridInv[2][2];
	int gridCells[2];          // cells per axis
};
constexpr int kSvMaxCells = 16;

struct SvMoves {
	int n;
	const float *starts, *ends;
	const float *mins, *maxs;            // per move, or null: shared
	float sharedMins[3], sharedMaxs[3];
	const int *passEnts;                 // per move, or null: shared
	int sharedPass;
	const int *masks;                    // per move, or null: shared
	int sharedMask;
	uint2 *results;                      // world traces in, merged traces out
	unsigned *offsets;                   // [n+1]: counts, then (after the scan) first candidate of every move
	int4 *firstCands;                    // [n]: the first four candidates the counting pass found (saves the second walk)
};

struct SvItems {                         // the entity pass's batch, one item per candidate
	float *starts, *ends, *mins, *maxs;  // mins/maxs only written when the moves have their own hulls
	int *models, *masks;                 // masks only written when the moves have their own masks
	float *origins;
	int *rotIndex;
	float *boxmins, *boxmaxs;
	int *entIndex;                       // the candidate's place in SV_AreaEntities order (its index in the snapshot)
};

// pass entity, its missiles, its owner's other missiles, unwanted contents (sv_world.c:496-516)
struct SvFilter {
	int passEnt, passOwner, mask;
	bool on;
};

__device__ __forceinline__ SvFilter svMakeFilter(const SvWorld &w, int passEnt, int mask, bool on) {
	SvFilter f{passEnt, -1, mask, on};
	if (on && passEnt != kEntityNumNone) {                                     // :484-493
		f.passOwner = (passEnt >= 0 && passEnt < w.numGentities) ? __ldg(&w.ownerByNum[passEnt]) : kEntityNumNone;
		if (f.passOwner == kEntityNumNone) f.passOwner = -1;
	}
	return f;
}

__device__ __forceinline__ bool svTouches(const SvWorld &w, int e, float lox, float loy, float loz, float hix, float hiy, float hiz,
                                          const SvFilter &f) {
	const float4 lo = __ldg(&w.entLo[e]);
	const float4 hi = __ldg(&w.entHi[e]);
	if (lo.x > hix || lo.y > hiy || lo.z > hiz || hi.x < lox || hi.y < loy || hi.z < loz) return false;   // sv_world.c:349-356
	if (f.on) {
		const int owner = __float_as_int(lo.w);
		if (f.passEnt != kEntityNumNone) {                                     // :500-510
			if (__ldg(&w.entInfo[e]).x == f.passEnt) return false;
			if (owner == f.passEnt) return false;
			if (owner == f.passOwner) return false;
		}
		if (!(f.mask & __float_as_int(hi.w))) return false;                    // :514-516
	}
	return true;
}

// cell of a coordinate on one axis of a grid level: monotone in x, and the SAME float ops as the host uses for the
// entities' cell ranges (svCellOf in cm_gpu.cu), so overlapping boxes always share a cell
__host__ __device__ __forceinline__ int svCellOf(float x, float lo, float inv, int cells) {
	const float t = (x - lo) * inv;
	if (!(t > 0.0f)) return 0;
	if (t >= (float)cells) return cells - 1;
	return (int)t;
}

// The candidates of one move, each once, in no particular order: the entities listed in the cells that the move's
// box covers, at the finer of the two levels where those are at most kSvMaxCells cells; a larger box walks the
// sector tree.  Entity indices ARE the order SV_AreaEntities lists them in, so the merge kernel can restore it.
template <typename F>
__device__ __forceinline__ void svForEachCandidate(const SvWorld &w, float lox, float loy, float loz, float hix, float hiy, float hiz,
                                                   int passEnt, int mask, bool wantFilters, F f) {
	const SvFilter flt = svMakeFilter(w, passEnt, mask, wantFilters);
	#pragma unroll 1
	for (int level = 0; level < 2; level++) {
		const int G = w.gridCells[level];
		const int cx0 = svCellOf(lox, w.gridLo[0], w.gridInv[level][0], G), cx1 = svCellOf(hix, w.gridLo[0], w.gridInv[level][0], G);
		const int cy0 = svCellOf(loy, w.gridLo[1], w.gridInv[level][1], G), cy1 = svCellOf(hiy, w.gridLo[1], w.gridInv[level][1], G);
		const int nx = cx1 - cx0 + 1, nc = nx * (cy1 - cy0 + 1);
		if (nc > kSvMaxCells) continue;
		const int *cellStart = w.cellStart[level], *cellEnts = w.cellEnts[level];
		for (int cy = cy0; cy <= cy1; cy++) {
			for (int cx = cx0; cx <= cx1; cx++) {
				const int cell = cy * G + cx;
				const int p1 = __ldg(&cellStart[cell + 1]);
				for (int p = __ldg(&cellStart[cell]); p < p1; p++) {
					const int e = __ldg(&cellEnts[p]);
					// an entity that reaches several of the covered cells is taken in the first of them
					const int first = __ldg(&w.entInfo[e]).w >> (16 * level);
					if (max(first & 0xff, cx0) != cx || max((first >> 8) & 0xff, cy0) != cy) continue;
					if (svTouches(w, e, lox, loy, loz, hix, hiy, hiz, flt)) f(e);
				}
			}
		}
		return;
	}
	// the walk of SV_AreaEntities_r (sv_world.c:344-377) over the depth-first node array
	int i = 0;
	while (i < w.numNodes) {
		const int4 a = __ldg(&w.nodes[2 * i]);
		if (a.z >= 0) {
			const float dist = __int_as_float(a.y);
			const float v = a.z == 0 ? sel3(hix, hiy, hiz, a.x) : sel3(lox, loy, loz, a.x);
			const bool enter = a.z == 0 ? (v > dist) : (v < dist);         // sv_world.c:370-375
			if (!enter) { i = a.w; continue; }
		}
		const int4 b = __ldg(&w.nodes[2 * i + 1]);
		for (int e = b.x; e < b.y; e++) {
			if (svTouches(w, e, lox, loy, loz, hix, hiy, hiz, flt)) f(e);
		}
		i++;
	}
}

template <bool FILL>
__global__ void __launch_bounds__(128) svGatherKernel(const SvWorld w, const SvMoves q, const SvItems it) {
	const int m = blockIdx.x * blockDim.x + threadIdx.x;
	if (m >= q.n) return;
	int *rec = reinterpret_cast<int *>(q.results) + 14 * (size_t)m;
	const float fraction = __int_as_float(rec[2]);
	if (!FILL) rec[13] = fraction != 1.0f ? kEntityNumWorld : kEntityNumNone;      // sv_world.c:577
	if (fraction == 0.0f) {                                                         // blocked at once by the world (:578-581)
		if (!FILL) q.offsets[m] = 0;
		return;
	}
	const size_t o = 3 * (size_t)m;
	const float s[3] = {q.starts[o], q.starts[o + 1], q.starts[o + 2]};
	const float e3[3] = {q.ends[o], q.ends[o + 1], q.ends[o + 2]};
	float mn[3], mx[3];
	#pragma unroll
	for (int a = 0; a < 3; a++) {
		mn[a] = q.mins ? q.mins[o + a] : q.sharedMins[a];
		mx[a] = q.maxs ? q.maxs[o + a] : q.sharedMaxs[a];
	}
	// the box of the whole move, one unit wider (sv_world.c:596-604): float add, then float subtract of 1
	float lo[3], hi[3];
	#pragma unroll
	for (int a = 0; a < 3; a++) {
		if (e3[a] > s[a]) { lo[a] = (s[a] + mn[a]) - 1.0f; hi[a] = (e3[a] + mx[a]) + 1.0f; }
		else { lo[a] = (e3[a] + mn[a]) - 1.0f; hi[a] = (s[a] + mx[a]) + 1.0f; }
	}
	const int passEnt = q.passEnts ? q.passEnts[m] : q.sharedPass;
	const int mask = q.masks ? q.masks[m] : q.sharedMask;
	if (!FILL) {
		unsigned count = 0;
		int4 c4 = make_int4(0, 0, 0, 0);
		svForEachCandidate(w, lo[0], lo[1], lo[2], hi[0], hi[1], hi[2], passEnt, mask, true, [&](int e) {
			if (count == 0) c4.x = e; else if (count == 1) c4.y = e; else if (count == 2) c4.z = e; else if (count == 3) c4.w = e;
			count++;
		});
		q.offsets[m] = count;
		if (count) q.firstCands[m] = c4;
	} else {
		size_t k = q.offsets[m];
		const unsigned count = q.offsets[m + 1] - (unsigned)k;
		if (count == 0) return;
		auto emit = [&](int e) {
			const size_t k3 = 3 * k, e3o = 3 * (size_t)e;
			const int4 info = __ldg(&w.entInfo[e]);
			#pragma unroll
			for (int a = 0; a < 3; a++) {
				it.starts[k3 + a] = s[a];
				it.ends[k3 + a] = e3[a];
				it.origins[k3 + a] = __ldg(&w.entOrigin[e3o + a]);
				it.boxmins[k3 + a] = __ldg(&w.entMins[e3o + a]);
				it.boxmaxs[k3 + a] = __ldg(&w.entMaxs[e3o + a]);
			}
			if (q.mins) {
				#pragma unroll
				for (int a = 0; a < 3; a++) { it.mins[k3 + a] = mn[a]; it.maxs[k3 + a] = mx[a]; }
			}
			if (q.masks) it.masks[k] = mask;
			it.models[k] = info.y;
			it.rotIndex[k] = info.z;
			it.entIndex[k] = e;
			k++;
		};
		if (count <= 4) {
			const int4 c4 = q.firstCands[m];
			emit(c4.x);
			if (count > 1) emit(c4.y);
			if (count > 2) emit(c4.z);
			if (count > 3) emit(c4.w);
		} else {
			svForEachCandidate(w, lo[0], lo[1], lo[2], hi[0], hi[1], hi[2], passEnt, mask, true, emit);
		}
	}
}

// SV_ClipMoveToEntities' loop body after the clip (sv_world.c:490-549), over the candidates of one move IN LIST
// ORDER: the gather emits them in cell order, so each round takes the candidate with the smallest entity index
// that has not been taken yet (a move has a handful of candidates)
__global__ void __launch_bounds__(128) svMergeKernel(int n, uint2 *results, const unsigned *offsets, const uint2 *cand, const int *entIndex,
                                                     const int4 *entInfo) {
	const int m = blockIdx.x * blockDim.x + threadIdx.x;
	if (m >= n) return;
	const unsigned k0 = offsets[m], k1 = offsets[m + 1];
	if (k0 == k1) return;
	uint2 *dst = results + 7 * (size_t)m;
	uint2 clip[7];
	#pragma unroll
	for (int j = 0; j < 7; j++) clip[j] = dst[j];
	// words: 0 allsolid, 1 startsolid, 2 fraction, ..., 13 entityNum  ->  clip[0] = (allsolid, startsolid), clip[1].x = fraction, clip[6].y = entityNum
	int last = -1;
	for (unsigned round = k0; round < k1; round++) {
		if (clip[0].x) break;                                          // :491-493
		unsigned k = k0;
		int e = 0x7fffffff;
		for (unsigned j = k0; j < k1; j++) {
			const int ej = entIndex[j];
			if (ej > last && ej < e) { e = ej; k = j; }
		}
		last = e;
		const uint2 *src = cand + 7 * (size_t)k;
		uint2 t[7];
		#pragma unroll
		for (int j = 0; j < 7; j++) t[j] = src[j];
		const unsigned num = (unsigned)__ldg(&entInfo[e]).x;           // touch->s.number
		if (t[0].x) { clip[0].x = 1; t[6].y = num; }                   // :532-534
		else if (t[0].y) { clip[0].y = 1; t[6].y = num; }              // :535-538
		if (__uint_as_float(t[1].x) < __uint_as_float(clip[1].x)) {    // :540-549
			const unsigned oldStart = clip[0].y;
			t[6].y = num;
			#pragma unroll
			for (int j = 0; j < 7; j++) clip[j] = t[j];
			clip[0].y |= oldStart;
		}
	}
	#pragma unroll
	for (int j = 0; j < 7; j++) dst[j] = clip[j];
}

// SV_PointContents (sv_world.c:616-647): pass 0 counts the entities whose box holds the point (minus the pass
// entity), pass 1 writes one CM_TransformedPointContents item per candidate
template <bool FILL>
__global__ void __launch_bounds__(128) svPointGatherKernel(const SvWorld w, int n, const float *points, const int *passEnts, int sharedPass,
                                                           unsigned *offsets, float *itPoints, int *itModels, float *itOrigins,
                                                           int *itRotIndex, float *itBoxMins, float *itBoxMaxs) {
	const int m = blockIdx.x * blockDim.x + threadIdx.x;
	if (m >= n) return;
	const size_t o = 3 * (size_t)m;
	const float px = points[o], py = points[o + 1], pz = points[o + 2];
	const int passEnt = passEnts ? passEnts[m] : sharedPass;
	if (!FILL) {
		unsigned count = 0;
		svForEachCandidate(w, px, py, pz, px, py, pz, kEntityNumNone, 0, false, [&](int e) {
			if (__ldg(&w.entInfo[e]).x != passEnt) count++;            // :632-634
		});
		offsets[m] = count;
	} else {
		size_t k = offsets[m];
		if (k == offsets[m + 1]) return;
		svForEachCandidate(w, px, py, pz, px, py, pz, kEntityNumNone, 0, false, [&](int e) {
			const int4 info = __ldg(&w.entInfo[e]);
			if (info.x == passEnt) return;
			const size_t k3 = 3 * k, e3o = 3 * (size_t)e;
			itPoints[k3] = px; itPoints[k3 + 1] = py; itPoints[k3 + 2] = pz;
			#pragma unroll
			for (int a = 0; a < 3; a++) {
				itOrigins[k3 + a] = __ldg(&w.entOrigin[e3o + a]);
				itBoxMins[k3 + a] = __ldg(&w.entMins[e3o + a]);
				itBoxMaxs[k3 + a] = __ldg(&w.entMaxs[e3o + a]);
			}
			itModels[k] = info.y;
			itRotIndex[k] = info.z;
			k++;
		});
	}
}

// contents[m] |= every candidate's (sv_world.c:643)
__global__ void __launch_bounds__(128) svPointMergeKernel(int n, int *contents, const unsigned *offsets, const int *cand) {
	const int m = blockIdx.x * blockDim.x + threadIdx.x;
	if (m >= n) return;
	int c = contents[m];
	for (unsigned k = offsets[m]; k < offsets[m + 1]; k++) c |= cand[k];
	contents[m] = c;
}

}  // namespace cmgpu
