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
	const int4   *entInfo;     // number, clip handle (inline model or 1023), rotation index (-1: angles are zero), 0
	const float  *entOrigin;   // [e][3]
	const float  *entMins;     // [e][3]  r.mins / r.maxs: the CM_TempBoxModel box
	const float  *entMaxs;
	const float  *entRot;      // [e][9]
	const int    *ownerByNum;  // r.ownerNum of every gentity
	int numNodes, numEnts, numGentities;
};

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
};

struct SvItems {                         // the entity pass's batch, one item per candidate
	float *starts, *ends, *mins, *maxs;  // mins/maxs only written when the moves have their own hulls
	int *models, *masks;                 // masks only written when the moves have their own masks
	float *origins;
	int *rotIndex;
	float *boxmins, *boxmaxs;
	int *entNumber;                      // touch->s.number of the candidate
};

// the walk of SV_AreaEntities_r + the filters of SV_ClipMoveToEntities; f(e) is called per survivor, in list order
template <typename F>
__device__ __forceinline__ void svForEachCandidate(const SvWorld &w, float lox, float loy, float loz, float hix, float hiy, float hiz,
                                                   int passEnt, int mask, bool wantFilters, F f) {
	int passOwner = -1;
	if (wantFilters && passEnt != kEntityNumNone) {
		passOwner = (passEnt >= 0 && passEnt < w.numGentities) ? __ldg(&w.ownerByNum[passEnt]) : kEntityNumNone;
		if (passOwner == kEntityNumNone) passOwner = -1;
	}
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
			const float4 lo = __ldg(&w.entLo[e]);
			const float4 hi = __ldg(&w.entHi[e]);
			if (lo.x > hix || lo.y > hiy || lo.z > hiz || hi.x < lox || hi.y < loy || hi.z < loz) continue;   // sv_world.c:349-356
			if (wantFilters) {
				const int number = __ldg(&w.entInfo[e]).x;
				const int owner = __float_as_int(lo.w);
				if (passEnt != kEntityNumNone) {                       // sv_world.c:500-510
					if (number == passEnt) continue;
					if (owner == passEnt) continue;
					if (owner == passOwner) continue;
				}
				if (!(mask & __float_as_int(hi.w))) continue;          // sv_world.c:514-516
			}
			f(e);
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
		svForEachCandidate(w, lo[0], lo[1], lo[2], hi[0], hi[1], hi[2], passEnt, mask, true, [&](int) { count++; });
		q.offsets[m] = count;
	} else {
		size_t k = q.offsets[m];
		if (k == q.offsets[m + 1]) return;
		svForEachCandidate(w, lo[0], lo[1], lo[2], hi[0], hi[1], hi[2], passEnt, mask, true, [&](int e) {
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
			it.entNumber[k] = info.x;
			k++;
		});
	}
}

// SV_ClipMoveToEntities' loop body after the clip (sv_world.c:490-549), over the candidates of one move
__global__ void __launch_bounds__(128) svMergeKernel(int n, uint2 *results, const unsigned *offsets, const uint2 *cand, const int *entNumber) {
	const int m = blockIdx.x * blockDim.x + threadIdx.x;
	if (m >= n) return;
	const unsigned k0 = offsets[m], k1 = offsets[m + 1];
	if (k0 == k1) return;
	uint2 *dst = results + 7 * (size_t)m;
	uint2 clip[7];
	#pragma unroll
	for (int j = 0; j < 7; j++) clip[j] = dst[j];
	// words: 0 allsolid, 1 startsolid, 2 fraction, ..., 13 entityNum  ->  clip[0] = (allsolid, startsolid), clip[1].x = fraction, clip[6].y = entityNum
	for (unsigned k = k0; k < k1; k++) {
		if (clip[0].x) break;                                          // :491-493
		const uint2 *src = cand + 7 * (size_t)k;
		uint2 t[7];
		#pragma unroll
		for (int j = 0; j < 7; j++) t[j] = src[j];
		const unsigned num = (unsigned)entNumber[k];
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
