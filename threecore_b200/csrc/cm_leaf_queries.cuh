// cm_leaf_queries.cuh -- CM_PointLeafnum and CM_BoxLeafnums as batches (cm_test.c:31-61,73-88,131-181): the tree
// queries behind entity linking (SV_LinkEntity's cluster / area lists, sv_world.c:266-303) and snapshot culling.
// One thread per query, the reference's own loops; the recursion of CM_BoxLeafnums_r becomes a stack of the
// children[1] still to visit, which keeps the reference's order (children[0] first).
#pragma once

#include "cm_kernels.cuh"

namespace cmgpu {

__global__ void __launch_bounds__(128) pointLeafnumKernel(const DevMap m, int n, const float *points, int *out) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float px = points[3 * (size_t)i], py = points[3 * (size_t)i + 1], pz = points[3 * (size_t)i + 2];
	int num = 0;
	while (num >= 0) {                                           // CM_PointLeafnum_r, cm_test.c:31-54
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
	out[i] = -1 - num;
}

struct LeafListBatch {
	int n;
	const float *mins, *maxs;            // [n][3]
	int listsize;
	int *lists;                          // [n][listsize]; entries past counts[i] are left untouched
	int *counts;                         // leafs stored (<= listsize)
	int *lastLeafs;                      // last leaf with a cluster, counted even when the list is full (cm_test.c:78-81)
	const int *leafCluster;              // [numLeafs][2] cluster, area; null: no leaf has a cluster
	int *status;
};

template <int DEPTH>
__global__ void __launch_bounds__(128) boxLeafnumsKernel(const DevMap m, const LeafListBatch q) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= q.n) return;
	const size_t o = 3 * (size_t)i;
	const float lox = q.mins[o], loy = q.mins[o + 1], loz = q.mins[o + 2];
	const float hix = q.maxs[o], hiy = q.maxs[o + 1], hiz = q.maxs[o + 2];
	int *list = q.lists + (size_t)i * q.listsize;
	int stack[DEPTH];
	int sp = 0, count = 0, lastLeaf = 0, num = 0;
	for (;;) {
		while (num >= 0) {                                       // CM_BoxLeafnums_r, cm_test.c:131-156
			const int4 nd = __ldg(&m.nodes[num]);
			const int s = boxPlaneSide(m, num, nd, lox, loy, loz, hix, hiy, hiz);
			if (s == 1) num = nd.x;
			else if (s == 2) num = nd.y;
			else {
				if (sp < DEPTH) stack[sp++] = nd.y; else atomicOr(q.status, 1);
				num = nd.x;
			}
		}
		const int leaf = -1 - num;                               // CM_StoreLeafs, cm_test.c:73-88
		if (q.leafCluster && __ldg(&q.leafCluster[2 * leaf]) != -1) lastLeaf = leaf;
		if (count < q.listsize) list[count++] = leaf;
		if (sp == 0) break;
		num = stack[--sp];
	}
	q.counts[i] = count;
	q.lastLeafs[i] = lastLeaf;
}

// ---- entity linking, PVS part (server/sv_world.c:255-303), one walk per entity -----------------------------
// SV_LinkEntity lists up to 128 leafs with CM_BoxLeafnums and then loops over the list twice (areas, clusters).
// Both loops only look at one leaf at a time, so they run INSIDE the walk, on the leafs that would have been
// stored; the leaf list itself is never materialised.  lastLeaf keeps counting past the 128th (cm_test.c:78-81).
struct EntClusters {                     // == cmgpu_entity_clusters_t
	int numLeafs, areanum, areanum2, numClusters, lastCluster;
	int clusternums[16];
};
constexpr int kMaxEntLeafs = 128;        // MAX_TOTAL_ENT_LEAFS, sv_world.c:178
constexpr int kMaxEntClusters = 16;      // MAX_ENT_CLUSTERS, server.h:35

struct EntClusterBatch {
	int n;
	const float *absmins, *absmaxs;      // [n][3]
	EntClusters *out;
	const int *leafCluster;              // [numLeafs][2] cluster, area; null: every leaf is (-1, -1)
	int *status;
};

template <int DEPTH>
__global__ void __launch_bounds__(128) entityClustersKernel(const DevMap m, const EntClusterBatch q) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= q.n) return;
	const size_t o = 3 * (size_t)i;
	const float lox = q.absmins[o], loy = q.absmins[o + 1], loz = q.absmins[o + 2];
	const float hix = q.absmaxs[o], hiy = q.absmaxs[o + 1], hiz = q.absmaxs[o + 2];
	int stack[DEPTH];
	int sp = 0, count = 0, lastLeaf = 0, num = 0;
	int areanum = -1, areanum2 = -1, numClusters = 0;
	bool full = false;                                       // the cluster loop hit its break (sv_world.c:294-296)
	int *clusters = q.out[i].clusternums;
	for (;;) {
		while (num >= 0) {                                   // CM_BoxLeafnums_r, cm_test.c:131-156
			const int4 nd = __ldg(&m.nodes[num]);
			const int s = boxPlaneSide(m, num, nd, lox, loy, loz, hix, hiy, hiz);
			if (s == 1) num = nd.x;
			else if (s == 2) num = nd.y;
			else {
				if (sp < DEPTH) stack[sp++] = nd.y; else atomicOr(q.status, 1);
				num = nd.x;
			}
		}
		const int leaf = -1 - num;
		int2 ca = make_int2(-1, -1);
		if (q.leafCluster) ca = __ldg(reinterpret_cast<const int2 *>(q.leafCluster) + leaf);
		if (ca.x != -1) lastLeaf = leaf;                     // CM_StoreLeafs, cm_test.c:78-81
		if (count < kMaxEntLeafs) {
			count++;
			if (ca.y != -1) {                                // sv_world.c:272-286
				if (areanum != -1 && areanum != ca.y) areanum2 = ca.y;
				else areanum = ca.y;
			}
			if (!full && ca.x != -1) {                       // :289-297
				clusters[numClusters++] = ca.x;
				if (numClusters == kMaxEntClusters) full = true;
			}
		}
		if (sp == 0) break;
		num = stack[--sp];
	}
	for (int k = numClusters; k < kMaxEntClusters; k++) clusters[k] = 0;
	int lastCluster = 0;
	if (full) lastCluster = q.leafCluster ? __ldg(&q.leafCluster[2 * lastLeaf]) : -1;     // :300-303
	q.out[i].numLeafs = count;
	q.out[i].areanum = areanum;
	q.out[i].areanum2 = areanum2;
	q.out[i].numClusters = numClusters;
	q.out[i].lastCluster = lastCluster;
}

// ---- snapshot culling, PVS stage (server/sv_snapshot.c:316-355) ---------------------------------------------
// One thread per (viewpoint, entity) pair, a warp's 32 verdicts leave as one word.  The viewpoint's leaf comes
// from pointLeafnumKernel; area connection is the flood number table the host keeps (cm_test.c:404-424).
struct VisBatch {
	int nViews, nEnts;
	const int *viewLeafs;                // [nViews]
	const EntClusters *ents;             // [nEnts]
	unsigned *visible;                   // [nViews][(nEnts+31)/32]
	const int *leafCluster;              // [numLeafs][2]
	const int *floodnum;                 // [numAreas]; null: cm_noAreas, everything connected
	int numAreas;
	const unsigned char *vis;            // [numClusters][clusterBytes]; null: not vised, every bit set
	int numClusters, clusterBytes;
};

__device__ __forceinline__ bool areasConnected(const VisBatch &q, int a1, int a2) {       // cm_test.c:404-424
	if (!q.floodnum) return true;
	if (a1 < 0 || a2 < 0 || a1 >= q.numAreas || a2 >= q.numAreas) return false;
	return __ldg(&q.floodnum[a1]) == __ldg(&q.floodnum[a2]);
}

__global__ void __launch_bounds__(128) visibleFromPointsKernel(const VisBatch q) {
	const int v = blockIdx.y;
	const int e = blockIdx.x * blockDim.x + threadIdx.x;
	const int words = (q.nEnts + 31) >> 5;
	bool visible = false;
	if (e < q.nEnts) {
		const int leaf = __ldg(&q.viewLeafs[v]);
		int2 ca = make_int2(-1, -1);
		if (q.leafCluster) ca = __ldg(reinterpret_cast<const int2 *>(q.leafCluster) + leaf);
		const int *ent = reinterpret_cast<const int *>(q.ents + e);
		const int areanum = __ldg(ent + 1), areanum2 = __ldg(ent + 2), n = __ldg(ent + 3), lastCluster = __ldg(ent + 4);
		if ((areasConnected(q, ca.y, areanum) || areasConnected(q, ca.y, areanum2)) && n) {
			// CM_ClusterPVS, cm_test.c:306-312: a cluster out of range reads row 0
			const unsigned char *row = nullptr;
			if (q.vis) row = q.vis + ((ca.x < 0 || ca.x >= q.numClusters) ? (size_t)0 : (size_t)ca.x * q.clusterBytes);
			// (the records come from the caller: a cluster number beyond the row reads as "not set" instead of reading past it)
			const unsigned rowBits = 8u * (unsigned)q.clusterBytes;
			const int nn = n < kMaxEntClusters ? n : kMaxEntClusters;
			int k, l = 0;
			for (k = 0; k < nn; k++) {
				l = __ldg(ent + 5 + k);
				if (!row || ((unsigned)l < rowBits && (__ldg(&row[l >> 3]) & (1 << (l & 7))))) break;
			}
			visible = true;
			if (k == nn) {
				if (lastCluster) {                           // the overflow range, sv_snapshot.c:347-353 as written
					const int lastScan = (unsigned)lastCluster < rowBits ? lastCluster : (int)rowBits;   // nothing is set beyond the row
					for (; l <= lastScan; l++) {
						if (!row || ((unsigned)l < rowBits && (__ldg(&row[l >> 3]) & (1 << (l & 7))))) break;
					}
					if (l == lastCluster) visible = false;
				} else {
					visible = false;
				}
			}
		}
	}
	const unsigned word = __ballot_sync(0xffffffffu, visible);
	if ((threadIdx.x & 31) == 0 && e < q.nEnts) q.visible[(size_t)v * words + (e >> 5)] = word;
}

}  // namespace cmgpu
