// cm_aas.cuh -- botlib's AAS tracer, batched (first half).
//
// AAS_TraceClientBBox (code/botlib/be_aas_sample.c:399-665) is what a bot's movement prediction asks hundreds of times
// per think: does the bot's bounding box, moved from start to end, stay inside areas it can be in?  The AAS file has
// the answer precomputed into its own BSP -- brushes already expanded by the standing and the crouching box -- so the
// query is a LINE through a tree of planes whose leaves are areas (child < 0), each with the presence types that fit,
// or solid (child 0).  One thread per line, the reference's explicit stack of 127 entries in local memory, float
// arithmetic with the two double-typed literals of the reference where they are (:624-625, :459/518).  Entity
// collisions (passent >= 0, :472-487) are server traces, i.e. cmgpu_sv_trace_batch; this is passent = -1.
// AAS_PointAreaNum (:212-258) is the same tree walked by a point.
#pragma once

#include "cm_kernels.cuh"

namespace cmgpu {

struct AasMap {
	const float4 *nodePlane;    // per node: normal.xyz, dist of its plane
	const int4   *nodeInfo;     // per node: planenum, children[0], children[1], 0   (node 0 is the reference's dummy)
	const float4 *planes;       // per plane: normal.xyz, dist (for the "plane faces the start" flip, :468)
	const int    *presence;     // per area: areasettings[].presencetype
	int numNodes, numPlanes, numAreas;
};

struct AasTraceBatch {
	int n;
	const float *starts, *ends;          // [n][3]
	const int *presencetypes;            // per item, or null: sharedPresence
	int sharedPresence;
	int *out;                            // aas_trace_t records, 9 ints each (be_aas.h:84-93)
};

constexpr int kAasStack = 127;           // tracestack[127], be_aas_sample.c:405
struct AasSeg { float sx, sy, sz, ex, ey, ez; int planenum, nodenum; };

__global__ void __launch_bounds__(128) aasTraceClientBBoxKernel(const AasMap m, const AasTraceBatch q) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= q.n) return;
	AasSeg stack[kAasStack];
	const size_t o = 3 * (size_t)i;
	const float s0x = q.starts[o], s0y = q.starts[o + 1], s0z = q.starts[o + 2];
	const float e0x = q.ends[o], e0y = q.ends[o + 1], e0z = q.ends[o + 2];
	const int presencetype = q.presencetypes ? q.presencetypes[i] : q.sharedPresence;
	int startsolid = 0, ent = 0, lastarea = 0, area = 0, planenum = 0;
	float fraction = 0.0f, px = 0.0f, py = 0.0f, pz = 0.0f;      // Com_Memset(&trace, 0, ...), :412
	int sp = 0;
	stack[0].sx = s0x; stack[0].sy = s0y; stack[0].sz = s0z;
	stack[0].ex = e0x; stack[0].ey = e0y; stack[0].ez = e0z;
	stack[0].planenum = 0;
	stack[0].nodenum = 1;
	sp = 1;
	for (;;) {
		sp--;
		if (sp < 0) {                                            // nothing was hit, :421-434
			startsolid = 0; fraction = 1.0f;
			px = e0x; py = e0y; pz = e0z;
			ent = 0; area = 0; planenum = 0;
			break;
		}
		const int nodenum = stack[sp].nodenum;
		bool blocked = false;
		int blockedArea = 0;
		if (nodenum < 0) {                                       // an area, :438-490
			if (!(__ldg(&m.presence[-nodenum]) & presencetype)) { blocked = true; blockedArea = -nodenum; }
			else { lastarea = -nodenum; continue; }
		} else if (nodenum == 0) {                               // a solid leaf, :492-528
			blocked = true;
		}
		if (blocked) {
			float cx = stack[sp].sx, cy = stack[sp].sy, cz = stack[sp].sz;
			float v1x, v1y, v1z;
			if (cx == s0x && cy == s0y && cz == s0z) {
				startsolid = 1; fraction = 0.0f;
				v1x = v1y = v1z = 0.0f;
			} else {
				startsolid = 0;
				v1x = e0x - s0x; v1y = e0y - s0y; v1z = e0z - s0z;
				const float v2x = cx - s0x, v2y = cy - s0y, v2z = cz - s0z;
				const float l2 = sqrtf((v2x * v2x + v2y * v2y) + v2z * v2z);          // VectorLength, q_shared.h:558
				const float len2 = (v1x * v1x + v1y * v1y) + v1z * v1z;               // VectorNormalize, q_math.c:857-874
				float length = len2;
				if (len2 != 0.0f) {
					const float ilength = 1.0f / (float)sqrt((double)len2);
					length = len2 * ilength;
					v1x = v1x * ilength; v1y = v1y * ilength; v1z = v1z * ilength;
				}
				fraction = l2 / length;
				cx = (float)((double)cx + (double)v1x * (-0.125));                     // VectorMA with a double scale, :459
				cy = (float)((double)cy + (double)v1y * (-0.125));
				cz = (float)((double)cz + (double)v1z * (-0.125));
			}
			px = cx; py = cy; pz = cz;
			ent = 0;
			area = blockedArea;
			planenum = stack[sp].planenum;
			const float4 pl = __ldg(&m.planes[planenum]);
			if ((v1x * pl.x + v1y * pl.y) + v1z * pl.z > 0) planenum ^= 1;             // :468-469
			break;
		}
		const float4 pl = __ldg(&m.nodePlane[nodenum]);
		const int4 nd = __ldg(&m.nodeInfo[nodenum]);
		const float csx = stack[sp].sx, csy = stack[sp].sy, csz = stack[sp].sz;
		const float cex = stack[sp].ex, cey = stack[sp].ey, cez = stack[sp].ez;
		float front = ((csx * pl.x + csy * pl.y) + csz * pl.z) - pl.w;
		const float back = ((cex * pl.x + cey * pl.y) + cez * pl.z) - pl.w;
		if (front >= 0 && back >= 0) {                           // ON_EPSILON is 0, :590-603
			stack[sp].nodenum = nd.y;
			sp++;
			if (sp >= kAasStack) break;                          // the reference's "stack overflow": the trace as it stands
		} else if (front < 0 && back < 0) {
			stack[sp].nodenum = nd.z;
			sp++;
			if (sp >= kAasStack) break;
		} else {
			const int tmpplanenum = stack[sp].planenum;
			if (front == back) front -= 0.001f;
			float frac;
			if (front < 0) frac = (float)(((double)front + 0.125) / (double)(front - back));
			else frac = (float)(((double)front - 0.125) / (double)(front - back));
			if (frac < 0) frac = 0.001f;
			else if (frac > 1) frac = 0.999f;
			const float mx = csx + (cex - csx) * frac, my = csy + (cey - csy) * frac, mz = csz + (cez - csz) * frac;
			const int side = front < 0 ? 1 : 0;
			stack[sp].sx = mx; stack[sp].sy = my; stack[sp].sz = mz;       // the far part keeps its end
			stack[sp].planenum = nd.x;
			stack[sp].nodenum = side ? nd.y : nd.z;                        // children[!side]
			sp++;
			if (sp >= kAasStack) break;
			stack[sp].sx = csx; stack[sp].sy = csy; stack[sp].sz = csz;
			stack[sp].ex = mx; stack[sp].ey = my; stack[sp].ez = mz;
			stack[sp].planenum = tmpplanenum;
			stack[sp].nodenum = side ? nd.z : nd.y;                        // children[side]
			sp++;
			if (sp >= kAasStack) break;
		}
	}
	int *out = q.out + 9 * (size_t)i;
	out[0] = startsolid; out[1] = __float_as_int(fraction);
	out[2] = __float_as_int(px); out[3] = __float_as_int(py); out[4] = __float_as_int(pz);
	out[5] = ent; out[6] = lastarea; out[7] = area; out[8] = planenum;
}

__global__ void __launch_bounds__(128) aasPointAreaNumKernel(const AasMap m, const int n, const float *points, int *out) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	const float x = points[3 * (size_t)i], y = points[3 * (size_t)i + 1], z = points[3 * (size_t)i + 2];
	int nodenum = 1;
	while (nodenum > 0) {
		const float4 pl = __ldg(&m.nodePlane[nodenum]);
		const int4 nd = __ldg(&m.nodeInfo[nodenum]);
		const float dist = ((x * pl.x + y * pl.y) + z * pl.z) - pl.w;
		nodenum = dist > 0 ? nd.y : nd.z;
	}
	out[i] = nodenum ? -nodenum : 0;
}

}  // namespace cmgpu
