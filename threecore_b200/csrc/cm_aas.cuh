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
	const int    *areaContents; // per area: areasettings[].contents (AREACONTENTS_*, aasfile.h:73-89)
	int numNodes, numPlanes, numAreas;
};

// aassettings' movement physics (be_aas_def.h:86-104, defaults in AAS_InitSettings, be_aas_move.c:74-92)
struct AasPhysics {
	float friction, stopspeed, gravity, waterfriction, watergravity, maxwalkvelocity, maxcrouchvelocity, maxswimvelocity,
	      walkaccelerate, airaccelerate, swimaccelerate, maxstep, maxsteepness, maxbarrier, jumpvel;
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

struct AasTrace { int startsolid; float fraction, px, py, pz; int ent, lastarea, area, planenum; };   // aas_trace_t

// AAS_TraceClientBBox(start, end, presencetype, -1); `stack`: kAasStack entries of the caller's (local memory)
__device__ __forceinline__ AasTrace aasTraceClientBBox(const AasMap &m, AasSeg *stack, const float s0x, const float s0y, const float s0z,
                                                       const float e0x, const float e0y, const float e0z, const int presencetype) {
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
	AasTrace t;
	t.startsolid = startsolid; t.fraction = fraction; t.px = px; t.py = py; t.pz = pz;
	t.ent = ent; t.lastarea = lastarea; t.area = area; t.planenum = planenum;
	return t;
}

#ifndef CMGPU_AAS_TRACE_BLOCKS
#define CMGPU_AAS_TRACE_BLOCKS 8
#endif
__global__ void __launch_bounds__(128, CMGPU_AAS_TRACE_BLOCKS) aasTraceClientBBoxKernel(const AasMap m, const AasTraceBatch q) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= q.n) return;
	AasSeg stack[kAasStack];
	const size_t o = 3 * (size_t)i;
	const AasTrace t = aasTraceClientBBox(m, stack, q.starts[o], q.starts[o + 1], q.starts[o + 2], q.ends[o], q.ends[o + 1], q.ends[o + 2],
	                                      q.presencetypes ? q.presencetypes[i] : q.sharedPresence);
	int *out = q.out + 9 * (size_t)i;
	out[0] = t.startsolid; out[1] = __float_as_int(t.fraction);
	out[2] = __float_as_int(t.px); out[3] = __float_as_int(t.py); out[4] = __float_as_int(t.pz);
	out[5] = t.ent; out[6] = t.lastarea; out[7] = t.area; out[8] = t.planenum;
}

__device__ __forceinline__ int aasPointAreaNum(const AasMap &m, const float x, const float y, const float z) {   // :212-258
	int nodenum = 1;
	while (nodenum > 0) {
		const float4 pl = __ldg(&m.nodePlane[nodenum]);
		const int4 nd = __ldg(&m.nodeInfo[nodenum]);
		const float dist = ((x * pl.x + y * pl.y) + z * pl.z) - pl.w;
		nodenum = dist > 0 ? nd.y : nd.z;
	}
	return nodenum ? -nodenum : 0;
}

__global__ void __launch_bounds__(128) aasPointAreaNumKernel(const AasMap m, const int n, const float *points, int *out) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	out[i] = aasPointAreaNum(m, points[3 * (size_t)i], points[3 * (size_t)i + 1], points[3 * (size_t)i + 2]);
}

// ---- AAS_PredictClientMovement (code/botlib/be_aas_move.c:495-979) ------------------------------------------------
// The bot library's movement predictor: up to maxframes frames of gravity, friction, acceleration by the command, a
// slide along the AAS with step-ups, and a dozen reasons to stop (SE_*, be_aas.h:163-176).  One thread per prediction;
// every query inside is one of the walks above or CM_PointContents of the world.  entnum is -1: entity collisions
// (AAS_AreaEntityCollision, be_aas_sample.c:472-487) are server traces and not part of this path.  Arithmetic widths
// follow the reference expression by expression -- including VectorMA's scale argument being evaluated once per
// component while the vector it reads is being overwritten (be_aas_move.c:792-799 with q_shared.h:495).
struct AasMoveBatch {
	int n;
	const float *origins, *velocities, *cmdmoves;        // [n][3]
	const int *presencetypes, *ongrounds, *cmdframes, *maxframes, *stopevents, *stopareanums;   // [n]
	const float *frametimes;                             // [n]
	int *out;                                            // aas_clientmove_t records, 21 ints each (be_aas.h:178-189)
	int *results;                                        // the function's return value per item (qtrue / qfalse)
};

enum : int { kSeHitGround = 1, kSeLeaveGround = 2, kSeEnterWater = 4, kSeEnterSlime = 8, kSeEnterLava = 16, kSeHitGroundDamage = 32, kSeGap = 64,
             kSeTouchJumppad = 128, kSeTouchTeleporter = 256, kSeEnterArea = 512, kSeHitGroundArea = 1024, kSeHitBoundingBox = 2048,
             kSeTouchClusterPortal = 4096 };

// AAS_TraceAreas, be_aas_sample.c:676-850: the areas a line passes through, with the points where it enters them
__device__ __forceinline__ int aasTraceAreas(const AasMap &m, AasSeg *stack, const float s0x, const float s0y, const float s0z,
                                             const float e0x, const float e0y, const float e0z, int *areas, float *points, const int maxareas) {
	int numareas = 0;
	areas[0] = 0;
	int sp = 1;
	stack[0].sx = s0x; stack[0].sy = s0y; stack[0].sz = s0z;
	stack[0].ex = e0x; stack[0].ey = e0y; stack[0].ez = e0z;
	stack[0].planenum = 0;
	stack[0].nodenum = 1;
	for (;;) {
		sp--;
		if (sp < 0) return numareas;
		const int nodenum = stack[sp].nodenum;
		if (nodenum < 0) {
			areas[numareas] = -nodenum;
			points[3 * numareas] = stack[sp].sx; points[3 * numareas + 1] = stack[sp].sy; points[3 * numareas + 2] = stack[sp].sz;
			numareas++;
			if (numareas >= maxareas) return numareas;
			continue;
		}
		if (!nodenum) continue;
		const float4 pl = __ldg(&m.nodePlane[nodenum]);
		const int4 nd = __ldg(&m.nodeInfo[nodenum]);
		const float csx = stack[sp].sx, csy = stack[sp].sy, csz = stack[sp].sz;
		const float cex = stack[sp].ex, cey = stack[sp].ey, cez = stack[sp].ez;
		const float front = ((csx * pl.x + csy * pl.y) + csz * pl.z) - pl.w;
		const float back = ((cex * pl.x + cey * pl.y) + cez * pl.z) - pl.w;
		if (front > 0 && back > 0) {
			stack[sp].nodenum = nd.y;
			sp++;
			if (sp >= kAasStack) return numareas;
		} else if (front <= 0 && back <= 0) {
			stack[sp].nodenum = nd.z;
			sp++;
			if (sp >= kAasStack) return numareas;
		} else {
			const int tmpplanenum = stack[sp].planenum;
			float frac = front / (front - back);
			if (frac < 0) frac = 0;
			else if (frac > 1) frac = 1;
			const float mx = csx + (cex - csx) * frac, my = csy + (cey - csy) * frac, mz = csz + (cez - csz) * frac;
			const int side = front < 0 ? 1 : 0;
			stack[sp].sx = mx; stack[sp].sy = my; stack[sp].sz = mz;
			stack[sp].planenum = nd.x;
			stack[sp].nodenum = side ? nd.y : nd.z;
			sp++;
			if (sp >= kAasStack) return numareas;
			stack[sp].sx = csx; stack[sp].sy = csy; stack[sp].sz = csz;
			stack[sp].ex = mx; stack[sp].ey = my; stack[sp].ez = mz;
			stack[sp].planenum = tmpplanenum;
			stack[sp].nodenum = side ? nd.z : nd.y;
			sp++;
			if (sp >= kAasStack) return numareas;
		}
	}
}

// DotProduct(normal, up) with up = {0, 0, 1}, as the macro evaluates it
__device__ __forceinline__ float aasDotUp(const float4 pl) { return (pl.x * 0.0f + pl.y * 0.0f) + pl.z * 1.0f; }

// AAS_ClipToBBox, be_aas_move.c:414-476 (changes the trace even when it returns false)
__device__ __forceinline__ bool aasClipToBBox(AasTrace &tr, const float *start, const float *end, const int presencetype, const float *mins, const float *maxs) {
	const float bmin[3] = {-15.0f, -15.0f, -24.0f};
	const float bmaxN[3] = {15.0f, 15.0f, 32.0f}, bmaxC[3] = {15.0f, 15.0f, 8.0f};     // AAS_PresenceTypeBoundingBox, be_aas_sample.c:72-88
	const float *bmax = presencetype == 2 ? bmaxN : bmaxC;
	float absmins[3], absmaxs[3], dir[3], mid[3] = {0.0f, 0.0f, 0.0f};
	for (int i = 0; i < 3; i++) { absmins[i] = mins[i] - bmax[i]; absmaxs[i] = maxs[i] - bmin[i]; }
	tr.px = end[0]; tr.py = end[1]; tr.pz = end[2];
	tr.fraction = 1;
	for (int i = 0; i < 3; i++) {
		if (start[i] < absmins[i] && end[i] < absmins[i]) return false;
		if (start[i] > absmaxs[i] && end[i] > absmaxs[i]) return false;
	}
	for (int i = 0; i < 3; i++) dir[i] = end[i] - start[i];
	float frac = 1;
	int i;
	for (i = 0; i < 3; i++) {
		if (fabsf(dir[i]) < 0.001f) continue;
		const float planedist = dir[i] > 0 ? absmins[i] : absmaxs[i];
		const float front = start[i] - planedist, back = end[i] - planedist;
		frac = front / (front - back);
		int side = i + 1;
		if (side > 2) side = 0;
		mid[side] = start[side] + dir[side] * frac;
		if (mid[side] > absmins[side] && mid[side] < absmaxs[side]) {
			side++;
			if (side > 2) side = 0;
			mid[side] = start[side] + dir[side] * frac;
			if (mid[side] > absmins[side] && mid[side] < absmaxs[side]) { mid[i] = planedist; break; }
		}
	}
	if (i != 3) {
		tr.startsolid = 0; tr.fraction = frac; tr.ent = 0; tr.planenum = 0; tr.area = 0; tr.lastarea = 0;
		tr.px = start[0] + dir[0] * frac; tr.py = start[1] + dir[1] * frac; tr.pz = start[2] + dir[2] * frac;
		return true;
	}
	return false;
}

#ifndef CMGPU_AAS_MOVE_BLOCKS
#define CMGPU_AAS_MOVE_BLOCKS 16   // 64 registers: 40.9 -> 55.8 M predictions/s (the predictor is latency-bound; its walks hide each other)
#endif
__global__ void __launch_bounds__(64, CMGPU_AAS_MOVE_BLOCKS) aasPredictClientMovementKernel(const AasMap m, const DevMap cm, const int haveWorld, const AasPhysics ph, const AasMoveBatch q) {
	const int item = blockIdx.x * blockDim.x + threadIdx.x;
	if (item >= q.n) return;
	AasSeg stack[kAasStack];
	int areas[20];
	float points[60];
	const size_t o3 = 3 * (size_t)item;
	int presencetype = q.presencetypes[item], onground = q.ongrounds[item];
	const int cmdframes = q.cmdframes[item], maxframes = q.maxframes[item], stopevent = q.stopevents[item], stopareanum = q.stopareanums[item];
	float frametime = q.frametimes[item];
	const float cmdx = q.cmdmoves[o3], cmdy = q.cmdmoves[o3 + 1], cmdz = q.cmdmoves[o3 + 2];
	const float bbmins[3] = {-4.0f, -4.0f, -4.0f}, bbmaxs[3] = {4.0f, 4.0f, 4.0f};      // AAS_PredictClientMovement, :973-974
	// the record: Com_Memset( move, 0 ) (:540)
	float endx = 0, endy = 0, endz = 0, velx = 0, vely = 0, velz = 0, mtime = 0;
	int endarea = 0, mpresence = 0, mstop = 0, mcontents = 0, mframes = 0, ret = 1;
	AasTrace mtrace;
	mtrace.startsolid = 0; mtrace.fraction = 0; mtrace.px = mtrace.py = mtrace.pz = 0; mtrace.ent = mtrace.lastarea = mtrace.area = mtrace.planenum = 0;
	AasTrace trace = mtrace;                               // Com_Memset( &trace, 0 ) (:541)

	auto pointContents = [&](const float x, const float y, const float z) -> int { return haveWorld ? pointContentsWorld(cm, x, y, z) : 0; };
	auto swimmingAt = [&](const float x, const float y, const float z) -> bool {     // AAS_Swimming, :215-223
		return (pointContents(x, y, z - 2.0f) & (8 | 16 | 32)) != 0;
	};
	if (frametime <= 0) frametime = 0.1f;
	const float jumpvel = ph.jumpvel * frametime;
	float ox = q.origins[o3], oy = q.origins[o3 + 1], oz = q.origins[o3 + 2];
	oz = (float)((double)oz + 0.25);
	float fvx = q.velocities[o3] * frametime, fvy = q.velocities[o3 + 1] * frametime, fvz = q.velocities[o3 + 2] * frametime;
	int jump_frame = -1;
	int n;
	bool stopped = false;
#define AAS_STOP(ex_, ey_, ez_, area_, scaled_, withTrace_, ev_, contents_) { \
		endx = (ex_); endy = (ey_); endz = (ez_); endarea = (area_); \
		if (scaled_) { const float inv_ = 1 / frametime; velx = fvx * inv_; vely = fvy * inv_; velz = fvz * inv_; } else { velx = fvx; vely = fvy; velz = fvz; } \
		if (withTrace_) mtrace = trace; \
		mstop = (ev_); mpresence = presencetype; mcontents = (contents_); mtime = n * frametime; mframes = n; stopped = true; }
	for (n = 0; n < maxframes && !stopped; n++) {
		const bool swimming = swimmingAt(ox, oy, oz);
		const float gravity = swimming ? ph.watergravity : ph.gravity;
		fvz = (float)((double)fvz - ((double)gravity * 0.1 * (double)frametime));
		if (onground || swimming) {
			const float friction = swimming ? ph.waterfriction : ph.friction;
			const float inv = 1 / frametime;
			fvx = fvx * inv; fvy = fvy * inv; fvz = fvz * inv;
			{	// AAS_ApplyFriction, :391-407
				const float speed = (float)sqrt((double)(fvx * fvx + fvy * fvy));
				if (speed != 0.0f) {
					const float control = speed < ph.stopspeed ? ph.stopspeed : speed;
					float newspeed = speed - frametime * control * friction;
					if (newspeed < 0) newspeed = 0;
					newspeed /= speed;
					fvx *= newspeed; fvy *= newspeed;
				}
			}
			fvx = fvx * frametime; fvy = fvy * frametime; fvz = fvz * frametime;
		}
		bool crouch = false;
		if (n < cmdframes) {
			float maxvel = ph.maxwalkvelocity, accelerate = ph.airaccelerate;
			float wx = cmdx, wy = cmdy, wz = cmdz;
			if (onground) {
				if (cmdz < -300) { crouch = true; maxvel = ph.maxcrouchvelocity; }
				if (!swimming && cmdz > 1) {
					fvz = (float)((double)jumpvel - ((double)gravity * 0.1 * (double)frametime) + 5);
					jump_frame = n;
					accelerate = ph.airaccelerate;
				} else {
					accelerate = ph.walkaccelerate;
				}
			}
			if (swimming) { maxvel = ph.maxswimvelocity; accelerate = ph.swimaccelerate; }
			else wz = 0;
			float wishspeed;
			{	// VectorNormalize, q_math.c:857-874
				const float len2 = (wx * wx + wy * wy) + wz * wz;
				wishspeed = len2;
				if (len2 != 0.0f) {
					const float il = 1 / (float)sqrt((double)len2);
					wishspeed = len2 * il;
					wx *= il; wy *= il; wz *= il;
				}
			}
			if (wishspeed > maxvel) wishspeed = maxvel;
			const float inv = 1 / frametime;
			fvx = fvx * inv; fvy = fvy * inv; fvz = fvz * inv;
			{	// AAS_Accelerate, :364-383
				const float currentspeed = (fvx * wx + fvy * wy) + fvz * wz;
				const float addspeed = wishspeed - currentspeed;
				if (!(addspeed <= 0)) {
					float accelspeed = accelerate * frametime * wishspeed;
					if (accelspeed > addspeed) accelspeed = addspeed;
					fvx += accelspeed * wx; fvy += accelspeed * wy; fvz += accelspeed * wz;
				}
			}
			fvx = fvx * frametime; fvy = fvy * frametime; fvz = fvz * frametime;
		}
		if (crouch) presencetype = 4;
		else if (presencetype == 4) {
			const int a = aasPointAreaNum(m, ox, oy, oz);                        // AAS_PointPresenceType, be_aas_sample.c:340-349
			const int pt = a ? __ldg(&m.presence[a]) : 1;
			if (pt & 2) presencetype = 2;
		}
		const float lastx = ox, lasty = oy, lastz = oz;
		float lvx = fvx, lvy = fvy, lvz = fvz;
		int j = 0;
		do {
			const float ex = ox + lvx, ey = oy + lvy, ez = oz + lvz;
			trace = aasTraceClientBBox(m, stack, ox, oy, oz, ex, ey, ez, presencetype);
			if (stopevent & (kSeEnterArea | kSeTouchJumppad | kSeTouchTeleporter | kSeTouchClusterPortal)) {
				const int numareas = aasTraceAreas(m, stack, ox, oy, oz, trace.px, trace.py, trace.pz, areas, points, 20);
				for (int i = 0; i < numareas && !stopped; i++) {
					const int ac = __ldg(&m.areaContents[areas[i]]);
					if ((stopevent & kSeEnterArea) && areas[i] == stopareanum) AAS_STOP(points[3 * i], points[3 * i + 1], points[3 * i + 2], areas[i], true, true, kSeEnterArea, 0)
					else if ((stopevent & kSeTouchJumppad) && n && (ac & 128)) AAS_STOP(points[3 * i], points[3 * i + 1], points[3 * i + 2], areas[i], true, true, kSeTouchJumppad, 0)
					else if ((stopevent & kSeTouchTeleporter) && (ac & 64)) AAS_STOP(points[3 * i], points[3 * i + 1], points[3 * i + 2], areas[i], true, true, kSeTouchTeleporter, 0)
					else if ((stopevent & kSeTouchClusterPortal) && (ac & 8)) AAS_STOP(points[3 * i], points[3 * i + 1], points[3 * i + 2], areas[i], true, true, kSeTouchClusterPortal, 0)
				}
				if (stopped) break;
			}
			if (stopevent & kSeHitBoundingBox) {
				const float st[3] = {ox, oy, oz}, en[3] = {trace.px, trace.py, trace.pz};
				if (aasClipToBBox(trace, st, en, presencetype, bbmins, bbmaxs)) {
					AAS_STOP(trace.px, trace.py, trace.pz, aasPointAreaNum(m, trace.px, trace.py, trace.pz), true, true, kSeHitBoundingBox, 0)
					break;
				}
			}
			ox = trace.px; oy = trace.py; oz = trace.pz;
			if (trace.fraction < 1.0f) {
				const float4 plane = __ldg(&m.planes[trace.planenum]);
				if (stopevent & kSeHitGroundArea) {
					if (aasDotUp(plane) > ph.maxsteepness) {
						const float sz2 = (float)((double)oz + 0.5);
						if (aasPointAreaNum(m, ox, oy, sz2) == stopareanum) {
							AAS_STOP(ox, oy, sz2, stopareanum, true, true, kSeHitGroundArea, 0)
							break;
						}
					}
				}
				bool step = false;
				if (plane.z == 0 && (jump_frame < 0 || n - jump_frame > 2)) {
					float stx = (float)((double)ox + (double)plane.x * (-0.25));
					float sty = (float)((double)oy + (double)plane.y * (-0.25));
					float stz = (float)((double)oz + (double)plane.z * (-0.25));
					const float sex = stx, sey = sty, sez = stz;
					stz += ph.maxstep;
					const AasTrace steptrace = aasTraceClientBBox(m, stack, stx, sty, stz, sex, sey, sez, presencetype);
					if (!steptrace.startsolid) {
						const float4 plane2 = __ldg(&m.planes[steptrace.planenum]);
						if (aasDotUp(plane2) > ph.maxsteepness) {
							lvx = ex - steptrace.px; lvy = ey - steptrace.py; lvz = ez - steptrace.pz;
							lvz = 0;
							fvz = 0;
							oz = steptrace.pz;
							step = true;
						}
					}
				}
				if (!step) {
					// VectorMA( v, -DotProduct( v, normal ), normal, v ): the scale is evaluated again for every component,
					// with the components already written (q_shared.h:495)
					lvx = lvx + plane.x * (-((lvx * plane.x + lvy * plane.y) + lvz * plane.z));
					lvy = lvy + plane.y * (-((lvx * plane.x + lvy * plane.y) + lvz * plane.z));
					lvz = lvz + plane.z * (-((lvx * plane.x + lvy * plane.y) + lvz * plane.z));
					const float oldz = fvz;
					fvx = fvx + plane.x * (-((fvx * plane.x + fvy * plane.y) + fvz * plane.z));
					fvy = fvy + plane.y * (-((fvx * plane.x + fvy * plane.y) + fvz * plane.z));
					fvz = fvz + plane.z * (-((fvx * plane.x + fvy * plane.y) + fvz * plane.z));
					if (aasDotUp(plane) > ph.maxsteepness) onground = 1;
					if (stopevent & kSeHitGroundDamage) {
						float delta = 0;
						if (oldz < 0 && fvz > oldz && !onground) delta = oldz;
						else if (onground) delta = fvz - oldz;
						if (delta != 0.0f) {
							delta = delta * 10;
							delta = (float)((double)(delta * delta) * 0.0001);
							if (swimming) delta = 0;
							if (delta > 40) {
								AAS_STOP(ox, oy, oz, aasPointAreaNum(m, ox, oy, oz), false, true, kSeHitGroundDamage, 0)
								break;
							}
						}
					}
				}
			}
			if (++j > 20) { ret = 0; stopped = true; break; }              // "extra check to prevent endless loop": qfalse, the record zeroed
		} while (trace.fraction < 1.0f);
		if (stopped) break;
		if (fvz <= 10) {
			const float fz = oz - 22.0f;
			const int pc = pointContents(ox, oy, fz);
			int event = 0;
			if (pc & 8) event |= kSeEnterLava;
			if (pc & 16) event |= kSeEnterSlime;
			if (pc & 32) event |= kSeEnterWater;
			const int areanum = aasPointAreaNum(m, ox, oy, oz);
			const int ac = __ldg(&m.areaContents[areanum]);
			if (ac & 2) event |= kSeEnterLava;
			if (ac & 4) event |= kSeEnterSlime;
			if (ac & 1) event |= kSeEnterWater;
			if (event & stopevent) {
				AAS_STOP(ox, oy, oz, areanum, true, false, event & stopevent, pc)
				break;
			}
		}
		{	// AAS_OnGround, :185-207
			const AasTrace g = aasTraceClientBBox(m, stack, ox, oy, oz, ox, oy, oz - 10.0f, presencetype);
			onground = 0;
			if (!g.startsolid && !(g.fraction >= 1.0f) && !(oz - g.pz > 10)) {
				if (!(aasDotUp(__ldg(&m.planes[g.planenum])) < ph.maxsteepness)) onground = 1;
			}
		}
		if (onground) {
			if (stopevent & kSeHitGround) { AAS_STOP(ox, oy, oz, aasPointAreaNum(m, ox, oy, oz), true, true, kSeHitGround, 0) break; }
		} else if (stopevent & kSeLeaveGround) {
			AAS_STOP(ox, oy, oz, aasPointAreaNum(m, ox, oy, oz), true, true, kSeLeaveGround, 0)
			break;
		} else if (stopevent & kSeGap) {
			const float gz = oz - (48 + ph.maxbarrier);
			const AasTrace gap = aasTraceClientBBox(m, stack, ox, oy, oz, ox, oy, gz, 4);
			if (!gap.startsolid) {
				if (gap.pz < oz - ph.maxstep - 1) {
					if (!(pointContents(ox, oy, gz) & 32)) {
						AAS_STOP(lastx, lasty, lastz, aasPointAreaNum(m, lastx, lasty, lastz), true, true, kSeGap, 0)
						break;
					}
				}
			}
		}
	}
	if (!stopped) AAS_STOP(ox, oy, oz, aasPointAreaNum(m, ox, oy, oz), true, false, 0, 0)
#undef AAS_STOP
	int *out = q.out + 21 * (size_t)item;
	if (!ret) {                                            // qfalse: nothing but the memset reached the record
		for (int k = 0; k < 21; k++) out[k] = 0;
	} else {
		out[0] = __float_as_int(endx); out[1] = __float_as_int(endy); out[2] = __float_as_int(endz);
		out[3] = endarea;
		out[4] = __float_as_int(velx); out[5] = __float_as_int(vely); out[6] = __float_as_int(velz);
		out[7] = mtrace.startsolid; out[8] = __float_as_int(mtrace.fraction);
		out[9] = __float_as_int(mtrace.px); out[10] = __float_as_int(mtrace.py); out[11] = __float_as_int(mtrace.pz);
		out[12] = mtrace.ent; out[13] = mtrace.lastarea; out[14] = mtrace.area; out[15] = mtrace.planenum;
		out[16] = mpresence; out[17] = mstop; out[18] = mcontents; out[19] = __float_as_int(mtime); out[20] = mframes;
	}
	if (q.results) q.results[item] = ret;
}

}  // namespace cmgpu
