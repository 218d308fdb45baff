// cm_gpu.cu -- C ABI of libcmgpu.so: context, map upload (flat arrays -> packed
// device layout), batch launches and the host<->device pipeline.
//
// Boundary it implements: include/cm_gpu.h.  Reference call sites it stands in
// for: CM_BoxTrace / CM_TransformedBoxTrace / CM_PointContents
// (code/qcommon/cm_public.h:42-52) and the upload hook after
// code/qcommon/cm_load.c:734.
#include "cm_kernels.cuh"
#include "cm_trace_fast.cuh"
#include "cm_trace_pool.cuh"
#include "cm_trace_simple.cuh"
#include "cm_sv_world.cuh"
#include "cm_leaf_queries.cuh"
#include "cm_aas.cuh"
#include "../../include/cm_gpu.h"
#include "cm_patchgen.h"
#include "cm_host_records.h"
#include "cm_patchgen_core.h"

#include <cuda.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <chrono>
#include <memory>
#include <new>
#include <stddef.h>
#include <string>
#include <thread>
#include <vector>

using namespace cmgpu;

static_assert(sizeof(cmgpu_trace_t) == 56, "trace_t layout is ABI (q_shared.h:976-986)");
static_assert(sizeof(Counters) == sizeof(cmgpu_counters), "counter layouts must agree");

namespace {

thread_local std::string g_createError;

constexpr int kCursorSlots = 64;
constexpr size_t kPoolSmem = (size_t)(kBlock / 32) * PF_COUNT * CMGPU_POOL_K * 32 * sizeof(float);
#ifndef CMGPU_PIPE_STREAMS
#define CMGPU_PIPE_STREAMS 3
#endif
constexpr int kBinSlots = CMGPU_PIPE_STREAMS;   // pipeline streams of the host-pointer calls, each with its own workspaces
constexpr int kPipeChunks = 64;          // upper bound on the pieces a host batch is cut into (copies overlap the kernels)
constexpr size_t kMinChunkItems = 1 << 16;
constexpr int kHullTables = 8;            // hulls whose node tables and entry grids are kept per map

// calls of a few items (traceHostTiny): where their arguments and results live in the context's mapped pinned block
constexpr int kTinyMax = 128;
struct TinyLayout {
	static constexpr size_t v3 = 12 * (size_t)kTinyMax, i4 = 4 * (size_t)kTinyMax;
	static constexpr size_t starts = 0, ends = starts + v3, mins = ends + v3, maxs = mins + v3, models = maxs + v3, masks = models + i4,
	                        origins = masks + i4, rot = origins + v3, rotIdx = rot + 36 * (size_t)kTinyMax, boxmins = rotIdx + i4,
	                        boxmaxs = boxmins + v3, out = boxmaxs + v3, hit = out + sizeof(cmgpu_trace_t) * (size_t)kTinyMax,
	                        status = hit + i4, cursor = status + 64, total = cursor + 64;
};


struct DevBuf {
	void *p = nullptr;
	size_t cap = 0;
};

}  // namespace

struct cmgpu_ctx_s {
	int device = 0;
	std::string err;
	bool hasMap = false;
	DevMap map{};
	std::vector<void *> mapAllocs;
	std::vector<char> slabHost;          // staging of the map slab during an upload
	std::vector<std::pair<const void **, size_t>> slabFixups;
	void *mapSlab = nullptr;
	size_t mapSlabBytes = 0, mapWindowBytes = 0;
	// per-hull tables live right behind the map in the same allocation (kHullTables slots of hullArenaPer bytes): one L2
	// access-policy window per launch is all CUDA gives, and without it their lines count as streaming data
	size_t hullArenaPer = 0, hullArenaOff = 0;
	int hullSlotsUsed = 0;
	bool hullWindow = true;              // "hull_window" 0: the window covers the map only (as in round 1)
	bool l2Window = true;
	int numModels = 0;
	int noCurves = 0, playerCurveClip = 1;
	bool counting = false;
	Counters *dCounters = nullptr;
	int *dStatus = nullptr;
	int *dCursor = nullptr;              // work cursors of the persistent kernel, one per launch slot
	unsigned cursorSlot = 0;
	int gridPlain = 0, gridCount = 0;    // resident CTAs of the two kernel variants
	int gridFast = 0, gridFastCount = 0;
	int gridPool = 0, gridPoolCount = 0;
	bool usePool = true;                 // hot kernel variant with K items per lane in shared memory (cm_trace_pool.cuh)
	int poolReps = 1;                    // its repeats of a light step per round (a node step is up to 3 levels, a leaf step 2 entries)
	int minChunk = 4;                    // fewest items a warp takes at a time in a small launch
	int simpleSpread = -1;               // thread-per-item kernel: log2 of the lanes per item; -1 = as many as fill the GPU's warp slots once
	int gridSimple = 0;                  // resident CTAs of the thread-per-item kernel
	int simpleMaxItems = 786432;         // launches up to this size use the thread-per-item kernel (lowest latency)
	int poolMinItems = 196609;           // device-pointer launches from this size up take the pool kernel, smaller ones the thread-per-item
	                                     // kernel.  The crossover depends on the sweeps: ~128 Ki items for the bench's long ones (128 Ki 313 / 293 us,
	                                     // 256 Ki 369 / 350, 512 Ki 576 / 467, 768 Ki 752 / 571), beyond 192 Ki for a rollout's short ones
	int pipePoolMinItems = 786433;       // the same for the pieces of the host pipeline (PCIe-bound; each piece a launch on one of three streams)
	FastMap fast{};                      // layout of the hot kernel (cm_trace_fast.cuh)
	int numPatches = 0;
	int treeDepth = 1;                   // levels of the world BSP (validate()): the size of every traversal stack
	bool useFast = true;
	int pipeChunks = 32;                 // pieces a host batch is cut into
	bool pipeRamp = true;                // ... the first ones smaller
	// optional CUDA-event timing of the trace kernel alone (cmgpu_kernel_timing)
	bool timing = false;
	std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timingPairs;
	size_t timingUsed = 0;
	bool slabOK = false;                 // all brush bounds below 2^15 in magnitude
	bool slabTest = true;
	int gang = kGang;
	int gangPool = 3;                    // pool kernel: eighths of the most popular kind's lane count a kind needs to run (measured: 3 beats 4 by 1 %)
	int gangFetch = kGang;
	int gangBrush = 5;                   // pool kernel: eighths of the most popular kind's lane count that must wait for a brush clip ...
	int gangFetchPool = 6;               // ... and for the fetch step, before it runs (measured: +0.7 % over 4 / 4)
	int lightReps = 3;
	int facetSteps = 8;                  // lanes that must have reached a patch before the warp takes the patches together (counting variant: planes per round)
	// spatial binning of large batches (counting sort by start cell)
	BinGrid grid{};
	bool binning = true;
	bool binDir = false;
	int binMinItems = 98304;             // smaller batches are traced in input order (the five binning launches cost ~50 us)
	float binCell = 256.0f, binLongLen = 512.0f;
	DevBuf bKeys, bRecs, bCompact, bHist[kBinSlots];
	bool capturing = false;              // the stream of the current device-pointer call is being captured into a CUDA graph
	// A captured launch bakes the addresses of the workspaces it used (binning buffers, stack arena, pool headers, hull
	// tables) into the graph.  Once a capture has used this context, a workspace that has to grow -- or a hull table that
	// has to make room -- is RETIRED instead of freed: replays keep reading and writing their own, still valid blocks.
	// Retired blocks are released with the map (a graph does not outlive the map it was recorded on) or the context.
	// records that leave this GPU (a peer's HBM over NVLink, sharding.PeerResults): the batch is traced in pieces and the
	// expansion of piece p -- the only kernel that writes records -- runs on a side stream with a few CTAs, next to the
	// walk of piece p + 1: the transfer hides behind the walks instead of following them (tracePieces)
	cudaStream_t sideStream = nullptr;
	cudaEvent_t evWalked = nullptr, evSideDone = nullptr;
	int remotePieces = 4, remoteExpandCtas = 64;
	std::vector<std::pair<uintptr_t, uintptr_t>> peerRanges;   // peer buffers opened with cmgpu_ipc_open: [base, end)
	bool remotePipeline = false;         // measured on 2 B200s: the per-piece launches cost more than the overlap saves; "remote_pipeline" turns it on
	bool graphPinned = false;
	std::vector<void *> retired;
	bool twoPhase = true, twoPhaseHost = false, twoPhaseSmall = false, compactByItem = true;
	// pool kernel: per-hull node threshold tables (nodeThresholdKernel), built on first use, dropped with the map
	struct HullTable { NodeHull key; int4 *nodes; uint2 *cells; bool usable, gridUsable; };   // slots of the arena behind the map
	EntryGrid entry{};                   // geometry of the entry grid and the heap of the top levels (cellRec is per hull: HullTable::cells)
	bool entryGrid = true;
	float entryCell = 128.0f;            // target cell size of the entry grid, world units
	std::vector<HullTable> hullTables;
	bool nodeThresholds = true;                // binned hot launches: 16-byte results by slot + expandRecordsKernel
	DevBuf bPoolHdr[kBinSlots + 1];
	DevBuf bStackArena[kBinSlots + 1];   // traversal stacks of the hot kernel: one per pipeline stream + one for device-pointer calls
	// Device-pointer calls share one workspace (binning buffers, stack arena).  extDone fires when the last such
	// launch has finished; a call on a different stream, or a host-pointer call, waits for it first.
	cudaEvent_t extDone = nullptr;
	cudaStream_t extStream = nullptr;
	bool extBusy = false;
	int *hStatus = nullptr;              // pinned
	unsigned long long launches = 0;
	cudaStream_t streams[kBinSlots] = {};
	cudaEvent_t evt[kPipeChunks] = {};
	// grow-only staging for the host-pointer entry points
	DevBuf bStarts, bEnds, bMins, bMaxs, bModels, bMasks, bOrigins, bRot, bRotIdx, bBoxMins, bBoxMaxs, bOut, bHit;
	std::vector<float> hostRot;
	std::vector<int32_t> hostRotIdx;
	// SV_Trace on the device (cm_sv_world.cuh): the linked entities in sector order, and the workspaces of a batch
	float worldBounds[6] = {};           // CM_ModelBounds(0): what SV_ClearWorld builds the sector tree from
	SvWorld sv{};
	bool hasSvWorld = false;
	DevBuf bSvWorld, bSvOffsets, bSvTiles, bSvItems, bSvCand, bSvPass;
	const int *leafCluster = nullptr;    // [numLeafs][2] cluster, area (cm_leaf_queries.cuh); part of the map slab
	DevBuf bLeafLists, bLeafCounts;
	// visibility and areas (cm_leaf_queries.cuh).  The matrix can be tens of MB: its own allocation, not the map slab
	DevBuf bVis, bFlood, bEntRecs, bVisible;
	bool vised = false;
	int numClusters = 0, clusterBytes = 0, numAreas = 0, noAreas = 0;
	std::vector<int> areaPortals, floodnum;          // host copies: [numAreas][numAreas] reference counts, [numAreas]
	unsigned *hSvTotal = nullptr;        // pinned
	// A call of a few items (the drop-in CM_BoxTrace is a batch of one) is all latency: its arguments and results live
	// in one block of MAPPED pinned memory that the kernel reads and writes over PCIe itself -- one launch, one
	// synchronisation, no copy calls, the status word included (traceHostTiny).
	char *tinyHost = nullptr, *tinyDev = nullptr;
	bool tinyCalls = true;
	// botlib's AAS world (cm_aas.cuh): its own tree, independent of the clip map
	AasMap aas{};
	AasPhysics aasPhys{6.0f, 100.0f, 800.0f, 1.0f, 400.0f, 320.0f, 100.0f, 150.0f, 10.0f, 1.0f, 4.0f, 19.0f, 0.7f, 33.0f, 270.0f};   // AAS_InitSettings, be_aas_move.c:74-92
	DevBuf bAas, bAasIn, bAasOut;
	bool hasAas = false;
	// host regions pinned through this context (cmgpu_host_register, or "auto_register"): page-aligned [lo, hi)
	std::vector<std::pair<uintptr_t, uintptr_t>> hostRegs;
	bool autoRegister = false;
	// Host half of a big host batch's results (traceHostHybrid, cm_host_records.h): a share of the pieces comes back as the
	// walk's 16-byte results and the calling process's worker threads write the 56-byte records from them.
	int hostRecords = -1;                // "host_records": -1 = the split follows the measured rates, 0 = off (every record over
	                                     // PCIe), 1 = every record written by the host, 2 = the new pipeline with no host share,
	                                     // 3 = the new pipeline with the measured share, whatever the plain one would do
	int hostThreads = -1;                // "host_threads": -1 = host cores per visible GPU, minus one for the caller
	size_t hybridMinItems = (size_t)1 << 21, hybridWalkItems = (size_t)1 << 20;
	hostrec::Pool *hostPool = nullptr;
	std::vector<hostrec::SideRec> hostSides;         // [numBrushSides + 1], entry 0 = no side
	std::vector<int32_t> hostBrushContents;          // [numBrushes + 1], entry 0 = no brush
	char *hCompact = nullptr;            // pinned: where the 16-byte results land
	size_t hCompactCap = 0;
	float *hStage = nullptr;             // pinned: pageable starts / ends on their way up (the threads copy, the link uploads)
	size_t hStageCap = 0;
	bool stagePageable = true;           // "stage_pageable"
	bool hybridUnavailable = false;      // the host would not pin the staging memory: big batches take the plain pipeline
	std::vector<cudaEvent_t> hybridEvents;
	// the share of a batch the threads get: set from the thread count before the first call, then corrected after every
	// call by who finished later, the link or the threads (both are busy to the end when the share is right)
	double hostShare = -1.0;
	int hybridGridEighths = 5;           // "hybrid_grid_eighths": share of the resident CTAs one piece's walk may take
	int hybridInFlight = 2 * kBinSlots;  // pieces queued ahead of the oldest one whose results are still on their way
	double hostShareFixed = -1.0;        // "host_share": >= 0 pins the share (experiments)
	int hybridDecouple = 2;              // "hybrid_decouple": 0 = a piece's upload, walk and result copy on one stream; 1 = uploads on a stream
	                                     // of their own and result copies on two more (one per kind); 2 = one stream for all result copies
	cudaStream_t inStream = nullptr, outHostStream = nullptr, outDevStream = nullptr;
	// big host batches, both pipelines timed (ns per item, running means): [0] = every record over PCIe (round 1), [1] = shared
	double pipeNs[2] = {0.0, 0.0};
	unsigned pipeCalls[2] = {0, 0};
	double d2hRate = 50e9;               // bytes per second of a result copy, measured on the way
	unsigned long long hostRecordsWritten = 0, deviceRecordsWritten = 0;   // of the last hybrid call
	unsigned long long svCandidates = 0;
};

namespace {

int fail(cmgpu_ctx *ctx, int code, const char *fmt, ...) {
	char buf[512];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	if (ctx) ctx->err = buf; else g_createError = buf;
	return code;
}

#define CUDA_TRY(ctx, expr)                                                                     \
	do {                                                                                        \
		cudaError_t e_ = (expr);                                                                \
		if (e_ != cudaSuccess)                                                                  \
			return fail(ctx, CMGPU_ERR_CUDA, "%s: %s (%s:%d)", #expr, cudaGetErrorString(e_),   \
			            __FILE__, __LINE__);                                                    \
	} while (0)

int reserve(cmgpu_ctx *ctx, DevBuf &b, size_t bytes) {
	if (bytes <= b.cap) return CMGPU_OK;
	if (ctx->capturing)                              // cudaFree / cudaMalloc are not allowed while a stream is being captured
		return fail(ctx, CMGPU_ERR_INVALID, "a workspace would have to grow during stream capture: run the same call once before capturing it");
	if (b.p && ctx->graphPinned) ctx->retired.push_back(b.p);   // a recorded graph may still name it (see graphPinned)
	else if (b.p) cudaFree(b.p);                 // cudaFree synchronises the device: nothing is still reading the old block
	b.p = nullptr;
	b.cap = 0;
	size_t want = bytes + bytes / 4 + 256;
	CUDA_TRY(ctx, cudaMalloc(&b.p, want));
	b.cap = want;
	return CMGPU_OK;
}

void releaseBuf(DevBuf &b) {
	if (b.p) cudaFree(b.p);
	b.p = nullptr;
	b.cap = 0;
}

// All arrays of a map go into ONE device allocation (256-byte aligned pieces), so that a single L2 access-policy
// window can cover the map: the kernels' L1 absorbs most hits on it, which makes the map look cold to L2's
// replacement next to the GBs of rays and records streaming through -- without the window its lines are evicted and
// re-read from HBM all the time.  uploadArray stages a piece and notes where its device pointer goes; commitMapSlab
// allocates, copies and patches the pointers.
template <typename T>
int uploadArray(cmgpu_ctx *ctx, const std::vector<T> &host, const T **dev) {
	const size_t bytes = (host.size() ? host.size() : 1) * sizeof(T);
	const size_t off = (ctx->slabHost.size() + 255) & ~(size_t)255;
	ctx->slabHost.resize(off + bytes, 0);
	if (host.size()) memcpy(ctx->slabHost.data() + off, host.data(), host.size() * sizeof(T));
	ctx->slabFixups.push_back({reinterpret_cast<const void **>(dev), off});
	*dev = nullptr;
	return CMGPU_OK;
}

int commitMapSlab(cmgpu_ctx *ctx) {
	void *p = nullptr;
	const size_t mapBytes = ((ctx->slabHost.size() ? ctx->slabHost.size() : 256) + 255) & ~(size_t)255;
	const size_t bytes = mapBytes + (size_t)kHullTables * ctx->hullArenaPer;
	ctx->hullArenaOff = mapBytes;
	ctx->hullSlotsUsed = 0;
	CUDA_TRY(ctx, cudaMalloc(&p, bytes));
	ctx->mapAllocs.push_back(p);
	CUDA_TRY(ctx, cudaMemcpy(p, ctx->slabHost.data(), ctx->slabHost.size(), cudaMemcpyHostToDevice));
	for (auto &fx : ctx->slabFixups) *fx.first = (const char *)p + fx.second;
	ctx->slabFixups.clear();
	std::vector<char>().swap(ctx->slabHost);
	ctx->mapSlab = p;
	ctx->mapSlabBytes = bytes;
	// reserve that much of L2 for persisting lines (the hardware caps both numbers)
	cudaDeviceProp prop;
	CUDA_TRY(ctx, cudaGetDeviceProperties(&prop, ctx->device));
	ctx->mapWindowBytes = 0;
	if (ctx->l2Window && prop.persistingL2CacheMaxSize > 0 && prop.accessPolicyMaxWindowSize > 0) {
		// set aside room for the map and two hulls' tables; the window of a launch covers the map and the slots in use
		const size_t aside = mapBytes + 2 * ctx->hullArenaPer;
		size_t want = aside < (size_t)prop.persistingL2CacheMaxSize ? aside : (size_t)prop.persistingL2CacheMaxSize;
		if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess)
			ctx->mapWindowBytes = bytes < (size_t)prop.accessPolicyMaxWindowSize ? bytes : (size_t)prop.accessPolicyMaxWindowSize;
		else cudaGetLastError();
	}
	return CMGPU_OK;
}

// launch with the map's L2 window attached to this launch only (no stream attribute is touched: device-pointer
// calls run on the caller's stream)
template <typename... KArgs, typename... Args>
cudaError_t launchOnMap(cmgpu_ctx *ctx, void (*kernel)(KArgs...), int blocks, int threads, size_t smem, cudaStream_t s, Args... args) {
	cudaLaunchConfig_t cfg{};
	cfg.gridDim = dim3((unsigned)blocks);
	cfg.blockDim = dim3((unsigned)threads);
	cfg.dynamicSmemBytes = smem;
	cfg.stream = s;
	cudaLaunchAttribute attr[1];
	int nattr = 0;
	if (ctx->mapWindowBytes) {
		attr[0].id = cudaLaunchAttributeAccessPolicyWindow;
		attr[0].val.accessPolicyWindow.base_ptr = ctx->mapSlab;
		size_t inUse = ctx->hullArenaOff + ctx->hullTables.size() * ctx->hullArenaPer;
		if (inUse > ctx->mapWindowBytes || !ctx->hullWindow) inUse = ctx->hullWindow ? ctx->mapWindowBytes : (ctx->hullArenaOff < ctx->mapWindowBytes ? ctx->hullArenaOff : ctx->mapWindowBytes);
		attr[0].val.accessPolicyWindow.num_bytes = inUse;
		attr[0].val.accessPolicyWindow.hitRatio = 1.0f;
		attr[0].val.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
		attr[0].val.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
		nattr = 1;
	}
	cfg.attrs = attr;
	cfg.numAttrs = nattr;
	return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

void freeMap(cmgpu_ctx *ctx) {
	for (void *p : ctx->mapAllocs) cudaFree(p);
	ctx->mapAllocs.clear();
	if (ctx->mapWindowBytes) cudaCtxResetPersistingL2Cache();
	ctx->mapSlab = nullptr;
	ctx->mapSlabBytes = ctx->mapWindowBytes = 0;
	ctx->slabFixups.clear();
	std::vector<char>().swap(ctx->slabHost);
	ctx->hasMap = false;
	ctx->numModels = 0;
	memset(&ctx->map, 0, sizeof(ctx->map));
	memset(&ctx->fast, 0, sizeof(ctx->fast));
	ctx->numPatches = 0;
	ctx->hullTables.clear();
	for (void *p : ctx->retired) cudaFree(p);
	ctx->retired.clear();
	ctx->graphPinned = false;
	ctx->leafCluster = nullptr;
	ctx->vised = false;
	ctx->numClusters = ctx->clusterBytes = ctx->numAreas = 0;
	ctx->areaPortals.clear();
	ctx->floodnum.clear();
}

// CM_FloodAreaConnections (cm_test.c:322-365): flood numbers from the portal reference counts, then to the device.
// Areas are numbered by the lowest area of their connected set, in index order, like the reference's recursion.
int floodAreas(cmgpu_ctx *ctx) {
	const int n = ctx->numAreas;
	ctx->floodnum.assign((size_t)n, 0);
	std::vector<int> work;
	int next = 0;
	for (int i = 0; i < n; i++) {
		if (ctx->floodnum[i]) continue;
		ctx->floodnum[i] = ++next;
		work.assign(1, i);
		while (!work.empty()) {
			const int a = work.back();
			work.pop_back();
			for (int j = 0; j < n; j++) {
				if (ctx->areaPortals[(size_t)a * n + j] > 0 && !ctx->floodnum[j]) {
					ctx->floodnum[j] = next;
					work.push_back(j);
				}
			}
		}
	}
	if (n == 0) return CMGPU_OK;
	int rc = reserve(ctx, ctx->bFlood, (size_t)n * sizeof(int));
	if (rc) return rc;
	CUDA_TRY(ctx, cudaMemcpy(ctx->bFlood.p, ctx->floodnum.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
	return CMGPU_OK;
}

inline int signbitsOf(const float *n) {
	return (n[0] < 0.0f ? 1 : 0) | (n[1] < 0.0f ? 2 : 0) | (n[2] < 0.0f ? 4 : 0);
}

// Everything the kernels index is range-checked here so that a malformed map is
// a CMGPU_ERR_BAD_MAP on the host instead of an out-of-bounds access on the GPU.
int validate(cmgpu_ctx *ctx, const cmgpu_flatmap *f, int *treeLevels) {
#define NEED(cond, ...) do { if (!(cond)) return fail(ctx, CMGPU_ERR_BAD_MAP, __VA_ARGS__); } while (0)
	NEED(f->numPlanes >= 12 && f->planes && f->planeType && f->planeSignbits, "map with no planes");
	NEED(f->numNodes >= 1 && f->nodes, "map has no nodes");
	NEED(f->numLeafs >= 1 && f->leafs, "map with no leafs");
	NEED(f->numBrushes >= 1 && f->brushes && f->brushBounds, "map with no brushes");
	NEED(f->numBrushSides >= 6 && f->brushsides, "map with no brush sides");
	NEED(f->numModels >= 1 && f->numModels <= CMGPU_MAX_SUBMODELS, "MAX_SUBMODELS exceeded");
	NEED(f->numLeafBrushes >= 1 && f->leafbrushes, "no leafbrushes");
	NEED(f->boxBrush >= 0 && f->boxBrush < f->numBrushes, "bad boxBrush");
	NEED(f->boxLeafBrush >= 0 && f->boxLeafBrush < f->numLeafBrushes, "bad boxLeafBrush");
	NEED(f->boxFirstPlane >= 0 && f->boxFirstPlane + 12 <= f->numPlanes, "bad boxFirstPlane");
	NEED(f->leafbrushes[f->boxLeafBrush] == f->boxBrush, "boxLeafBrush does not name boxBrush");
	NEED(f->brushes[3 * (size_t)f->boxBrush + 2] == 6, "box brush must have 6 sides");
	NEED(f->numClusters >= 0 && f->numAreas >= 0 && f->numAreas <= CMGPU_MAX_AREAS, "bad cluster / area counts");
	NEED(!f->vised || (f->visibility && f->clusterBytes > 0), "vised map without a visibility matrix");
	if (f->leafClusterArea) {
		for (int i = 0; i < f->numLeafs; i++) {
			const int c = f->leafClusterArea[2 * (size_t)i], a = f->leafClusterArea[2 * (size_t)i + 1];
			NEED(c >= -1 && (c == -1 || c < f->numClusters || !f->vised), "leaf %i: bad cluster %i", i, c);
			NEED(a >= -1 && (a == -1 || a < f->numAreas), "leaf %i: bad area %i", i, a);
			NEED(!f->vised || c < 8 * f->clusterBytes, "leaf %i: cluster %i beyond the visibility rows", i, c);
		}
	}

	for (int i = 0; i < f->numNodes; i++) {
		const int32_t *nd = f->nodes + 3 * (size_t)i;
		NEED(nd[0] >= 0 && nd[0] < f->numPlanes, "node %d: bad plane %d", i, nd[0]);
		for (int c = 1; c <= 2; c++) {
			if (nd[c] >= 0) NEED(nd[c] < f->numNodes, "node %d: bad child %d", i, nd[c]);
			else NEED(-1 - nd[c] < f->numLeafs, "node %d: bad leaf %d", i, -1 - nd[c]);
		}
	}
	{   // the node graph must be a tree: a cycle would never terminate on the device
		// ... and its depth bounds the traversal stacks: the pending far halves of a sweep (cm_trace.c:516-537) and the
		// pending children[1] of a box query (cm_test.c:131-156) each belong to a different node on the path from the
		// root, so a walk never holds more of them than the tree has levels.  The reference recurses without a limit.
		std::vector<unsigned char> seen((size_t)f->numNodes, 0);
		std::vector<std::pair<int, int>> st;
		st.push_back({0, 1});
		int levels = 0;
		while (!st.empty()) {
			const int n = st.back().first, d = st.back().second;
			st.pop_back();
			NEED(!seen[n], "node %d reachable twice: BSP is not a tree", n);
			seen[n] = 1;
			if (d > levels) levels = d;
			for (int c = 1; c <= 2; c++) {
				int ch = f->nodes[3 * (size_t)n + c];
				if (ch >= 0) st.push_back({ch, d + 1});
			}
		}
		if (levels > kDeepDepth - 1)
			return fail(ctx, CMGPU_ERR_LIMIT, "the map's BSP is %d levels deep; the traversal stacks hold %d pending halves", levels, kDeepDepth - 1);
		*treeLevels = levels;
	}
	auto checkLeaf = [&](const int32_t *lf, const char *what, int i) -> int {
		NEED(lf[1] >= 0 && lf[0] >= 0 && (long long)lf[0] + lf[1] <= f->numLeafBrushes, "%s %d: bad leafbrush range", what, i);
		NEED(lf[3] >= 0 && lf[2] >= 0 && (long long)lf[2] + lf[3] <= f->numLeafSurfaces, "%s %d: bad leafsurface range", what, i);
		return CMGPU_OK;
	};
	for (int i = 0; i < f->numLeafs; i++) {
		int rc = checkLeaf(f->leafs + 4 * (size_t)i, "leaf", i);
		if (rc) return rc;
	}
	if (f->numModels > 1) NEED(f->modelLeafs != nullptr, "modelLeafs missing");
	for (int i = 1; i < f->numModels; i++) {
		int rc = checkLeaf(f->modelLeafs + 4 * (size_t)i, "model", i);
		if (rc) return rc;
	}
	for (int i = 0; i < f->numLeafBrushes; i++)
		NEED(f->leafbrushes[i] >= 0 && f->leafbrushes[i] < f->numBrushes, "leafbrush %d out of range", i);
	for (int i = 0; i < f->numLeafSurfaces; i++)
		NEED(f->leafsurfaces[i] >= 0 && f->leafsurfaces[i] < f->numSurfaces, "leafsurface %d out of range", i);
	for (int i = 0; i < f->numBrushes; i++) {
		const int32_t *b = f->brushes + 3 * (size_t)i;
		NEED(b[2] >= 0 && b[1] >= 0 && (long long)b[1] + b[2] <= f->numBrushSides, "brush %d: bad side range", i);
	}
	for (int i = 0; i < f->numBrushSides; i++) {
		int pl = f->brushsides[2 * (size_t)i];
		NEED(pl >= 0 && pl < f->numPlanes, "brushside %d: bad plane", i);
		NEED(f->planeSignbits[pl] == signbitsOf(f->planes + 4 * (size_t)pl),
		     "plane %d: signbits do not follow the normal's signs", pl);
	}
	for (int i = 0; i < f->numSurfaces; i++)
		NEED(f->surf2patch[i] >= -1 && f->surf2patch[i] < f->numPatches, "surface %d: bad patch", i);
	for (int i = 0; i < f->numPatches; i++) {
		const int32_t *p = f->patches + 6 * (size_t)i;
		NEED(p[2] >= 0 && p[3] >= 0 && (long long)p[2] + p[3] <= f->numPatchPlanes, "patch %d: bad plane range", i);
		NEED(p[4] >= 0 && p[5] >= 0 && (long long)p[4] + p[5] <= f->numFacets, "patch %d: bad facet range", i);
		NEED(p[5] <= 65535, "patch %d: %d facets (the reference stops at MAX_FACETS = 1024, cm_patch.h:23; the kernels index a patch's facets with 16 bits)", i, p[5]);
		for (int k = 0; k < p[5]; k++) {
			const int32_t *fc = f->facets + 3 * (size_t)(p[4] + k);
			NEED(fc[0] >= 0 && fc[0] < p[3], "patch %d facet %d: bad surface plane", i, k);
			NEED(fc[1] >= 0 && fc[2] >= 0 && (long long)fc[2] + fc[1] <= f->numBorders, "patch %d facet %d: bad borders", i, k);
			NEED(fc[1] <= 26, "patch %d facet %d: %d borders (the reference's facet_t holds 4 + 6 + 16)", i, k, fc[1]);
			for (int j = 0; j < fc[1]; j++) {
				int bp = f->borders[2 * (size_t)(fc[2] + j)];
				NEED(bp >= 0 && bp < p[3], "patch %d facet %d border %d: bad plane", i, k, j);
			}
		}
	}
	for (int i = 0; i < f->numPatchPlanes; i++)
		NEED(f->patchPlaneSignbits[i] == signbitsOf(f->patchPlanes + 4 * (size_t)i),
		     "patch plane %d: signbits do not follow the normal's signs", i);
#undef NEED
	return CMGPU_OK;
}

inline float bitsToFloat(int32_t v) { float f; memcpy(&f, &v, 4); return f; }
inline int32_t floatToBits(float f) { int32_t v; memcpy(&v, &f, 4); return v; }

// true when sides 0..5 of brush i are exactly the planes -x,+x,-y,+y,-z,+z through its bounds
// (what CM_BoundBrush, cm_load.c:230-239, assumes); such sides are evaluated from the box alone.
bool canonicalAxialSides(const cmgpu_flatmap *f, int i) {
	const int32_t *b = f->brushes + 3 * (size_t)i;
	if (b[2] < 6 || i == f->boxBrush) return false;
	const float *bb = f->brushBounds + 6 * (size_t)i;
	for (int s = 0; s < 6; s++) {
		const float *pl = f->planes + 4 * (size_t)f->brushsides[2 * (size_t)(b[1] + s)];
		const int axis = s >> 1;
		const float want = (s & 1) ? 1.0f : -1.0f;
		for (int j = 0; j < 3; j++) {
			if (j == axis ? pl[j] != want : pl[j] != 0.0f) return false;
		}
		const float dist = (s & 1) ? bb[3 + axis] : -bb[axis];
		if (memcmp(&dist, &pl[3], 4) != 0) return false;
	}
	return true;
}

int packAndUpload(cmgpu_ctx *ctx, const cmgpu_flatmap *f) {
	DevMap &m = ctx->map;
	std::vector<int4> nodes((size_t)f->numNodes);
	std::vector<float4> nodePlanes((size_t)f->numNodes);
	for (int i = 0; i < f->numNodes; i++) {
		const int32_t *nd = f->nodes + 3 * (size_t)i;
		const float *pl = f->planes + 4 * (size_t)nd[0];
		int type = f->planeType[nd[0]] < 3 ? f->planeType[nd[0]] : 3;
		nodes[i] = make_int4(nd[1], nd[2], type, floatToBits(pl[3]));
		nodePlanes[i] = make_float4(pl[0], pl[1], pl[2], pl[3]);
	}
	std::vector<int4> leafs((size_t)f->numLeafs);
	for (int i = 0; i < f->numLeafs; i++) {
		const int32_t *l = f->leafs + 4 * (size_t)i;
		leafs[i] = make_int4(l[0], l[1], l[2], l[3]);
	}
	std::vector<int> leafsurfaces;
	if (f->numLeafSurfaces) leafsurfaces.assign(f->leafsurfaces, f->leafsurfaces + f->numLeafSurfaces);
	// every leafbrush entry carries the brush's box, contents and index, so a leaf's brush list is
	// one sequential stream of 32-byte records instead of index -> brush -> bounds pointer chasing
	std::vector<float4> leafBrushBox(2 * (size_t)f->numLeafBrushes);
	for (int k = 0; k < f->numLeafBrushes; k++) {
		const int bi = f->leafbrushes[k];
		const int32_t *b = f->brushes + 3 * (size_t)bi;
		const float *bb = f->brushBounds + 6 * (size_t)bi;
		leafBrushBox[2 * (size_t)k] = make_float4(bb[0], bb[1], bb[2], bitsToFloat(b[0]));
		leafBrushBox[2 * (size_t)k + 1] = make_float4(bb[3], bb[4], bb[5], bitsToFloat(bi));
	}
	std::vector<int4> brushInfo((size_t)f->numBrushes);
	for (int i = 0; i < f->numBrushes; i++) {
		const int32_t *b = f->brushes + 3 * (size_t)i;
		brushInfo[i] = make_int4(b[1], b[2], b[0], canonicalAxialSides(f, i) ? 1 : 0);
	}
	std::vector<float4> sidePlanes((size_t)f->numBrushSides);
	std::vector<int2> sideMeta((size_t)f->numBrushSides);
	for (int i = 0; i < f->numBrushSides; i++) {
		int pl = f->brushsides[2 * (size_t)i];
		const float *p = f->planes + 4 * (size_t)pl;
		sidePlanes[i] = make_float4(p[0], p[1], p[2], p[3]);
		sideMeta[i] = make_int2(f->brushsides[2 * (size_t)i + 1], (int)f->planeType[pl] | ((int)f->planeSignbits[pl] << 8));
	}
	ctx->hostSides.assign((size_t)f->numBrushSides + 1, hostrec::SideRec{});
	for (int i = 0; i < f->numBrushSides; i++) {
		hostrec::SideRec &r = ctx->hostSides[(size_t)i + 1];
		r.plane[0] = sidePlanes[i].x; r.plane[1] = sidePlanes[i].y; r.plane[2] = sidePlanes[i].z; r.plane[3] = sidePlanes[i].w;
		r.typeSignbits = (uint32_t)sideMeta[i].y & 0xffffu;
		r.surfaceFlags = (uint32_t)sideMeta[i].x;
	}
	ctx->hostBrushContents.assign((size_t)f->numBrushes + 1, 0);
	for (int i = 0; i < f->numBrushes; i++) ctx->hostBrushContents[(size_t)i + 1] = f->brushes[3 * (size_t)i];
	std::vector<int4> modelLeafs((size_t)f->numModels);
	for (int i = 0; i < f->numModels; i++) {
		if (i == 0 || !f->modelLeafs) { modelLeafs[i] = make_int4(0, 0, 0, 0); continue; }
		const int32_t *l = f->modelLeafs + 4 * (size_t)i;
		modelLeafs[i] = make_int4(l[0], l[1], l[2], l[3]);
	}
	std::vector<int> surf2patch;
	if (f->numSurfaces) surf2patch.assign(f->surf2patch, f->surf2patch + f->numSurfaces);
	std::vector<int4> patchInfo(2 * (size_t)f->numPatches);
	std::vector<float4> patchBox(2 * (size_t)f->numPatches);
	for (int i = 0; i < f->numPatches; i++) {
		const int32_t *p = f->patches + 6 * (size_t)i;
		const float *pb = f->patchBounds + 6 * (size_t)i;
		patchInfo[2 * (size_t)i] = make_int4(p[0], p[1], p[2], p[3]);
		patchInfo[2 * (size_t)i + 1] = make_int4(p[4], p[5], 0, 0);     // .z: first chunk of 32 facets, set below
		patchBox[2 * (size_t)i] = make_float4(pb[0], pb[1], pb[2], 0.0f);
		patchBox[2 * (size_t)i + 1] = make_float4(pb[3], pb[4], pb[5], 0.0f);
	}
	std::vector<float4> patchPlanes((size_t)f->numPatchPlanes);
	for (int i = 0; i < f->numPatchPlanes; i++) {
		const float *p = f->patchPlanes + 4 * (size_t)i;
		patchPlanes[i] = make_float4(p[0], p[1], p[2], p[3]);
	}
	std::vector<int4> facets((size_t)f->numFacets);
	for (int i = 0; i < f->numFacets; i++) {
		const int32_t *fc = f->facets + 3 * (size_t)i;
		facets[i] = make_int4(fc[0], fc[1], fc[2], 0);
	}
	std::vector<int> borders((size_t)f->numBorders);
	for (int i = 0; i < f->numBorders; i++) {
		borders[i] = f->borders[2 * (size_t)i] | (f->borders[2 * (size_t)i + 1] ? (int)0x80000000u : 0);
	}
	// Sweeps walk a facet's planes one per step: every facet gets its planes as one contiguous run (surface plane,
	// then the borders in order) so that a step is ONE dependent load instead of facet -> border -> plane.
	// facetRuns[f] = first plane of the run, numBorders, borderInward bits (validate() caps numBorders at 26), 0
	std::vector<int4> facetRuns((size_t)f->numFacets);
	std::vector<float4> facetPlanes;
	for (int p = 0; p < f->numPatches; p++) {
		const int32_t *pt = f->patches + 6 * (size_t)p;
		for (int k = 0; k < pt[5]; k++) {
			const int fi = pt[4] + k;
			const int32_t *fc = f->facets + 3 * (size_t)fi;
			unsigned inward = 0;
			const int start = (int)facetPlanes.size();
			const float *sp = f->patchPlanes + 4 * (size_t)(pt[2] + fc[0]);
			facetPlanes.push_back(make_float4(sp[0], sp[1], sp[2], sp[3]));
			for (int j = 0; j < fc[1]; j++) {
				const int32_t *bd = f->borders + 2 * (size_t)(fc[2] + j);
				const float *bp = f->patchPlanes + 4 * (size_t)(pt[2] + bd[0]);
				facetPlanes.push_back(make_float4(bp[0], bp[1], bp[2], bp[3]));
				if (bd[1]) inward |= 1u << j;
			}
			facetRuns[fi] = make_int4(start, fc[1], (int)inward, 0);
		}
	}

	// the axial borders of every facet, and their maxima over the chunks of 32 facets a warp takes per pass (DevMap::facetAxial)
	std::vector<float4> facetAxial(2 * (size_t)f->numFacets), facetChunkAxial;
	for (int p = 0; p < f->numPatches; p++) {
		const int32_t *pt = f->patches + 6 * (size_t)p;
		patchInfo[2 * (size_t)p + 1].z = (int)(facetChunkAxial.size() / 2);
		for (int k0 = 0; k0 < pt[5]; k0 += 32) {
			float cmax[6];
			for (int d = 0; d < 6; d++) cmax[d] = -INFINITY;
			for (int k = k0; k < pt[5] && k < k0 + 32; k++) {
				const int fi = pt[4] + k;
				const int32_t *fc = f->facets + 3 * (size_t)fi;
				float D[6];
				for (int d = 0; d < 6; d++) D[d] = INFINITY;
				for (int j = 0; j < fc[1]; j++) {
					const int32_t *bd = f->borders + 2 * (size_t)(fc[2] + j);
					const float *bp = f->patchPlanes + 4 * (size_t)(pt[2] + bd[0]);
					float n[3] = {bp[0], bp[1], bp[2]}, w = bp[3];
					if (bd[1]) { n[0] = -n[0]; n[1] = -n[1]; n[2] = -n[2]; w = -w; }      // borderInward: the plane is flipped (cm_patch.c:1413-1417)
					for (int a = 0; a < 3; a++) {
						const int b = (a + 1) % 3, c = (a + 2) % 3;
						if (n[b] == 0.0f && n[c] == 0.0f && (n[a] == 1.0f || n[a] == -1.0f)) {
							const int d = 2 * a + (n[a] < 0.0f ? 1 : 0);
							if (w < D[d]) D[d] = w;
						}
					}
				}
				facetAxial[2 * (size_t)fi] = make_float4(D[0], D[1], D[2], D[3]);
				facetAxial[2 * (size_t)fi + 1] = make_float4(D[4], D[5], 0.0f, 0.0f);
				for (int d = 0; d < 6; d++) if (D[d] > cmax[d]) cmax[d] = D[d];
			}
			facetChunkAxial.push_back(make_float4(cmax[0], cmax[1], cmax[2], cmax[3]));
			facetChunkAxial.push_back(make_float4(cmax[4], cmax[5], 0.0f, 0.0f));
		}
	}

	int rc = CMGPU_OK;
#define UP(vec, field) if ((rc = uploadArray(ctx, vec, &m.field)) != CMGPU_OK) return rc
	UP(nodes, nodes); UP(nodePlanes, nodePlanes); UP(leafs, leafs); UP(leafBrushBox, leafBrushBox);
	UP(leafsurfaces, leafsurfaces); UP(brushInfo, brushInfo);
	UP(sidePlanes, sidePlanes); UP(sideMeta, sideMeta); UP(modelLeafs, modelLeafs); UP(surf2patch, surf2patch);
	UP(patchInfo, patchInfo); UP(patchBox, patchBox); UP(patchPlanes, patchPlanes); UP(facets, facets);
	UP(borders, borders); UP(facetRuns, facetRuns); UP(facetPlanes, facetPlanes);
	UP(facetAxial, facetAxial); UP(facetChunkAxial, facetChunkAxial);
	ctx->leafCluster = nullptr;
	if (f->leafClusterArea) {
		std::vector<int> lca(f->leafClusterArea, f->leafClusterArea + 2 * (size_t)f->numLeafs);
		if ((rc = uploadArray(ctx, lca, &ctx->leafCluster)) != CMGPU_OK) return rc;
	}
	{	// ---- layout of the hot kernel: leafs as sentinel-terminated streams of widened brush boxes ----
		std::vector<int> streamStart((size_t)f->numLeafs);
		std::vector<float4> stream;
		stream.reserve(2 * ((size_t)f->numLeafBrushes + (size_t)f->numLeafs));
		for (int l = 0; l < f->numLeafs; l++) {
			const int32_t *lf = f->leafs + 4 * (size_t)l;
			streamStart[l] = (int)(stream.size() / 2);
			for (int k = 0; k < lf[1]; k++) {
				const int bi = f->leafbrushes[lf[0] + k];
				const float *bb = f->brushBounds + 6 * (size_t)bi;
				// CM_BoundsIntersect's own float ops (cm_test.c:481-494), done once
				stream.push_back(make_float4(bb[0] - kBoundsEps, bb[1] - kBoundsEps, bb[2] - kBoundsEps,
				                             bitsToFloat(f->brushes[3 * (size_t)bi])));
				stream.push_back(make_float4(bb[3] + kBoundsEps, bb[4] + kBoundsEps, bb[5] + kBoundsEps, bitsToFloat(bi)));
			}
			stream.push_back(make_float4(0.0f, 0.0f, 0.0f, bitsToFloat(0)));
			stream.push_back(make_float4(0.0f, 0.0f, 0.0f, bitsToFloat(-1)));
		}
		// spare sentinels, so that a kernel may always read a few entries past the one it is looking at
		for (int spare = 0; spare < 7; spare++) {
			stream.push_back(make_float4(0.0f, 0.0f, 0.0f, bitsToFloat(0)));
			stream.push_back(make_float4(0.0f, 0.0f, 0.0f, bitsToFloat(-1)));
		}
		std::vector<int4> fnodes((size_t)f->numNodes);
		for (int i = 0; i < f->numNodes; i++) {
			int4 nd = nodes[i];
			if (nd.x < 0) nd.x = -1 - streamStart[-1 - nd.x];
			if (nd.y < 0) nd.y = -1 - streamStart[-1 - nd.y];
			fnodes[i] = nd;
		}
		// the first kEntryLevels levels of the tree in heap order (EntryGrid::top): entry (1 << L) - 1 + path is where the L
		// decisions `path` (first decision = most significant bit, 1 = children[0]) lead; positions below a leaf stay 0
		std::vector<int> top(((size_t)2 << kEntryLevels), 0);
		{
			struct T { int ref, level; unsigned path; };
			std::vector<T> st;
			st.push_back({0, 0, 0u});
			while (!st.empty()) {
				const T t = st.back();
				st.pop_back();
				top[((size_t)1 << t.level) - 1 + t.path] = t.ref;
				if (t.ref < 0 || t.level == kEntryLevels) continue;
				st.push_back({fnodes[t.ref].x, t.level + 1, (t.path << 1) | 1u});
				st.push_back({fnodes[t.ref].y, t.level + 1, t.path << 1});
			}
		}
		if ((rc = uploadArray(ctx, top, &ctx->entry.top)) != CMGPU_OK) return rc;
		std::vector<float4> brushRec(2 * (size_t)f->numBrushes);
		std::vector<int> brushContents((size_t)f->numBrushes);
		for (int i = 0; i < f->numBrushes; i++) {
			const int32_t *b = f->brushes + 3 * (size_t)i;
			const float *bb = f->brushBounds + 6 * (size_t)i;
			brushRec[2 * (size_t)i] = make_float4(bb[0], bb[1], bb[2], bitsToFloat(b[1]));
			brushRec[2 * (size_t)i + 1] = make_float4(bb[3], bb[4], bb[5], bitsToFloat((b[2] & 0xffffff) | (brushInfo[i].w ? (1 << 30) : 0)));
			brushContents[i] = b[0];
		}
		FastMap &fm = ctx->fast;
		if ((rc = uploadArray(ctx, fnodes, &fm.nodes)) != CMGPU_OK) return rc;
		if ((rc = uploadArray(ctx, stream, &fm.stream)) != CMGPU_OK) return rc;
		if ((rc = uploadArray(ctx, brushRec, &fm.brushRec)) != CMGPU_OK) return rc;
		if ((rc = uploadArray(ctx, brushContents, &fm.brushContents)) != CMGPU_OK) return rc;
		ctx->numPatches = f->numPatches;
	}
#undef UP
	{	// entry grid (EntryGrid): cells of about entryCell units over the bounds of the world's brushes, at most 2^20 of them
		float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
		for (int i = 0; i < f->numBrushes; i++) {
			if (i == f->boxBrush) continue;
			const float *bb = f->brushBounds + 6 * (size_t)i;
			for (int a = 0; a < 3; a++) {
				if (bb[a] < lo[a]) lo[a] = bb[a];
				if (bb[3 + a] > hi[a]) hi[a] = bb[3 + a];
			}
		}
		EntryGrid &e = ctx->entry;
		int bits[3], total = 0;
		for (int a = 0; a < 3; a++) {
			float ext = hi[a] - lo[a];
			if (!(ext > 1.0f) || !(ext < 1e9f)) ext = 1.0f;
			bits[a] = 0;
			while (bits[a] < 10 && ext / (float)(1 << bits[a]) > ctx->entryCell) bits[a]++;
			total += bits[a];
		}
		while (total > 20) {
			const int a = bits[0] >= bits[1] ? (bits[0] >= bits[2] ? 0 : 2) : (bits[1] >= bits[2] ? 1 : 2);
			bits[a]--;
			total--;
		}
		for (int a = 0; a < 3; a++) {
			float ext = hi[a] - lo[a];
			if (!(ext > 1.0f) || !(ext < 1e9f)) ext = 1.0f;
			e.cells[a] = 1 << bits[a];
			e.lo[a] = lo[a];
			e.inv[a] = (float)e.cells[a] / ext;
		}
		e.cellRec = nullptr;
	}
	// room behind the map for the per-hull tables (node thresholds + entry grid), so that the map's L2 window covers them too
	ctx->hullArenaPer = (((size_t)f->numNodes * sizeof(int4) + 16 + 255) & ~(size_t)255) +
	                    (((size_t)ctx->entry.cells[0] * ctx->entry.cells[1] * ctx->entry.cells[2] * sizeof(uint2) + 255) & ~(size_t)255);
	if ((rc = commitMapSlab(ctx)) != CMGPU_OK) return rc;
	ctx->fast.nodePlanes = m.nodePlanes;
	ctx->fast.sidePlanes = m.sidePlanes;
	ctx->fast.sideMeta = m.sideMeta;
	{	// binning grid over the bounds of the world's brushes
		float lo[3] = {1e30f, 1e30f, 1e30f}, hi[3] = {-1e30f, -1e30f, -1e30f};
		for (int i = 0; i < f->numBrushes; i++) {
			if (i == f->boxBrush) continue;
			const float *bb = f->brushBounds + 6 * (size_t)i;
			for (int a = 0; a < 3; a++) {
				if (bb[a] < lo[a]) lo[a] = bb[a];
				if (bb[3 + a] > hi[a]) hi[a] = bb[3 + a];
			}
		}
		ctx->slabOK = true;
		for (int a = 0; a < 3; a++) if (!(lo[a] > -32768.0f && hi[a] < 32768.0f)) ctx->slabOK = false;
		BinGrid &g = ctx->grid;
		int totalBits = 0;
		for (int a = 0; a < 3; a++) {
			float ext = hi[a] - lo[a];
			if (!(ext > 1.0f) || !(ext < 1e9f)) { ext = 1.0f; lo[a] = 0.0f; }
			int bits = 0;
			while (bits < 8 && (float)(1 << bits) * ctx->binCell < ext) bits++;
			g.bits[a] = bits;
			totalBits += bits;
		}
		while (totalBits > 19) {                     // keep the histogram at a few MB
			int a = g.bits[0] >= g.bits[1] ? (g.bits[0] >= g.bits[2] ? 0 : 2) : (g.bits[1] >= g.bits[2] ? 1 : 2);
			g.bits[a]--;
			totalBits--;
		}
		for (int a = 0; a < 3; a++) {
			float ext = hi[a] - lo[a];
			if (!(ext > 1.0f) || !(ext < 1e9f)) ext = 1.0f;
			g.cells[a] = 1 << g.bits[a];
			g.lo[a] = lo[a];
			g.inv[a] = (float)g.cells[a] / ext;
		}
		g.longLen2 = ctx->binLongLen * ctx->binLongLen;
		g.dirBins = ctx->binDir ? 8 : 1;
		g.numBins = 2 * kBinClasses * g.cells[0] * g.cells[1] * g.cells[2] * g.dirBins;   // x2: point hulls / box hulls of per-item batches
	}
	m.numNodes = f->numNodes;
	m.numModels = f->numModels;
	m.boxBrush = f->boxBrush;
	m.boxLeafBrush = f->boxLeafBrush;
	m.noCurves = ctx->noCurves;
	m.playerCurveClip = ctx->playerCurveClip;
	ctx->numModels = f->numModels;
	// visibility matrix and areas; every portal starts closed (cm_load.c:320,734)
	ctx->vised = f->vised != 0;
	ctx->numClusters = f->numClusters;
	ctx->clusterBytes = f->clusterBytes;
	ctx->numAreas = f->leafClusterArea ? f->numAreas : 0;
	if (ctx->vised) {
		const size_t bytes = (size_t)f->numClusters * (size_t)f->clusterBytes;
		if ((rc = reserve(ctx, ctx->bVis, bytes + 1))) return rc;
		CUDA_TRY(ctx, cudaMemcpy(ctx->bVis.p, f->visibility, bytes, cudaMemcpyHostToDevice));
	}
	ctx->areaPortals.assign((size_t)ctx->numAreas * ctx->numAreas, 0);
	if ((rc = floodAreas(ctx))) return rc;
	ctx->hasMap = true;
	return CMGPU_OK;
}

// CM_ClipHandleToModel (cm_load.c:768-798): a handle is a loaded inline model or 1023
int checkHandles(cmgpu_ctx *ctx, int n, const int32_t *models) {
	if (!models) return CMGPU_OK;
	for (int i = 0; i < n; i++) {
		int h = models[i];
		if (h == CMGPU_BOX_MODEL_HANDLE) continue;
		if (h < 0 || h >= ctx->numModels) return fail(ctx, CMGPU_ERR_BAD_HANDLE, "CM_ClipHandleToModel: bad handle %i (item %d)", h, i);
	}
	return CMGPU_OK;
}

// Per-batch constants of the hot kernel.  The float ops are CM_Trace's own (cm_trace.c:582-588, 650-659).
// nodeHi/nodeLo: CM_TraceThroughTree compares the float differences t1,t2 (widened) with the doubles
// offset+1 and -offset-1 (cm_trace.c:483,487).  For a float t and a double x: t >= x <=> t >= RU(x) and
// t < -x <=> t < -RD(x), RU/RD = x rounded up/down to float; so the comparison can stay in float.
FastHull makeFastHull(const float *mins, const float *maxs, int mask, bool slabOK) {
	FastHull h{};
	bool point = true;
	for (int a = 0; a < 3; a++) {
		const float off = (mins[a] + maxs[a]) * 0.5f;
		h.off[a] = off;
		h.n0[a] = mins[a] - off;
		h.p0[a] = maxs[a] - off;
		if (h.n0[a] != 0.0f) point = false;
	}
	h.isPoint = point ? 1 : 0;
	for (int a = 0; a < 3; a++) {
		const double x = (double)(point ? 0.0f : h.p0[a]) + 1.0;
		float up = (float)x, dn = (float)x;
		if ((double)up < x) up = nextafterf(up, INFINITY);
		if ((double)dn > x) dn = nextafterf(dn, -INFINITY);
		h.nodeHi[a] = up;
		h.nodeLo[a] = -dn;
	}
	h.mask = mask;
	h.slabOK = slabOK ? 1 : 0;
	return h;
}

// The node table of a hull (cm_trace_pool.cuh, nodeThresholdKernel): looked up by the bits of the six thresholds, built
// on first use.  Building synchronises the stream once per (map, hull); a table whose boundaries could not be verified
// on the device is marked unusable and the launch takes the plain node step.
const cmgpu_ctx::HullTable *hullTable(cmgpu_ctx *ctx, const FastHull &h, cudaStream_t s) {
	NodeHull key;
	for (int a = 0; a < 3; a++) { key.hi[a] = h.nodeHi[a]; key.lo[a] = h.nodeLo[a]; }
	for (auto &t : ctx->hullTables) {
		if (!memcmp(&t.key, &key, sizeof(key))) return t.usable ? &t : nullptr;
	}
	if (ctx->capturing) return nullptr;                   // building synchronises: not inside a capture (plain node step instead)
	if ((int)ctx->hullTables.size() >= kHullTables) {
		// every slot behind the map is taken.  Start over -- unless a recorded graph may name a slot: then this hull simply
		// gets no table (the plain node step, same results)
		if (ctx->graphPinned) return nullptr;
		cudaDeviceSynchronize();                          // nothing may still read the old tables
		ctx->hullTables.clear();
	}
	cmgpu_ctx::HullTable t{};
	t.key = key;
	t.usable = false;
	t.gridUsable = false;
	const int n = ctx->map.numNodes;
	const size_t bytes = (size_t)n * sizeof(int4);
	const size_t nodesRoom = (bytes + 16 + 255) & ~(size_t)255;
	char *slot = (char *)ctx->mapSlab + ctx->hullArenaOff + ctx->hullTables.size() * ctx->hullArenaPer;
	if (ctx->mapSlab && ctx->hullArenaPer >= nodesRoom) {
		t.nodes = (int4 *)slot;
		t.cells = (uint2 *)(slot + nodesRoom);
		int *flag = (int *)(slot + bytes);
		int hostFlag = -1;
		bool ok = cudaMemsetAsync(flag, 0, sizeof(int), s) == cudaSuccess;
		if (ok) nodeThresholdKernel<<<(n + 255) / 256, 256, 0, s>>>(ctx->fast.nodes, n, key, t.nodes, flag);
		ok = ok && cudaGetLastError() == cudaSuccess;
		ok = ok && cudaMemcpyAsync(&hostFlag, flag, sizeof(int), cudaMemcpyDeviceToHost, s) == cudaSuccess;
		ok = ok && cudaStreamSynchronize(s) == cudaSuccess;
		t.usable = ok && hostFlag == 0;
		ctx->launches++;
		// the entry grid of this hull (entryGridKernel): built from the table just made, verified on the device like it
		const EntryGrid &e = ctx->entry;
		const int cells = e.cells[0] * e.cells[1] * e.cells[2];
		if (t.usable && ctx->entryGrid && e.top && cells > 1 && ctx->hullArenaPer >= nodesRoom + (size_t)cells * sizeof(uint2)) {
			hostFlag = -1;
			ok = cudaMemsetAsync(flag, 0, sizeof(int), s) == cudaSuccess;
			if (ok) entryGridKernel<<<(cells + 255) / 256, 256, 0, s>>>(t.nodes, e, t.cells, flag);
			ok = ok && cudaGetLastError() == cudaSuccess;
			ok = ok && cudaMemcpyAsync(&hostFlag, flag, sizeof(int), cudaMemcpyDeviceToHost, s) == cudaSuccess;
			ok = ok && cudaStreamSynchronize(s) == cudaSuccess;
			t.gridUsable = ok && hostFlag == 0;
			ctx->launches++;
		}
	}
	ctx->hullTables.push_back(t);
	return t.usable ? &ctx->hullTables.back() : nullptr;
}

// Lanes per item of the thread-per-item kernel (log2): the largest spread that still fits the launch into the GPU's
// resident warps at once.
static int simpleSpread(const cmgpu_ctx *ctx, const int n) {
	if (ctx->simpleSpread >= 0) return ctx->simpleSpread;
	const long long slots = (long long)ctx->gridSimple * kBlock;      // resident threads
	int lg = 0;
	while (lg < 5 && ((long long)n << (lg + 1)) <= slots) lg++;
	return lg;
}

// Device-pointer calls share the context's workspaces.  Whatever such a call leaves running on its stream, the next
// one on ANOTHER stream (or a host-pointer call) must wait for: extDone is recorded after the last kernel that reads or
// writes a shared workspace, and waited for at entry.
int extWait(cmgpu_ctx *ctx, cudaStream_t s) {
	if (ctx->extBusy && ctx->extStream != s) CUDA_TRY(ctx, cudaStreamWaitEvent(s, ctx->extDone, 0));
	return CMGPU_OK;
}
int extRecord(cmgpu_ctx *ctx, cudaStream_t s) {
	cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
	if (cudaStreamIsCapturing(s, &st) != cudaSuccess) cudaGetLastError();
	if (st != cudaStreamCaptureStatusNone) return CMGPU_OK;          // captured work: the caller orders replays
	CUDA_TRY(ctx, cudaEventRecord(ctx->extDone, s));
	ctx->extStream = s;
	ctx->extBusy = true;
	return CMGPU_OK;
}

// wsOffset: first item of this launch inside the context's binning workspace (the host pipeline gives
// every chunk its own range); histSlot: which histogram table to use (one per pipeline stream).
// hybrid (traceHostHybrid): the pool kernel whatever the size of the launch, leaving the walk's 16-byte results in
// bCompact[wsOffset + item]; 1 = and nothing more (the host writes the records), 2 = followed by expandRecordsKernel.
int launchTrace(cmgpu_ctx *ctx, TraceBatch q, cudaStream_t s, size_t wsOffset = 0, int histSlot = -1, bool sideExpand = false, int hybrid = 0) {
	const bool compactOnly = hybrid == 1;
	if (q.n <= 0) return CMGPU_OK;
	const unsigned *slotOf = nullptr;
	q.compact = nullptr;
	const int arenaSlot = histSlot < 0 ? kBinSlots : histSlot;
	const bool external = histSlot < 0;              // device-pointer call on the caller's stream
	// A call on a stream that is being captured (CUDA graph of a rollout tick, a server frame) only records work:
	// no allocation, no event bookkeeping with the host-pointer calls, no timing.  The captured launches use the
	// workspaces as they are, so the caller keeps graph replays and other calls on this context in stream order.
	struct CaptureScope { cmgpu_ctx *c; ~CaptureScope() { c->capturing = false; } } captureScope{ctx};
	{
		cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
		if (external && cudaStreamIsCapturing(s, &st) == cudaSuccess) ctx->capturing = st != cudaStreamCaptureStatusNone;
		else cudaGetLastError();
		if (ctx->capturing) ctx->graphPinned = true;
	}
	const bool capturing = ctx->capturing;
	if (external && !capturing && ctx->extBusy && ctx->extStream != s) CUDA_TRY(ctx, cudaStreamWaitEvent(s, ctx->extDone, 0));
	ctx->map.noCurves = ctx->noCurves;
	ctx->map.playerCurveClip = ctx->playerCurveClip;
	if (ctx->binning && q.n >= ctx->binMinItems) {
		int rc;
		if (external) {
			histSlot = 0;
			wsOffset = 0;
			if ((rc = reserve(ctx, ctx->bKeys, (size_t)q.n * 4))) return rc;
			if ((rc = reserve(ctx, ctx->bRecs, (size_t)q.n * 32))) return rc;
			if ((rc = reserve(ctx, ctx->bCompact, (size_t)q.n * 16))) return rc;
		}
		const int numTiles = (ctx->grid.numBins + kScanTile - 1) / kScanTile;
		if ((rc = reserve(ctx, ctx->bHist[histSlot], ((size_t)ctx->grid.numBins + (size_t)numTiles) * 4))) return rc;
		unsigned *keys = (unsigned *)ctx->bKeys.p + wsOffset;
		float4 *recs = (float4 *)ctx->bRecs.p + 2 * wsOffset;
		unsigned *hist = (unsigned *)ctx->bHist[histSlot].p;
		CUDA_TRY(ctx, cudaMemsetAsync(hist, 0, (size_t)ctx->grid.numBins * 4, s));
		const int g256 = (q.n + 255) / 256;
		binCountKernel<<<g256, 256, 0, s>>>(ctx->grid, q.n, q.starts, q.ends, q.models, q.mins, q.maxs, keys, hist);
		unsigned *tileSums = hist + ctx->grid.numBins;
		binTileSumKernel<<<numTiles, 256, 0, s>>>(hist, ctx->grid.numBins, tileSums);
		binTileScanKernel<<<1, 256, 0, s>>>(tileSums, numTiles);
		binScanKernel<<<numTiles, 256, 0, s>>>(hist, ctx->grid.numBins, tileSums);
		binScatterKernel<<<g256, 256, 0, s>>>(q.n, q.starts, q.ends, keys, hist, recs);
		ctx->launches += 5;
		q.recs = recs;
		slotOf = keys;
	}
	// persistent kernel: as many CTAs as fit on the GPU (fewer for tiny batches), each warp pulls
	// kChunk items at a time from a cursor that must be zero at launch
	q.cursor = ctx->dCursor + (ctx->cursorSlot++ % kCursorSlots);
	q.gang = ctx->gang;
	q.gangFetch = ctx->gangFetch;
	q.gangBrush = ctx->gangBrush;
	q.lightReps = ctx->lightReps;
	q.facetSteps = ctx->facetSteps;
	q.stackDepth = ctx->treeDepth;
	const bool deep = ctx->treeDepth > kMaxDepth;   // the kernels with local-memory stacks have a variant for such maps
	CUDA_TRY(ctx, cudaMemsetAsync(q.cursor, 0, sizeof(int), s));
	// the hot kernel serves world traces with one hull and one mask on maps without (active) patches
	const bool fast = ctx->useFast && !q.mins && !q.masks && !q.models && !q.origins && (ctx->numPatches == 0 || ctx->noCurves);
	// the pool variant pays off once every warp has several rounds of items to work through
	const bool pool = fast && ctx->usePool && (hybrid || q.n >= (external ? ctx->poolMinItems : ctx->pipePoolMinItems));
	// out must be 16-byte aligned for the expand pass's stores (any cudaMalloc'ed or record-aligned pointer is)
	// (not in the host pipeline: its pieces are small, their records leave over PCIe anyway, and the extra launch per
	//  piece costs the call 8 % -- measured 764 against 826 Mtraces/s end to end)
	// and not for the smaller launches of the one-item-per-lane kernels either (a rollout phase, a game tick: one more
	// launch is 5 % of their latency; "two_phase_small" turns it on for them)
	if (fast && (pool || ctx->twoPhaseSmall) && slotOf && ctx->twoPhase && (external || ctx->twoPhaseHost || sideExpand || hybrid) && !ctx->counting && ((uintptr_t)q.out & 15) == 0)
		q.compact = (uint4 *)ctx->bCompact.p + wsOffset;
	q.compactByItem = q.compact && pool && ctx->compactByItem ? 1 : 0;
	if (hybrid && !q.compactByItem) return fail(ctx, CMGPU_ERR_INVALID, "internal: this launch cannot leave its results as 16-byte records by item");
	const int resident = pool ? (ctx->counting ? ctx->gridPoolCount : ctx->gridPool)
	                   : fast ? (ctx->counting ? ctx->gridFastCount : ctx->gridFast) : (ctx->counting ? ctx->gridCount : ctx->gridPlain);
	// Items per cursor grab: kChunk for big batches; for small ones a share that spreads the batch over all
	// resident warps -- down to 8 items per warp: a batch's latency is the walk of its slowest warp, and a warp
	// with fewer items in flight runs fewer kinds of step per round.
	const long long residentWarps = (long long)resident * (kBlock / 32);
	int chunk = (int)(((long long)q.n / (residentWarps * (pool ? CMGPU_POOL_K : 1) * 4)) / 32 * 32);
	if (chunk < 32) {
		chunk = (int)(((long long)q.n + residentWarps - 1) / residentWarps);
		if (chunk < ctx->minChunk) chunk = ctx->minChunk;
		if (chunk > 32) chunk = 32;
	}
	if (chunk > kChunk) chunk = kChunk;
	q.chunk = chunk;
	const long long warpsWanted = ((long long)q.n + chunk - 1) / chunk;
	long long blocks = (warpsWanted + (kBlock / 32) - 1) / (kBlock / 32);
	if (blocks > resident) blocks = resident;
	// pieces of a host batch: a launch takes only a share of the GPU, so that the next piece's walk (another stream) runs
	// next to it and fills the SMs while this one's last warps finish their items
	if (hybrid && ctx->hybridGridEighths < 8 && blocks > (long long)resident * ctx->hybridGridEighths / 8) blocks = (long long)resident * ctx->hybridGridEighths / 8;
	if (blocks < 1) blocks = 1;
	cudaEvent_t tEnd = nullptr;
	if (ctx->timing && !capturing) {
		if (ctx->timingUsed == ctx->timingPairs.size()) {
			cudaEvent_t a, b;
			CUDA_TRY(ctx, cudaEventCreate(&a));
			CUDA_TRY(ctx, cudaEventCreate(&b));
			ctx->timingPairs.push_back({a, b});
		}
		auto &pr = ctx->timingPairs[ctx->timingUsed++];
		tEnd = pr.second;
		CUDA_TRY(ctx, cudaEventRecord(pr.first, s));
	}
	if (fast && !pool && q.n <= ctx->simpleMaxItems) {
		const FastHull h = makeFastHull(q.sharedMins, q.sharedMaxs, q.sharedMask, ctx->slabOK && ctx->slabTest);
		q.spreadLog2 = simpleSpread(ctx, q.n);
		const int nb = (int)((((long long)q.n << q.spreadLog2) + kBlock - 1) / kBlock);
		if (deep)               CUDA_TRY(ctx, launchOnMap(ctx, traceSimpleKernel<false, kDeepDepth>, nb, kBlock, 0, s, ctx->fast, q, h));
		else if (ctx->counting) CUDA_TRY(ctx, launchOnMap(ctx, traceSimpleKernel<true>, nb, kBlock, 0, s, ctx->fast, q, h));
		else                    CUDA_TRY(ctx, launchOnMap(ctx, traceSimpleKernel<false>, nb, kBlock, 0, s, ctx->fast, q, h));
	} else if (pool) {
		q.lightReps = ctx->poolReps;
		q.gangFetch = ctx->gangFetchPool;
		q.gang = ctx->gangPool;
		int rcA = reserve(ctx, ctx->bStackArena[arenaSlot], (size_t)resident * kBlock * CMGPU_POOL_K * q.stackDepth * sizeof(FastSeg));
		if (rcA) return rcA;
		q.stackArena = (FastSeg *)ctx->bStackArena[arenaSlot].p;
		rcA = reserve(ctx, ctx->bPoolHdr[arenaSlot], (size_t)resident * kBlock * CMGPU_POOL_K * sizeof(PoolHdr));
		if (rcA) return rcA;
		q.poolHdr = ctx->bPoolHdr[arenaSlot].p;
		const FastHull h = makeFastHull(q.sharedMins, q.sharedMaxs, q.sharedMask, ctx->slabOK && ctx->slabTest);
		const cmgpu_ctx::HullTable *ht = (ctx->nodeThresholds && !ctx->counting) ? hullTable(ctx, h, s) : nullptr;
		q.tnodes = ht ? ht->nodes : nullptr;
		q.entry = ctx->entry;
		q.entry.cellRec = (ht && ht->gridUsable && ctx->entryGrid) ? ht->cells : nullptr;
		if (ctx->counting)  CUDA_TRY(ctx, launchOnMap(ctx, tracePoolKernel<true, CMGPU_POOL_K, false>, (int)blocks, kBlock, kPoolSmem, s, ctx->fast, q, h));
		else if (q.tnodes)  CUDA_TRY(ctx, launchOnMap(ctx, tracePoolKernel<false, CMGPU_POOL_K, true>, (int)blocks, kBlock, kPoolSmem, s, ctx->fast, q, h));
		else                CUDA_TRY(ctx, launchOnMap(ctx, tracePoolKernel<false, CMGPU_POOL_K, false>, (int)blocks, kBlock, kPoolSmem, s, ctx->fast, q, h));
	} else if (fast) {
		int rcA = reserve(ctx, ctx->bStackArena[arenaSlot], (size_t)resident * kBlock * q.stackDepth * sizeof(FastSeg));
		if (rcA) return rcA;
		q.stackArena = (FastSeg *)ctx->bStackArena[arenaSlot].p;
		const FastHull h = makeFastHull(q.sharedMins, q.sharedMaxs, q.sharedMask, ctx->slabOK && ctx->slabTest);
		if (ctx->counting) CUDA_TRY(ctx, launchOnMap(ctx, traceFastKernel<true>, (int)blocks, kBlock, 0, s, ctx->fast, q, h));
		else               CUDA_TRY(ctx, launchOnMap(ctx, traceFastKernel<false>, (int)blocks, kBlock, 0, s, ctx->fast, q, h));
	} else {
		if (deep)               CUDA_TRY(ctx, launchOnMap(ctx, traceKernel<false, kDeepDepth>, (int)blocks, kBlock, 0, s, ctx->map, q));
		else if (ctx->counting) CUDA_TRY(ctx, launchOnMap(ctx, traceKernel<true>, (int)blocks, kBlock, 0, s, ctx->map, q));
		else                    CUDA_TRY(ctx, launchOnMap(ctx, traceKernel<false>, (int)blocks, kBlock, 0, s, ctx->map, q));
	}
	ctx->launches++;
	CUDA_TRY(ctx, cudaGetLastError());
	if (compactOnly) {
		// the records are written later, by expandRecordsKernel or by the host
	} else if (q.compact && sideExpand) {
		int ctas = (q.n + kExpandBlock - 1) / kExpandBlock;
		if (ctas > ctx->remoteExpandCtas) ctas = ctx->remoteExpandCtas;
		CUDA_TRY(ctx, cudaEventRecord(ctx->evWalked, s));
		CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->sideStream, ctx->evWalked, 0));
		expandRecordsKernel<<<ctas, kExpandBlock, 0, ctx->sideStream>>>(ctx->fast, q.n, q.compactByItem ? nullptr : slotOf, q.compact, q.starts,
		                                                                 q.ends, q.out, q.hitIds);
		ctx->launches++;
		CUDA_TRY(ctx, cudaGetLastError());
	} else if (q.compact) {
		expandRecordsKernel<<<(q.n + kExpandBlock - 1) / kExpandBlock, kExpandBlock, 0, s>>>(ctx->fast, q.n, q.compactByItem ? nullptr : slotOf, q.compact, q.starts,
		                                                                                   q.ends, q.out, q.hitIds);
		ctx->launches++;
		CUDA_TRY(ctx, cudaGetLastError());
	}
	if (tEnd) CUDA_TRY(ctx, cudaEventRecord(tEnd, s));     // the bracket covers both phases: walk + record expansion
	if (external && !capturing) {
		CUDA_TRY(ctx, cudaEventRecord(ctx->extDone, s));
		ctx->extStream = s;
		ctx->extBusy = true;
	}
	return CMGPU_OK;
}

}  // namespace

// ---- pinned host memory for the host-pointer calls --------------------------------------------------------------
// cudaMemcpyAsync from or to pageable memory is staged by the driver and blocks the calling thread, so the pieces of a
// host batch stop overlapping (copy in, trace, copy out run one after the other).  The engine's callers hand over
// long-lived memory -- the QVM's data segment (sv_game.c:198, VMA(1)), a results arena -- which can be pinned ONCE.
static int registerRange(cmgpu_ctx *ctx, const void *ptr, size_t bytes) {
	if (!ptr || !bytes) return CMGPU_OK;
	uintptr_t lo = (uintptr_t)ptr & ~(uintptr_t)4095, hi = ((uintptr_t)ptr + bytes + 4095) & ~(uintptr_t)4095;
	// regions of ours that the new one touches are merged into it (a registration cannot overlap another)
	for (size_t i = 0; i < ctx->hostRegs.size();) {
		auto &r = ctx->hostRegs[i];
		if (r.first <= lo && hi <= r.second) return CMGPU_OK;
		if (r.first < hi && lo < r.second) {
			cudaHostUnregister((void *)r.first);
			if (r.first < lo) lo = r.first;
			if (r.second > hi) hi = r.second;
			ctx->hostRegs.erase(ctx->hostRegs.begin() + (long)i);
		} else i++;
	}
	cudaError_t e = cudaHostRegister((void *)lo, hi - lo, cudaHostRegisterDefault);
	if (e == cudaErrorHostMemoryAlreadyRegistered) { cudaGetLastError(); return CMGPU_OK; }   // pinned by its owner: nothing to do
	if (e != cudaSuccess) {
		cudaGetLastError();
		return fail(ctx, CMGPU_ERR_CUDA, "cudaHostRegister(%p, %zu bytes): %s", (void *)lo, (size_t)(hi - lo), cudaGetErrorString(e));
	}
	ctx->hostRegs.push_back({lo, hi});
	return CMGPU_OK;
}

static bool isPageable(const void *p) {
	cudaPointerAttributes a{};
	if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return true; }
	return a.type == cudaMemoryTypeUnregistered;
}

// "auto_register": pin the big arrays of a host batch on first sight (see cm_gpu.h for when that is safe)
static void autoRegister(cmgpu_ctx *ctx, const void *ptr, size_t bytes) {
	if (!ctx->autoRegister || !ptr || bytes < ((size_t)1 << 20)) return;
	if (!isPageable(ptr) && !isPageable((const char *)ptr + bytes - 1)) return;
	if (registerRange(ctx, ptr, bytes) != CMGPU_OK) ctx->err.clear();      // not fatal: the copies are merely staged
}


// ---- records that land in another GPU's memory ---------------------------------------------------------------------
constexpr int CMGPU_ERR_UNSUPPORTED_INTERNAL = -1000;    // tracePieces: not a batch it handles, take the plain launch

bool isRemote(cmgpu_ctx *ctx, const void *p) {
	for (auto &r : ctx->peerRanges) {
		if (r.first <= (uintptr_t)p && (uintptr_t)p < r.second) return true;
	}
	return false;
}

// A big hot-path batch whose records go to a peer: pieces of >= poolMinItems items are binned and walked one after the
// other on the caller's stream; each piece's expansion -- narrow, NVLink-bound -- runs on the side stream while the next
// piece is walked.  The caller's stream waits for the last expansion before the call's work counts as done.
int tracePieces(cmgpu_ctx *ctx, const TraceBatch &q0, cudaStream_t s) {
	const bool hot = ctx->useFast && ctx->usePool && ctx->twoPhase && ctx->binning && !ctx->counting && !q0.mins && !q0.masks && !q0.models && !q0.origins &&
	                 (ctx->numPatches == 0 || ctx->noCurves) && ((uintptr_t)q0.out & 15) == 0;
	int pieces = ctx->remotePieces;
	if (pieces > q0.n / ctx->pipePoolMinItems) pieces = q0.n / ctx->pipePoolMinItems;
	if (!hot || pieces < 2) return CMGPU_ERR_UNSUPPORTED_INTERNAL;
	cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
	if (cudaStreamIsCapturing(s, &st) != cudaSuccess) cudaGetLastError();
	if (st != cudaStreamCaptureStatusNone) return CMGPU_ERR_UNSUPPORTED_INTERNAL;
	if (!ctx->sideStream) {
		CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->sideStream, cudaStreamNonBlocking));
		CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->evWalked, cudaEventDisableTiming));
		CUDA_TRY(ctx, cudaEventCreateWithFlags(&ctx->evSideDone, cudaEventDisableTiming));
	}
	int rc;
	if ((rc = extWait(ctx, s))) return rc;
	const size_t N = (size_t)q0.n;
	if ((rc = reserve(ctx, ctx->bKeys, N * 4))) return rc;
	if ((rc = reserve(ctx, ctx->bRecs, N * 32))) return rc;
	if ((rc = reserve(ctx, ctx->bCompact, N * 16))) return rc;
	for (int p = 0; p < pieces; p++) {
		const size_t lo = ((N * (size_t)p / (size_t)pieces) & ~(size_t)255), hi = p + 1 == pieces ? N : ((N * (size_t)(p + 1) / (size_t)pieces) & ~(size_t)255);
		TraceBatch q = q0;
		q.n = (int)(hi - lo);
		q.starts = q0.starts + 3 * lo;
		q.ends = q0.ends + 3 * lo;
		q.out = q0.out + 7 * lo;
		q.hitIds = q0.hitIds ? q0.hitIds + lo : nullptr;
		if ((rc = launchTrace(ctx, q, s, lo, kBinSlots - 1, true))) return rc;
	}
	CUDA_TRY(ctx, cudaEventRecord(ctx->evSideDone, ctx->sideStream));
	CUDA_TRY(ctx, cudaStreamWaitEvent(s, ctx->evSideDone, 0));
	return extRecord(ctx, s);
}

// =====================================================================================
// C ABI
// =====================================================================================
extern "C" {

int cmgpu_abi_version(void) { return CMGPU_ABI_VERSION; }

int cmgpu_device_count(void) {
	int n = 0;
	if (cudaGetDeviceCount(&n) != cudaSuccess) {
		cudaGetLastError();
		return 0;
	}
	return n;
}

const char *cmgpu_last_error(const cmgpu_ctx *ctx) { return ctx ? ctx->err.c_str() : g_createError.c_str(); }

int cmgpu_create(cmgpu_ctx **out, int device) {
	if (!out) return fail(nullptr, CMGPU_ERR_INVALID, "cmgpu_create: out is NULL");
	*out = nullptr;
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n == 0) {
		cudaGetLastError();
		return fail(nullptr, CMGPU_ERR_NO_DEVICE, "no CUDA device (%s); this engine has no CPU path",
		            e == cudaSuccess ? "count is 0" : cudaGetErrorString(e));
	}
	if (device < 0 || device >= n) return fail(nullptr, CMGPU_ERR_INVALID, "device %d out of range (%d devices)", device, n);
	cudaDeviceProp prop;
	if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return fail(nullptr, CMGPU_ERR_CUDA, "cudaGetDeviceProperties failed");
	if (prop.major != 10) {
		return fail(nullptr, CMGPU_ERR_NO_DEVICE, "device %d is sm_%d%d; libcmgpu is built for sm_100a only", device, prop.major, prop.minor);
	}
	cmgpu_ctx *ctx = new cmgpu_ctx_s();
	ctx->device = device;
	if (const char *g = getenv("CMGPU_GANG")) ctx->gang = atoi(g);   // tuning knobs
	if (const char *g = getenv("CMGPU_LIGHT_REPS")) ctx->lightReps = atoi(g) > 0 ? atoi(g) : 1;
	if (const char *g = getenv("CMGPU_POOL")) ctx->usePool = atoi(g) != 0;
	if (const char *g = getenv("CMGPU_PIPE_CHUNKS")) ctx->pipeChunks = atoi(g) > 0 ? (atoi(g) > kPipeChunks ? kPipeChunks : atoi(g)) : ctx->pipeChunks;
	if (const char *g = getenv("CMGPU_BIN")) ctx->binning = atoi(g) != 0;
	if (const char *g = getenv("CMGPU_BIN_CELL")) ctx->binCell = atof(g) > 1.0 ? (float)atof(g) : 128.0f;
	if (const char *g = getenv("CMGPU_BIN_LONG")) ctx->binLongLen = (float)atof(g);
	cudaSetDevice(device);
	bool ok = true;
	for (auto &s : ctx->streams) ok = ok && cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) == cudaSuccess;
	for (auto &ev : ctx->evt) ok = ok && cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) == cudaSuccess;
	ok = ok && cudaEventCreateWithFlags(&ctx->extDone, cudaEventDisableTiming) == cudaSuccess;
	ok = ok && cudaMalloc(&ctx->dCounters, sizeof(Counters)) == cudaSuccess;
	ok = ok && cudaMemset(ctx->dCounters, 0, sizeof(Counters)) == cudaSuccess;
	ok = ok && cudaMalloc(&ctx->dStatus, sizeof(int)) == cudaSuccess;
	ok = ok && cudaMemset(ctx->dStatus, 0, sizeof(int)) == cudaSuccess;
	ok = ok && cudaMallocHost(&ctx->hStatus, sizeof(int)) == cudaSuccess;
	if (ok && cudaHostAlloc((void **)&ctx->tinyHost, TinyLayout::total, cudaHostAllocMapped) == cudaSuccess) {
		if (cudaHostGetDevicePointer((void **)&ctx->tinyDev, ctx->tinyHost, 0) != cudaSuccess) { cudaFreeHost(ctx->tinyHost); ctx->tinyHost = nullptr; cudaGetLastError(); }
	} else if (ok) { ctx->tinyHost = nullptr; cudaGetLastError(); }
	ok = ok && cudaMalloc(&ctx->dCursor, kCursorSlots * sizeof(int)) == cudaSuccess;
	if (ok) {
		int perSm = 0;
		ok = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, traceKernel<false>, kBlock, 0) == cudaSuccess && perSm > 0;
		ctx->gridPlain = perSm * prop.multiProcessorCount;
		int perSmC = 0;
		ok = ok && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSmC, traceKernel<true>, kBlock, 0) == cudaSuccess && perSmC > 0;
		ctx->gridCount = perSmC * prop.multiProcessorCount;
		int perSmF = 0, perSmFC = 0;
		ok = ok && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSmF, traceFastKernel<false>, kBlock, 0) == cudaSuccess && perSmF > 0;
		ok = ok && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSmFC, traceFastKernel<true>, kBlock, 0) == cudaSuccess && perSmFC > 0;
		ctx->gridFast = perSmF * prop.multiProcessorCount;
		ctx->gridFastCount = perSmFC * prop.multiProcessorCount;
		int perSmS = 0;
		ok = ok && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSmS, traceSimpleKernel<false>, kBlock, 0) == cudaSuccess && perSmS > 0;
		ctx->gridSimple = perSmS * prop.multiProcessorCount;
		int perSmP = 0, perSmPC = 0;
		ok = ok && cudaFuncSetAttribute(tracePoolKernel<false, CMGPU_POOL_K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPoolSmem) == cudaSuccess;
		ok = ok && cudaFuncSetAttribute(tracePoolKernel<false, CMGPU_POOL_K, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPoolSmem) == cudaSuccess;
		ok = ok && cudaFuncSetAttribute(tracePoolKernel<true, CMGPU_POOL_K, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kPoolSmem) == cudaSuccess;
		int perSmPT = 0;
		ok = ok && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSmP, tracePoolKernel<false, CMGPU_POOL_K, false>, kBlock, kPoolSmem) == cudaSuccess && perSmP > 0;
		ok = ok && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSmPT, tracePoolKernel<false, CMGPU_POOL_K, true>, kBlock, kPoolSmem) == cudaSuccess && perSmPT > 0;
		if (perSmPT < perSmP) perSmP = perSmPT;
		ok = ok && cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSmPC, tracePoolKernel<true, CMGPU_POOL_K, false>, kBlock, kPoolSmem) == cudaSuccess && perSmPC > 0;
		ctx->gridPool = perSmP * prop.multiProcessorCount;
		ctx->gridPoolCount = perSmPC * prop.multiProcessorCount;
	}
	if (!ok) {
		g_createError = std::string("context setup failed: ") + cudaGetErrorString(cudaGetLastError());
		cmgpu_destroy(ctx);
		return CMGPU_ERR_CUDA;
	}
	*out = ctx;
	return CMGPU_OK;
}

void cmgpu_destroy(cmgpu_ctx *ctx) {
	if (!ctx) return;
	cudaSetDevice(ctx->device);
	cudaDeviceSynchronize();
	freeMap(ctx);
	for (DevBuf *b : {&ctx->bStarts, &ctx->bEnds, &ctx->bMins, &ctx->bMaxs, &ctx->bModels, &ctx->bMasks, &ctx->bOrigins,
	                  &ctx->bRot, &ctx->bRotIdx, &ctx->bBoxMins, &ctx->bBoxMaxs, &ctx->bOut, &ctx->bHit,
	                  &ctx->bKeys, &ctx->bRecs, &ctx->bCompact, &ctx->bSvWorld, &ctx->bSvOffsets, &ctx->bSvTiles, &ctx->bSvItems, &ctx->bSvCand,
	                  &ctx->bSvPass, &ctx->bLeafLists, &ctx->bLeafCounts, &ctx->bVis, &ctx->bFlood, &ctx->bEntRecs, &ctx->bVisible}) releaseBuf(*b);
	if (ctx->hSvTotal) cudaFreeHost(ctx->hSvTotal);
	if (ctx->tinyHost) cudaFreeHost(ctx->tinyHost);
	if (ctx->hostPool) hostrec::poolDestroy(ctx->hostPool);
	if (ctx->hCompact) cudaFreeHost(ctx->hCompact);
	if (ctx->hStage) cudaFreeHost(ctx->hStage);
	for (auto &e : ctx->hybridEvents) cudaEventDestroy(e);
	for (cudaStream_t st : {ctx->inStream, ctx->outHostStream, ctx->outDevStream}) if (st) cudaStreamDestroy(st);
	releaseBuf(ctx->bAas); releaseBuf(ctx->bAasIn); releaseBuf(ctx->bAasOut);
	for (auto &r : ctx->hostRegs) cudaHostUnregister((void *)r.first);
	ctx->hostRegs.clear();
	for (auto &b : ctx->bHist) releaseBuf(b);
	for (auto &b : ctx->bStackArena) releaseBuf(b);
	for (auto &b : ctx->bPoolHdr) releaseBuf(b);
	if (ctx->extDone) cudaEventDestroy(ctx->extDone);
	if (ctx->sideStream) { cudaStreamDestroy(ctx->sideStream); cudaEventDestroy(ctx->evWalked); cudaEventDestroy(ctx->evSideDone); }
	for (auto &pr : ctx->timingPairs) { cudaEventDestroy(pr.first); cudaEventDestroy(pr.second); }
	if (ctx->dCounters) cudaFree(ctx->dCounters);
	if (ctx->dStatus) cudaFree(ctx->dStatus);
	if (ctx->dCursor) cudaFree(ctx->dCursor);
	if (ctx->hStatus) cudaFreeHost(ctx->hStatus);
	for (auto &s : ctx->streams) if (s) cudaStreamDestroy(s);
	for (auto &ev : ctx->evt) if (ev) cudaEventDestroy(ev);
	delete ctx;
}

int cmgpu_upload(cmgpu_ctx *ctx, const cmgpu_flatmap *map) {
	if (!ctx || !map) return fail(ctx, CMGPU_ERR_INVALID, "cmgpu_upload: NULL argument");
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	int levels = 1;
	int rc = validate(ctx, map, &levels);
	if (rc) return rc;
	CUDA_TRY(ctx, cudaDeviceSynchronize());
	freeMap(ctx);
	ctx->treeDepth = levels;
	ctx->hasSvWorld = false;                     // entity snapshots refer to the inline models of the map they were made for
	rc = packAndUpload(ctx, map);
	if (rc) freeMap(ctx);
	else memcpy(ctx->worldBounds, map->modelBounds, sizeof(ctx->worldBounds));
	return rc;
}

int cmgpu_clear(cmgpu_ctx *ctx) {
	if (!ctx) return CMGPU_ERR_INVALID;
	cudaSetDevice(ctx->device);
	cudaDeviceSynchronize();
	freeMap(ctx);
	return CMGPU_OK;
}

int cmgpu_set_cvars(cmgpu_ctx *ctx, int noCurves, int playerCurveClip) {
	if (!ctx) return CMGPU_ERR_INVALID;
	ctx->noCurves = noCurves;
	ctx->playerCurveClip = playerCurveClip;
	return CMGPU_OK;
}

int cmgpu_num_inline_models(const cmgpu_ctx *ctx) { return ctx ? ctx->numModels : 0; }

int cmgpu_set_option(cmgpu_ctx *ctx, const char *name, double value) {
	if (!ctx || !name) return CMGPU_ERR_INVALID;
	const std::string k(name);
	if (k == "bin") ctx->binning = value != 0;
	else if (k == "bin_min_items") ctx->binMinItems = value < 1 ? 1 : (int)value;
	else if (k == "bin_cell") { if (!(value > 1.0)) return fail(ctx, CMGPU_ERR_INVALID, "bin_cell must be > 1"); ctx->binCell = (float)value; }
	else if (k == "bin_long_len") ctx->binLongLen = (float)value;
	else if (k == "bin_dir") ctx->binDir = value != 0;
	else if (k == "fast") ctx->useFast = value != 0;
	else if (k == "l2_window") ctx->l2Window = value != 0;      // takes effect at the next cmgpu_upload
	else if (k == "pool") ctx->usePool = value != 0;
	else if (k == "pool_reps") ctx->poolReps = value < 1 ? 1 : (int)value;
	else if (k == "min_chunk") ctx->minChunk = value < 1 ? 1 : (value > 32 ? 32 : (int)value);
	else if (k == "simple_max_items") ctx->simpleMaxItems = value < 0 ? 0 : (int)value;
	else if (k == "pool_min_items") ctx->poolMinItems = value < 1 ? 1 : (int)value;
	else if (k == "pipe_pool_min_items") ctx->pipePoolMinItems = value < 1 ? 1 : (int)value;
	else if (k == "pipe_ramp") ctx->pipeRamp = value != 0;
	else if (k == "pipe_chunks") ctx->pipeChunks = value < 1 ? 1 : (value > kPipeChunks ? kPipeChunks : (int)value);
	else if (k == "slab") ctx->slabTest = value != 0;
	else if (k == "two_phase") ctx->twoPhase = value != 0;
	else if (k == "two_phase_host") ctx->twoPhaseHost = value != 0;
	else if (k == "two_phase_small") ctx->twoPhaseSmall = value != 0;
	else if (k == "compact_by_item") ctx->compactByItem = value != 0;
	else if (k == "node_thresholds") ctx->nodeThresholds = value != 0;
	else if (k == "entry_grid") ctx->entryGrid = value != 0;
	else if (k == "hull_window") ctx->hullWindow = value != 0;
	else if (k == "auto_register") ctx->autoRegister = value != 0;
	else if (k == "host_records") ctx->hostRecords = value < 0 ? -1 : (value > 3 ? 3 : (int)value);
	else if (k == "host_threads") {
		const int t = value < 0 ? -1 : (value > 256 ? 256 : (int)value);
		if (t != ctx->hostThreads && ctx->hostPool) { hostrec::poolDestroy(ctx->hostPool); ctx->hostPool = nullptr; }
		ctx->hostThreads = t;
	}
	else if (k == "stage_pageable") ctx->stagePageable = value != 0;
	else if (k == "host_share") ctx->hostShareFixed = value < 0 ? -1.0 : (value > 1 ? 1.0 : value);
	else if (k == "hybrid_decouple") ctx->hybridDecouple = value < 0 ? 0 : (value > 2 ? 2 : (int)value);
	else if (k == "hybrid_grid_eighths") ctx->hybridGridEighths = value < 1 ? 1 : (value > 8 ? 8 : (int)value);
	else if (k == "hybrid_in_flight") ctx->hybridInFlight = value < 1 ? 1 : (int)value;
	else if (k == "hybrid_min_items") ctx->hybridMinItems = value < 1 ? 1 : (size_t)value;
	else if (k == "hybrid_walk_items") ctx->hybridWalkItems = value < 98304 ? 98304 : (size_t)value;
	else if (k == "tiny_calls") ctx->tinyCalls = value != 0;
	else if (k == "remote_pipeline") ctx->remotePipeline = value != 0;
	else if (k == "remote_pieces") ctx->remotePieces = value < 1 ? 1 : (int)value;
	else if (k == "remote_expand_ctas") ctx->remoteExpandCtas = value < 1 ? 1 : (int)value;
	else if (k == "entry_cell") { if (!(value >= 8.0)) return fail(ctx, CMGPU_ERR_INVALID, "entry_cell must be >= 8"); ctx->entryCell = (float)value; }   // takes effect at the next cmgpu_upload
	else if (k == "gang") ctx->gang = ctx->gangPool = (int)value;
	else if (k == "gang_pool") ctx->gangPool = (int)value;
	else if (k == "gang_fetch") ctx->gangFetch = ctx->gangFetchPool = (int)value;
	else if (k == "gang_brush") ctx->gangBrush = (int)value;
	else if (k == "facet_steps") ctx->facetSteps = value < 1 ? 1 : (int)value;
	else if (k == "light_reps") ctx->lightReps = value < 1 ? 1 : (int)value;
	else if (k == "simple_spread") ctx->simpleSpread = value < 0 ? -1 : (value > 5 ? 5 : (int)value);
	else return fail(ctx, CMGPU_ERR_INVALID, "cmgpu_set_option: unknown option '%s'", name);
	return CMGPU_OK;
}

void cmgpu_rotation_matrices(int n, const float *angles, float *matrices, int32_t *rot_index) {
	// AngleVectors, q_math.c:999-1032: angle scaled in double and rounded to float, sin/cos of that
	// float in double rounded to float, all products/sums in float (this TU is built with -fmad=false
	// for device AND host code, so nothing below is contracted).
	const double k = M_PI * 2 / 360;
	for (int i = 0; i < n; i++) {
		const float *a = angles + 3 * (size_t)i;
		if (!(a[0] || a[1] || a[2])) {
			rot_index[i] = -1;
			continue;
		}
		rot_index[i] = i;
		float ang;
		ang = (float)((double)a[1] * k); const float sy = (float)sin((double)ang), cy = (float)cos((double)ang);
		ang = (float)((double)a[0] * k); const float sp = (float)sin((double)ang), cp = (float)cos((double)ang);
		ang = (float)((double)a[2] * k); const float sr = (float)sin((double)ang), cr = (float)cos((double)ang);
		float *m = matrices + 9 * (size_t)i;
		// row 0: forward
		m[0] = cp * cy;
		m[1] = cp * sy;
		m[2] = -sp;
		// row 1: -right (CreateRotationMatrix negates the right vector, cm_trace.c:65-68)
		const float nsr = -1 * sr, ncr = -1 * cr;
		m[3] = -((nsr * sp) * cy + ncr * -sy);
		m[4] = -((nsr * sp) * sy + ncr * cy);
		m[5] = -(nsr * cp);
		// row 2: up
		m[6] = (cr * sp) * cy + -sr * -sy;
		m[7] = (cr * sp) * sy + -sr * cy;
		m[8] = cr * cp;
	}
}

int cmgpu_enable_counters(cmgpu_ctx *ctx, int enable) {
	if (!ctx) return CMGPU_ERR_INVALID;
	ctx->counting = enable != 0;
	return CMGPU_OK;
}

int cmgpu_read_counters(cmgpu_ctx *ctx, cmgpu_counters *out, int reset) {
	if (!ctx || !out) return CMGPU_ERR_INVALID;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	CUDA_TRY(ctx, cudaDeviceSynchronize());
	CUDA_TRY(ctx, cudaMemcpy(out, ctx->dCounters, sizeof(Counters), cudaMemcpyDeviceToHost));
	if (reset) CUDA_TRY(ctx, cudaMemset(ctx->dCounters, 0, sizeof(Counters)));
	return CMGPU_OK;
}

unsigned long long cmgpu_launch_count(const cmgpu_ctx *ctx) { return ctx ? ctx->launches : 0; }

// ---- result buffers shared between the processes of one node (one process per GPU) ------------------
// Rank A allocates, the others open A's buffer and hand (pointer + their offset) to the device-pointer entry points:
// their trace kernels then store their 56-byte records straight into A's memory over NVLink while they run -- the
// result gather is the kernels' own stores, not a pass after them.
int cmgpu_ipc_alloc(cmgpu_ctx *ctx, size_t bytes, void **devptr, unsigned char handle[64]) {
	if (!ctx || !devptr || !handle || !bytes) return fail(ctx, CMGPU_ERR_INVALID, "cmgpu_ipc_alloc: bad argument");
	static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	void *p = nullptr;
	CUDA_TRY(ctx, cudaMalloc(&p, bytes));
	cudaIpcMemHandle_t h;
	cudaError_t e = cudaIpcGetMemHandle(&h, p);
	if (e != cudaSuccess) {
		cudaFree(p);
		return fail(ctx, CMGPU_ERR_CUDA, "cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
	}
	memcpy(handle, &h, 64);
	*devptr = p;
	return CMGPU_OK;
}

int cmgpu_ipc_open(cmgpu_ctx *ctx, const unsigned char handle[64], void **devptr) {
	if (!ctx || !devptr || !handle) return fail(ctx, CMGPU_ERR_INVALID, "cmgpu_ipc_open: bad argument");
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	cudaIpcMemHandle_t h;
	memcpy(&h, handle, 64);
	CUDA_TRY(ctx, cudaIpcOpenMemHandle(devptr, h, cudaIpcMemLazyEnablePeerAccess));
	// remember the mapping: records stored into it travel over NVLink, and big batches are streamed there (tracePieces)
	CUdeviceptr base = (CUdeviceptr)*devptr;
	size_t size = 1;
	{	// the driver call through the runtime's entry-point lookup: libcmgpu.so does not link libcuda
		typedef CUresult (*RangeFn)(CUdeviceptr *, size_t *, CUdeviceptr);
		void *fn = nullptr;
		cudaDriverEntryPointQueryResult qr;
		if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr) == cudaSuccess && fn) {
			CUdeviceptr b = 0;
			size_t sz = 0;
			if (((RangeFn)fn)(&b, &sz, (CUdeviceptr)*devptr) == CUDA_SUCCESS && sz) { base = b; size = sz; }
		} else cudaGetLastError();
	}
	ctx->peerRanges.push_back({(uintptr_t)base, (uintptr_t)base + size});
	return CMGPU_OK;
}

int cmgpu_ipc_close(cmgpu_ctx *ctx, void *devptr) {
	if (!ctx) return CMGPU_ERR_INVALID;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	CUDA_TRY(ctx, cudaDeviceSynchronize());
	CUDA_TRY(ctx, cudaIpcCloseMemHandle(devptr));
	for (size_t i = 0; i < ctx->peerRanges.size(); i++) {
		if (ctx->peerRanges[i].first <= (uintptr_t)devptr && (uintptr_t)devptr < ctx->peerRanges[i].second) { ctx->peerRanges.erase(ctx->peerRanges.begin() + (long)i); break; }
	}
	return CMGPU_OK;
}

int cmgpu_ipc_free(cmgpu_ctx *ctx, void *devptr) {
	if (!ctx) return CMGPU_ERR_INVALID;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	CUDA_TRY(ctx, cudaDeviceSynchronize());
	CUDA_TRY(ctx, cudaFree(devptr));
	return CMGPU_OK;
}

int cmgpu_host_register(cmgpu_ctx *ctx, const void *ptr, size_t bytes) {
	if (!ctx) return CMGPU_ERR_INVALID;
	if (!ptr || !bytes) return fail(ctx, CMGPU_ERR_INVALID, "cmgpu_host_register: empty region");
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	return registerRange(ctx, ptr, bytes);
}

int cmgpu_format_records(int n, const uint32_t *results16, const float *starts, const float *ends,
                         int num_sides, const float *side_planes, const int32_t *side_type_signbits, const int32_t *side_surface_flags,
                         int num_brushes, const int32_t *brush_contents, cmgpu_trace_t *records, int32_t *hit_ids) {
	if (n < 0 || num_sides < 0 || num_brushes < 0 || (n > 0 && (!results16 || !starts || !ends || !records))) return CMGPU_ERR_INVALID;
	if ((num_sides > 0 && (!side_planes || !side_type_signbits || !side_surface_flags)) || (num_brushes > 0 && !brush_contents)) return CMGPU_ERR_INVALID;
	for (int i = 0; i < n; i++) {
		const int32_t side = (int32_t)results16[4 * (size_t)i + 2], brush = (int32_t)results16[4 * (size_t)i + 3];
		if (side < -1 || side >= num_sides || brush < -1 || brush >= num_brushes) return CMGPU_ERR_INVALID;
	}
	std::vector<hostrec::SideRec> sides((size_t)num_sides + 1, hostrec::SideRec{});
	std::vector<int32_t> contents((size_t)num_brushes + 1, 0);
	for (int i = 0; i < num_sides; i++) {
		hostrec::SideRec &r = sides[(size_t)i + 1];
		memcpy(r.plane, side_planes + 4 * (size_t)i, 16);
		r.typeSignbits = (uint32_t)side_type_signbits[i] & 0xffffu;
		r.surfaceFlags = (uint32_t)side_surface_flags[i];
	}
	for (int i = 0; i < num_brushes; i++) contents[(size_t)i + 1] = brush_contents[i];
	hostrec::expand(hostrec::Tables{sides.data(), contents.data()}, results16, starts, ends, records, hit_ids, 0, (size_t)n);
	return CMGPU_OK;
}

int cmgpu_format_records_threads(int threads, int n, const uint32_t *results16, const float *starts, const float *ends,
                                 int num_sides, const float *side_planes, const int32_t *side_type_signbits, const int32_t *side_surface_flags,
                                 int num_brushes, const int32_t *brush_contents, cmgpu_trace_t *records, int32_t *hit_ids) {
	// the same through the worker pool of a big host batch, the way traceHostHybrid drives it: the sweeps are staged into a
	// scratch copy a piece ahead, every piece's job is queued held and released later (here: by this thread, last piece first)
	if (threads < 1 || n < 0 || num_sides < 0 || num_brushes < 0 || (n > 0 && (!results16 || !starts || !ends || !records))) return CMGPU_ERR_INVALID;
	if ((num_sides > 0 && (!side_planes || !side_type_signbits || !side_surface_flags)) || (num_brushes > 0 && !brush_contents)) return CMGPU_ERR_INVALID;
	for (int i = 0; i < n; i++) {
		const int32_t side = (int32_t)results16[4 * (size_t)i + 2], brush = (int32_t)results16[4 * (size_t)i + 3];
		if (side < -1 || side >= num_sides || brush < -1 || brush >= num_brushes) return CMGPU_ERR_INVALID;
	}
	std::vector<hostrec::SideRec> sides((size_t)num_sides + 1, hostrec::SideRec{});
	std::vector<int32_t> contents((size_t)num_brushes + 1, 0);
	for (int i = 0; i < num_sides; i++) {
		hostrec::SideRec &r = sides[(size_t)i + 1];
		memcpy(r.plane, side_planes + 4 * (size_t)i, 16);
		r.typeSignbits = (uint32_t)side_type_signbits[i] & 0xffffu;
		r.surfaceFlags = (uint32_t)side_surface_flags[i];
	}
	for (int i = 0; i < num_brushes; i++) contents[(size_t)i + 1] = brush_contents[i];
	const size_t N = (size_t)n, piece = 50000;
	std::vector<float> stagedStarts(3 * N + 1), stagedEnds(3 * N + 1);
	hostrec::Pool *pool = hostrec::poolCreate(threads);
	hostrec::poolSetEpoch(pool);
	std::vector<int> tickets;
	for (size_t lo = 0; lo < N; lo += piece) {
		const size_t hi = lo + piece < N ? lo + piece : N;
		hostrec::Job st{};
		st.starts = starts; st.ends = ends; st.stageStarts = stagedStarts.data(); st.stageEnds = stagedEnds.data();
		st.lo = lo; st.hi = hi;
		hostrec::poolWaitJob(pool, hostrec::poolSubmit(pool, st));
		hostrec::Job job{};
		job.tables = hostrec::Tables{sides.data(), contents.data()};
		job.compact = results16;
		job.starts = stagedStarts.data(); job.ends = stagedEnds.data();      // the staged copies: both jobs are checked by the result
		job.records = records; job.hitIds = hit_ids;
		job.lo = lo; job.hi = hi;
		job.held = true;
		tickets.push_back(hostrec::poolSubmit(pool, job));
	}
	for (size_t k = tickets.size(); k-- > 0;) hostrec::poolRelease(pool, tickets[k]);
	hostrec::poolWaitAll(pool);
	const bool idle = hostrec::poolPendingRecords(pool) == 0;
	hostrec::poolDestroy(pool);
	return idle ? CMGPU_OK : CMGPU_ERR_INVALID;
}

int cmgpu_last_batch_split(const cmgpu_ctx *ctx, unsigned long long *by_host, unsigned long long *by_device) {
	if (!ctx) return CMGPU_ERR_INVALID;
	if (by_host) *by_host = ctx->hostRecordsWritten;
	if (by_device) *by_device = ctx->deviceRecordsWritten;
	return CMGPU_OK;
}

int cmgpu_host_unregister(cmgpu_ctx *ctx, const void *ptr) {
	if (!ctx || !ptr) return CMGPU_ERR_INVALID;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	CUDA_TRY(ctx, cudaDeviceSynchronize());          // nothing may still copy from or to it
	const uintptr_t p = (uintptr_t)ptr;
	for (size_t i = 0; i < ctx->hostRegs.size(); i++) {
		if (ctx->hostRegs[i].first <= p && p < ctx->hostRegs[i].second) {
			cudaHostUnregister((void *)ctx->hostRegs[i].first);
			cudaGetLastError();
			ctx->hostRegs.erase(ctx->hostRegs.begin() + (long)i);
			return CMGPU_OK;
		}
	}
	return fail(ctx, CMGPU_ERR_INVALID, "cmgpu_host_unregister: %p is not inside a region registered through this context", ptr);
}

int cmgpu_kernel_timing(cmgpu_ctx *ctx, int enable) {
	if (!ctx) return CMGPU_ERR_INVALID;
	ctx->timing = enable != 0;
	ctx->timingUsed = 0;
	return CMGPU_OK;
}

int cmgpu_kernel_time_ms(cmgpu_ctx *ctx, double *total_ms, int *launches) {
	if (!ctx || !total_ms) return CMGPU_ERR_INVALID;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	double sum = 0.0;
	for (size_t i = 0; i < ctx->timingUsed; i++) {
		CUDA_TRY(ctx, cudaEventSynchronize(ctx->timingPairs[i].second));
		float ms = 0.0f;
		CUDA_TRY(ctx, cudaEventElapsedTime(&ms, ctx->timingPairs[i].first, ctx->timingPairs[i].second));
		sum += ms;
	}
	*total_ms = sum;
	if (launches) *launches = (int)ctx->timingUsed;
	ctx->timingUsed = 0;
	return CMGPU_OK;
}

int cmgpu_check_status(cmgpu_ctx *ctx, void *stream) {
	if (!ctx) return CMGPU_ERR_INVALID;
	cudaStream_t s = (cudaStream_t)stream;
	CUDA_TRY(ctx, cudaMemcpyAsync(ctx->hStatus, ctx->dStatus, sizeof(int), cudaMemcpyDeviceToHost, s));
	CUDA_TRY(ctx, cudaStreamSynchronize(s));
	if (const int st = *ctx->hStatus) {
		CUDA_TRY(ctx, cudaMemsetAsync(ctx->dStatus, 0, sizeof(int), s));
		if (st & 2)
			return fail(ctx, CMGPU_ERR_BAD_HANDLE, "CM_ClipHandleToModel: a device-side model handle was neither a loaded inline model (0..%d) nor %d; "
			            "those items were traced against nothing", ctx->numModels - 1, CMGPU_BOX_MODEL_HANDLE);
		// cannot happen on a validated map (stacks are sized by the tree's depth, %d levels here); kept as a tripwire
		return fail(ctx, CMGPU_ERR_LIMIT, "traversal stack overflow: more than %d pending halves on a BSP of %d levels", ctx->treeDepth > kMaxDepth ? kDeepDepth : ctx->treeDepth, ctx->treeDepth);
	}
	return CMGPU_OK;
}

// ---- device-pointer entry points ---------------------------------------------------------
int cmgpu_box_trace_batch_device(cmgpu_ctx *ctx, int n,
                                 const float *starts, const float *ends,
                                 const float *mins, const float *maxs, int hull_stride,
                                 const int32_t *models, const int32_t *masks, int mask_stride,
                                 const float *origins, const float *rotations, const int32_t *rot_index,
                                 const float *boxmins, const float *boxmaxs,
                                 cmgpu_trace_t *results, int32_t *hit_ids, void *stream) {
	if (!ctx) return CMGPU_ERR_INVALID;
	if (n < 0 || (n > 0 && (!starts || !ends || !masks || !results)))
		return fail(ctx, CMGPU_ERR_INVALID, "cmgpu_box_trace_batch_device: bad argument");
	if (!ctx->hasMap) return fail(ctx, CMGPU_ERR_NO_MAP, "no map uploaded");
	if ((mins == nullptr) != (maxs == nullptr)) return fail(ctx, CMGPU_ERR_INVALID, "mins and maxs must both be given or both be NULL");
	if (hull_stride && !mins) hull_stride = 0;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	TraceBatch q{};
	q.n = n; q.starts = starts; q.ends = ends;
	if (hull_stride) { q.mins = mins; q.maxs = maxs; }
	else if (mins) { memcpy(q.sharedMins, mins, 12); memcpy(q.sharedMaxs, maxs, 12); }   // host [3], by value
	q.models = models;
	if (mask_stride) q.masks = masks; else q.sharedMask = masks[0];                         // host [1], by value
	q.origins = origins; q.rotations = rotations; q.rotIndex = rotations ? rot_index : nullptr;
	if (origins && !q.rotations) q.rotIndex = nullptr;
	q.boxmins = boxmins; q.boxmaxs = boxmins ? boxmaxs : nullptr;
	if (!q.boxmaxs) q.boxmins = nullptr;
	q.out = reinterpret_cast<uint2 *>(results); q.hitIds = hit_ids;
	q.status = ctx->dStatus; q.counters = ctx->dCounters;
	if (ctx->remotePipeline && isRemote(ctx, results)) {
		const int rc = tracePieces(ctx, q, (cudaStream_t)stream);
		if (rc != CMGPU_ERR_UNSUPPORTED_INTERNAL) return rc;
	}
	return launchTrace(ctx, q, (cudaStream_t)stream);
}

int cmgpu_point_contents_batch_device(cmgpu_ctx *ctx, int n, const float *points,
                                      const int32_t *models, const float *origins,
                                      const float *rotations, const int32_t *rot_index,
                                      const float *boxmins, const float *boxmaxs,
                                      int32_t *contents, void *stream) {
	if (!ctx) return CMGPU_ERR_INVALID;
	if (n < 0 || (n > 0 && (!points || !contents))) return fail(ctx, CMGPU_ERR_INVALID, "cmgpu_point_contents_batch_device: bad argument");
	if (!ctx->hasMap) return fail(ctx, CMGPU_ERR_NO_MAP, "no map uploaded");
	if (n == 0) return CMGPU_OK;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	PointBatch q{};
	q.n = n; q.points = points; q.models = models; q.origins = origins;
	q.rotations = rotations; q.rotIndex = rotations ? rot_index : nullptr;
	q.boxmins = boxmins; q.boxmaxs = boxmins ? boxmaxs : nullptr;
	if (!q.boxmaxs) q.boxmins = nullptr;
	q.out = contents;
	q.status = ctx->dStatus;
	const int grid = (n + 127) / 128;
	pointContentsKernel<<<grid, 128, 0, (cudaStream_t)stream>>>(ctx->map, q);
	ctx->launches++;
	CUDA_TRY(ctx, cudaGetLastError());
	return CMGPU_OK;
}

// ---- host-pointer entry points -----------------------------------------------------------
// The batch is cut into chunks; chunk c's H2D copy, kernel and D2H copy are issued on stream
// c % kBinSlots, so the copy engines (one per direction) and the SMs work on different chunks at once.
// ---- calls of a few items ------------------------------------------------------------------------------------------
static int traceHostTiny(cmgpu_ctx *ctx, int n, const float *starts, const float *ends, const float *mins, const float *maxs, int hull_stride,
                         const int32_t *models, const int32_t *masks, int mask_stride, const float *origins, const float *angles,
                         const float *boxmins, const float *boxmaxs, cmgpu_trace_t *results, int32_t *hit_ids) {
	typedef TinyLayout T;
	char *h = ctx->tinyHost, *d = ctx->tinyDev;
	const size_t N = (size_t)n;
	cudaStream_t s = ctx->streams[0];
	if (ctx->extBusy) {                              // an earlier device-pointer call may still use the status word
		CUDA_TRY(ctx, cudaEventSynchronize(ctx->extDone));
		ctx->extBusy = false;
	}
	memcpy(h + T::starts, starts, N * 12);
	memcpy(h + T::ends, ends, N * 12);
	TraceBatch q{};
	q.n = n;
	q.starts = (const float *)(d + T::starts);
	q.ends = (const float *)(d + T::ends);
	if (hull_stride) {
		memcpy(h + T::mins, mins, N * 12); memcpy(h + T::maxs, maxs, N * 12);
		q.mins = (const float *)(d + T::mins); q.maxs = (const float *)(d + T::maxs);
	} else if (mins) {
		memcpy(q.sharedMins, mins, 12); memcpy(q.sharedMaxs, maxs, 12);
	}
	if (models) { memcpy(h + T::models, models, N * 4); q.models = (const int *)(d + T::models); }
	if (mask_stride) { memcpy(h + T::masks, masks, N * 4); q.masks = (const int *)(d + T::masks); } else q.sharedMask = masks[0];
	if (origins) {
		memcpy(h + T::origins, origins, N * 12);
		cmgpu_rotation_matrices(n, angles, (float *)(h + T::rot), (int32_t *)(h + T::rotIdx));
		q.origins = (const float *)(d + T::origins);
		q.rotations = (const float *)(d + T::rot);
		q.rotIndex = (const int *)(d + T::rotIdx);
	}
	if (boxmins) {
		memcpy(h + T::boxmins, boxmins, N * 12); memcpy(h + T::boxmaxs, boxmaxs, N * 12);
		q.boxmins = (const float *)(d + T::boxmins); q.boxmaxs = (const float *)(d + T::boxmaxs);
	}
	q.out = reinterpret_cast<uint2 *>(d + T::out);
	q.hitIds = hit_ids ? (int *)(d + T::hit) : nullptr;
	q.status = (int *)(d + T::status);
	q.cursor = (int *)(d + T::cursor);
	*(volatile int *)(h + T::status) = 0;
	*(volatile int *)(h + T::cursor) = 0;
	q.counters = ctx->dCounters;
	q.gang = ctx->gang; q.gangFetch = ctx->gangFetch; q.gangBrush = ctx->gangBrush; q.lightReps = ctx->lightReps;
	q.facetSteps = ctx->facetSteps;
	q.stackDepth = ctx->treeDepth;
	q.chunk = ctx->minChunk;
	ctx->map.noCurves = ctx->noCurves;
	ctx->map.playerCurveClip = ctx->playerCurveClip;
	const bool deep = ctx->treeDepth > kMaxDepth;
	const bool fast = ctx->useFast && !q.mins && !q.masks && !q.models && !q.origins && (ctx->numPatches == 0 || ctx->noCurves);
	if (fast) {
		const FastHull fh = makeFastHull(q.sharedMins, q.sharedMaxs, q.sharedMask, ctx->slabOK && ctx->slabTest);
		q.spreadLog2 = simpleSpread(ctx, n);               // a handful of items: a warp each
		const int nb = ((n << q.spreadLog2) + kBlock - 1) / kBlock;
		if (deep) CUDA_TRY(ctx, launchOnMap(ctx, traceSimpleKernel<false, kDeepDepth>, nb, kBlock, 0, s, ctx->fast, q, fh));
		else      CUDA_TRY(ctx, launchOnMap(ctx, traceSimpleKernel<false>, nb, kBlock, 0, s, ctx->fast, q, fh));
	} else {
		const int warps = (n + q.chunk - 1) / q.chunk;
		const int nb = (warps + kBlock / 32 - 1) / (kBlock / 32);
		if (deep) CUDA_TRY(ctx, launchOnMap(ctx, traceKernel<false, kDeepDepth>, nb, kBlock, 0, s, ctx->map, q));
		else      CUDA_TRY(ctx, launchOnMap(ctx, traceKernel<false>, nb, kBlock, 0, s, ctx->map, q));
	}
	ctx->launches++;
	CUDA_TRY(ctx, cudaStreamSynchronize(s));
	memcpy(results, h + T::out, N * sizeof(cmgpu_trace_t));
	if (hit_ids) memcpy(hit_ids, h + T::hit, N * 4);
	if (const int st = *(volatile int *)(h + T::status)) {
		if (st & 2) return fail(ctx, CMGPU_ERR_BAD_HANDLE, "CM_ClipHandleToModel: bad handle");
		return fail(ctx, CMGPU_ERR_LIMIT, "traversal stack overflow on a BSP of %d levels", ctx->treeDepth);
	}
	return CMGPU_OK;
}

// ---- big host batches of the hot path: the records come back two ways at once ------------------------------------------
// A host batch waits for its records: 56 bytes per trace over PCIe against 24 on the way in and ~6 ns of GPU time
// (DESIGN.md 4.4).  The walk decides 16 of the 56 bytes; the rest follows from them, the caller's own start / end arrays
// and three tables of the map.  The batch is cut into pieces of `hybridWalkItems` items, and when a piece is queued --
// a few pieces ahead of the one whose results are leaving -- it is given one of two ways home:
//   device: walk, expandRecordsKernel, copy of the 56-byte records into the caller's array, all on the piece's stream;
//   host:   walk, copy of the 16-byte results into pinned memory of the context; worker threads of the calling process
//           write the records into the caller's array (cm_host_records.cpp -- the same float operations in the same order).
// The choice is list scheduling with the measured rates of the two resources: the piece goes to whichever of link and
// workers would be done earlier with everything it has been given so far plus this piece.  Few host cores (eight ranks
// on one box) => few pieces go to the host; neither side waits for the slower one.
struct HybridArrival { hostrec::Pool *pool; int ticket; };
static void CUDART_CB hybridArrived(void *arg) {       // host function behind the copy of a piece's 16-byte results
	const HybridArrival *a = static_cast<const HybridArrival *>(arg);
	hostrec::poolRelease(a->pool, a->ticket);
}

static bool hybridEligible(const cmgpu_ctx *ctx, size_t N, int hull_stride, const int32_t *models, int mask_stride, const float *origins,
                           const float *boxmins) {
	return ctx->hostRecords != 0 && ctx->useFast && ctx->usePool && ctx->twoPhase && ctx->compactByItem && ctx->binning && !ctx->counting &&
	       !hull_stride && !models && !mask_stride && !origins && !boxmins && (ctx->numPatches == 0 || ctx->noCurves) &&
	       N >= ctx->hybridMinItems && N <= ((size_t)1 << 26) && ctx->hybridWalkItems >= (size_t)ctx->binMinItems && !ctx->hostSides.empty();
}

static int traceHostHybrid(cmgpu_ctx *ctx, const size_t N, const float *starts, const float *ends, const float *mins, const float *maxs,
                           const int32_t mask, cmgpu_trace_t *results, int32_t *hit_ids) {
	int rc;
	const size_t v3 = 3 * sizeof(float);
	if (!ctx->hostPool && ctx->hostRecords != 2) {
		int t = ctx->hostThreads;
		if (t < 0) {
			int devs = 1;
			if (cudaGetDeviceCount(&devs) != cudaSuccess || devs < 1) { cudaGetLastError(); devs = 1; }
			const int hw = (int)std::thread::hardware_concurrency();
			t = hw / devs - 1;
			if (t > 32) t = 32;
		}
		if (t < 1) t = 1;
		ctx->hostPool = hostrec::poolCreate(t);
	}
	hostrec::Pool *pool = ctx->hostRecords == 2 ? nullptr : ctx->hostPool;
	if (!ctx->inStream) {
		CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->inStream, cudaStreamNonBlocking));
		CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->outHostStream, cudaStreamNonBlocking));
		CUDA_TRY(ctx, cudaStreamCreateWithFlags(&ctx->outDevStream, cudaStreamNonBlocking));
	}
	const int decouple = ctx->hybridDecouple;
	if (ctx->hostShareFixed >= 0.0) ctx->hostShare = ctx->hostShareFixed;
	if (pool && ctx->hCompactCap < N * 16) {
		if (ctx->hCompact) { CUDA_TRY(ctx, cudaDeviceSynchronize()); cudaFreeHost(ctx->hCompact); ctx->hCompact = nullptr; ctx->hCompactCap = 0; }
		const size_t want = N * 16 + N * 2 + 4096;
		// a host that will not pin that much is served by the plain pipeline (traceHost), not by an error
		if (cudaHostAlloc((void **)&ctx->hCompact, want, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); ctx->hCompact = nullptr; return CMGPU_ERR_UNSUPPORTED_INTERNAL; }
		ctx->hCompactCap = want;
	}
	if ((rc = reserve(ctx, ctx->bStarts, N * v3))) return rc;
	if ((rc = reserve(ctx, ctx->bEnds, N * v3))) return rc;
	if ((rc = reserve(ctx, ctx->bOut, N * sizeof(cmgpu_trace_t)))) return rc;
	if (hit_ids && (rc = reserve(ctx, ctx->bHit, N * 4))) return rc;
	if ((rc = reserve(ctx, ctx->bKeys, N * 4))) return rc;
	if ((rc = reserve(ctx, ctx->bRecs, N * 32))) return rc;
	if ((rc = reserve(ctx, ctx->bCompact, N * 16))) return rc;
	if (ctx->extBusy) {
		CUDA_TRY(ctx, cudaEventSynchronize(ctx->extDone));
		ctx->extBusy = false;
	}
	autoRegister(ctx, starts, N * v3);
	autoRegister(ctx, ends, N * v3);
	autoRegister(ctx, results, N * sizeof(cmgpu_trace_t));
	// Pageable start / end arrays: an upload straight from them is staged by the driver, one blocking copy at a time.  The
	// threads copy a piece ahead into pinned memory instead, and the link uploads from there.
	const bool stage = pool && ctx->stagePageable && N * 24 <= ((size_t)1 << 30) &&
	                   (isPageable(starts) || isPageable(ends) || isPageable((const char *)(starts + 3 * N) - 1) || isPageable((const char *)(ends + 3 * N) - 1));
	if (stage && ctx->hStageCap < N * 24) {
		if (ctx->hStage) { CUDA_TRY(ctx, cudaDeviceSynchronize()); cudaFreeHost(ctx->hStage); ctx->hStage = nullptr; ctx->hStageCap = 0; }
		const size_t want = N * 24 + N * 3 + 4096;
		if (cudaHostAlloc((void **)&ctx->hStage, want, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); ctx->hStage = nullptr; return CMGPU_ERR_UNSUPPORTED_INTERNAL; }
		ctx->hStageCap = want;
	}
	const float *upStarts = stage ? ctx->hStage : starts, *upEnds = stage ? ctx->hStage + 3 * N : ends;
	// a pageable result array is written by the CPU anyway (a copy into it would be staged by the driver, and block)
	const bool allHost = pool && (ctx->hostRecords == 1 || isPageable(results) || isPageable((const char *)(results + N) - 1));

	// pieces: 1/4, 1/2, then whole pieces of hybridWalkItems (the first records should leave early)
	std::vector<size_t> wb;
	wb.push_back(0);
	{
		const size_t piece = ctx->hybridWalkItems;
		size_t step = piece / 4 >= (size_t)ctx->binMinItems ? piece / 4 : piece;
		size_t at = 0;
		while (at < N) {
			at = at + step + step / 2 < N ? ((at + step) & ~(size_t)255) : N;
			wb.push_back(at);
			if (step < piece) step = step * 2 < piece ? step * 2 : piece;
		}
	}
	const int pieces = (int)wb.size() - 1;
	while (ctx->hybridEvents.size() < 4 * (size_t)pieces + 1) {
		cudaEvent_t e;
		CUDA_TRY(ctx, cudaEventCreate(&e));
		ctx->hybridEvents.push_back(e);
	}
	cudaEvent_t *copyStart = ctx->hybridEvents.data(), *pieceDone = ctx->hybridEvents.data() + pieces, *uploaded = ctx->hybridEvents.data() + 2 * pieces;
	cudaEvent_t *walked = ctx->hybridEvents.data() + 3 * pieces;
	const cudaEvent_t evBase = ctx->hybridEvents[4 * (size_t)pieces];

	// no worker outlives the call: on an early return every held job is released (its host function may never run) and waited for
	std::vector<HybridArrival> arrivals((size_t)pieces, HybridArrival{nullptr, 0});
	struct Drain {
		hostrec::Pool *p; std::vector<HybridArrival> *held;
		~Drain() {
			if (!p) return;
			cudaDeviceSynchronize();                         // no host function of this call may fire after `held` is gone
			cudaGetLastError();
			for (auto &a : *held) if (a.pool) hostrec::poolRelease(a.pool, a.ticket);
			hostrec::poolWaitAll(p);
		}
	} drain{pool, &arrivals};
	static const bool traceTimes = getenv("CMGPU_HYBRID_TRACE") != nullptr;
	const auto t0 = std::chrono::steady_clock::now();
	auto sinceMs = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count(); };
	if (pool) hostrec::poolSetEpoch(pool);
	CUDA_TRY(ctx, cudaEventRecord(evBase, ctx->streams[0]));
	std::vector<char> wentHost((size_t)pieces, 0);

	std::vector<size_t> outBytes((size_t)pieces, 0);
	int oldest = 0;
	double backlogBytes = 0.0;          // result bytes queued on the link and not yet seen to have arrived
	unsigned long long hostRecs = 0, devRecs = 0;
	auto harvest = [&](bool block) -> cudaError_t {
		while (oldest < (int)outBytes.size() && outBytes[oldest]) {
			cudaError_t e = block ? cudaEventSynchronize(pieceDone[oldest]) : cudaEventQuery(pieceDone[oldest]);
			if (e == cudaErrorNotReady) return cudaSuccess;
			if (e != cudaSuccess) return e;
			float ms = 0.0f;
			if (outBytes[oldest] >= ((size_t)8 << 20) && cudaEventElapsedTime(&ms, copyStart[oldest], pieceDone[oldest]) == cudaSuccess && ms > 0.0f)
			{	// a copy that shared the link with the other kind measures its share, not the link: follow the best samples
				const double rate = (double)outBytes[oldest] / (1e-3 * ms);
				ctx->d2hRate = rate > ctx->d2hRate ? 0.5 * ctx->d2hRate + 0.5 * rate : 0.95 * ctx->d2hRate + 0.05 * rate;
			}
			backlogBytes -= (double)outBytes[oldest];
			oldest++;
			block = false;
		}
		return cudaSuccess;
	};
	if (pool && ctx->hostShare < 0.0) {
		// ~60 M records per second and thread against a link that moves ~1000 M records per second: 56 / (Rp / Rh + 40)
		const double Rh = 60e6 * hostrec::poolThreads(pool);
		ctx->hostShare = 56.0 / (55e9 / Rh + 40.0);
	}
	const int inFlight = ctx->hybridInFlight;
	hostrec::Tables tables{ctx->hostSides.data(), ctx->hostBrushContents.data()};
	std::vector<int> stageTicket((size_t)pieces, 0);
	auto stagePiece = [&](int c) {
		if (!stage || c >= pieces || stageTicket[c]) return;
		hostrec::Job job{};
		job.starts = starts; job.ends = ends;
		job.stageStarts = ctx->hStage; job.stageEnds = ctx->hStage + 3 * N;
		job.lo = wb[c]; job.hi = wb[c + 1];
		stageTicket[c] = hostrec::poolSubmit(pool, job);
	};
	for (int c = 0; c < pieces; c++) {
		const size_t lo = wb[c], cnt = wb[c + 1] - lo;
		if (stage) {
			stagePiece(c);
			stagePiece(c + 1);                                  // one piece ahead of the upload
			hostrec::poolWaitJob(pool, stageTicket[c]);
		}
		CUDA_TRY(ctx, harvest(false));
		while (c - oldest >= inFlight) CUDA_TRY(ctx, harvest(true));
		bool host = allHost;
		if (pool && !allHost && (ctx->hostRecords < 0 || ctx->hostRecords == 3)) {
			// error diffusion: the threads' share of what has been handed out so far follows hostShare (the first pieces
			// go to them: they are small, and the threads need the head start more than the link does)
			host = (double)(hostRecs + cnt) <= ctx->hostShare * (double)(lo + cnt) + 0.5 * (double)cnt;
			if (traceTimes) fprintf(stderr, "[hybrid] piece %d (%zu items) queued at %.2f ms -> %s; share %.3f, link %.1f GB/s, threads %.0f Mrec/s\n", c, cnt,
			                        sinceMs(), host ? "host" : "device", ctx->hostShare, ctx->d2hRate * 1e-9, hostrec::poolRate(pool) * 1e-6);
		}
		// everything of a piece on its own stream, like the pieces of the round-1 pipeline
		cudaStream_t s = ctx->streams[c % kBinSlots];
		cudaStream_t si = decouple ? ctx->inStream : s;
		cudaStream_t so = decouple == 0 ? s : (decouple == 1 && host ? ctx->outHostStream : ctx->outDevStream);
		CUDA_TRY(ctx, cudaMemcpyAsync((char *)ctx->bStarts.p + lo * v3, (const char *)upStarts + lo * v3, cnt * v3, cudaMemcpyHostToDevice, si));
		CUDA_TRY(ctx, cudaMemcpyAsync((char *)ctx->bEnds.p + lo * v3, (const char *)upEnds + lo * v3, cnt * v3, cudaMemcpyHostToDevice, si));
		if (traceTimes || decouple) CUDA_TRY(ctx, cudaEventRecord(uploaded[c], si));
		if (decouple) CUDA_TRY(ctx, cudaStreamWaitEvent(s, uploaded[c], 0));
		TraceBatch q{};
		q.n = (int)cnt;
		q.starts = (const float *)ctx->bStarts.p + 3 * lo;
		q.ends = (const float *)ctx->bEnds.p + 3 * lo;
		if (mins) { memcpy(q.sharedMins, mins, 12); memcpy(q.sharedMaxs, maxs, 12); }
		q.sharedMask = mask;
		q.out = reinterpret_cast<uint2 *>((cmgpu_trace_t *)ctx->bOut.p + lo);
		q.hitIds = hit_ids && !host ? (int *)ctx->bHit.p + lo : nullptr;
		q.status = ctx->dStatus;
		q.counters = ctx->dCounters;
		wentHost[c] = host;
		if ((rc = launchTrace(ctx, q, s, lo, c % kBinSlots, false, host ? 1 : 2))) return rc;
		if (traceTimes || decouple) CUDA_TRY(ctx, cudaEventRecord(walked[c], s));
		if (decouple) CUDA_TRY(ctx, cudaStreamWaitEvent(so, walked[c], 0));
		CUDA_TRY(ctx, cudaEventRecord(copyStart[c], so));
		if (host) {
			outBytes[c] = cnt * 16;
			CUDA_TRY(ctx, cudaMemcpyAsync(ctx->hCompact + lo * 16, (const char *)ctx->bCompact.p + lo * 16, cnt * 16, cudaMemcpyDeviceToHost, so));
			CUDA_TRY(ctx, cudaEventRecord(pieceDone[c], so));
			hostrec::Job job{};
			job.tables = tables;
			job.compact = reinterpret_cast<const uint32_t *>(ctx->hCompact);
			job.starts = starts; job.ends = ends;
			job.records = results; job.hitIds = hit_ids;
			job.lo = lo; job.hi = lo + cnt;
			job.held = true;
			arrivals[c] = HybridArrival{pool, hostrec::poolSubmit(pool, job)};
			// if this fails the job would never run: release it now (the call fails anyway, Drain waits for the workers)
			if (cudaLaunchHostFunc(so, hybridArrived, &arrivals[c]) != cudaSuccess) { hostrec::poolRelease(pool, arrivals[c].ticket); CUDA_TRY(ctx, cudaGetLastError()); return fail(ctx, CMGPU_ERR_CUDA, "cudaLaunchHostFunc failed"); }
			hostRecs += cnt;
		} else {
			outBytes[c] = cnt * sizeof(cmgpu_trace_t);
			if (hit_ids) CUDA_TRY(ctx, cudaMemcpyAsync(hit_ids + lo, (int *)ctx->bHit.p + lo, cnt * 4, cudaMemcpyDeviceToHost, so));
			CUDA_TRY(ctx, cudaMemcpyAsync(results + lo, (cmgpu_trace_t *)ctx->bOut.p + lo, cnt * sizeof(cmgpu_trace_t), cudaMemcpyDeviceToHost, so));
			CUDA_TRY(ctx, cudaEventRecord(pieceDone[c], so));
			devRecs += cnt;
		}
		backlogBytes += (double)outBytes[c];
	}
	while (oldest < pieces) CUDA_TRY(ctx, harvest(true));
	for (auto &s : ctx->streams) CUDA_TRY(ctx, cudaStreamSynchronize(s));
	if (decouple) for (cudaStream_t st : {ctx->inStream, ctx->outHostStream, ctx->outDevStream}) CUDA_TRY(ctx, cudaStreamSynchronize(st));
	if (traceTimes) fprintf(stderr, "[hybrid] last copy done at %.2f ms\n", sinceMs());
	if (pool) hostrec::poolWaitAll(pool);
	if (traceTimes) {
		fprintf(stderr, "[hybrid] threads done at %.2f ms; %llu records by the host, %llu by the device\n", sinceMs(), hostRecs, devRecs);
		for (int c = 0; c < pieces; c++) {
			float a = 0, b = 0, d = 0;
			cudaEventElapsedTime(&a, evBase, uploaded[c]); cudaEventElapsedTime(&b, evBase, walked[c]); cudaEventElapsedTime(&d, evBase, pieceDone[c]);
			fprintf(stderr, "[hybrid] piece %d %s: uploaded %.2f, walked%s %.2f, results on the host %.2f ms\n", c, wentHost[c] ? "host  " : "device", a, wentHost[c] ? "" : " + expanded", b, d);
		}
	}
	drain.p = nullptr;
	if (pool && !allHost && (ctx->hostRecords < 0 || ctx->hostRecords == 3) && ctx->hostShareFixed < 0.0) {
		const double tHost = hostRecs ? hostrec::poolLastDoneMs(pool) : 0.0;
		double tLink = 0.0;
		for (int c = 0; c < pieces; c++) {
			float ms = 0.0f;
			if (!wentHost[c] && cudaEventElapsedTime(&ms, evBase, pieceDone[c]) == cudaSuccess && ms > tLink) tLink = ms;
		}
		const double total = tHost > tLink ? tHost : tLink;
		if (total > 0.0) {
			// The threads are the tail: they got too much.  Otherwise they should be busy ~72 % of the call (measured: beyond
			// that every fourth call ends with the threads as the tail) -- their last piece arrives when the link delivers
			// it, so finishing together says nothing about whether they could take more.
			const double util = hostrec::poolBusyMs(pool) / total;
			double step;
			if (tHost > 1.03 * tLink && tLink > 0.0) step = 0.7 * ctx->hostShare * (tLink - tHost) / total;
			else step = 0.5 * (0.72 - util) * (ctx->hostShare > 0.1 ? ctx->hostShare : 0.1);
			if (!hostRecs) step = 0.05;                      // nothing went to the threads: try a little
			if (step > 0.05) step = 0.05;
			if (step < -0.08) step = -0.08;
			ctx->hostShare += step;
			if (ctx->hostShare < 0.0) ctx->hostShare = 0.0;
			if (ctx->hostShare > 1.0) ctx->hostShare = 1.0;
		}
		if (traceTimes) fprintf(stderr, "[hybrid] link done at %.2f ms, threads at %.2f ms (busy %.2f ms): share -> %.3f\n", tLink, tHost, hostrec::poolBusyMs(pool), ctx->hostShare);
	}
	ctx->hostRecordsWritten = hostRecs;
	ctx->deviceRecordsWritten = devRecs;
	return cmgpu_check_status(ctx, ctx->streams[0]);
}

static int traceHost(cmgpu_ctx *ctx, int n, const float *starts, const float *ends,
                     const float *mins, const float *maxs, int hull_stride,
                     const int32_t *models, const int32_t *masks, int mask_stride,
                     const float *origins, const float *angles,
                     const float *boxmins, const float *boxmaxs,
                     cmgpu_trace_t *results, int32_t *hit_ids) {
	if (!ctx) return CMGPU_ERR_INVALID;
	if (n < 0 || (n > 0 && (!starts || !ends || !masks || !results)))
		return fail(ctx, CMGPU_ERR_INVALID, "trace batch: bad argument");
	if ((mins == nullptr) != (maxs == nullptr)) return fail(ctx, CMGPU_ERR_INVALID, "mins and maxs must both be given or both be NULL");
	if (origins && !angles) return fail(ctx, CMGPU_ERR_INVALID, "origins without angles");
	if (!ctx->hasMap) return fail(ctx, CMGPU_ERR_NO_MAP, "no map uploaded");
	if (n == 0) return CMGPU_OK;
	int rc = checkHandles(ctx, n, models);
	if (rc) return rc;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	hull_stride = hull_stride ? 1 : 0;
	mask_stride = mask_stride ? 1 : 0;
	if (!boxmins || !boxmaxs) boxmins = boxmaxs = nullptr;
	if (!mins) hull_stride = 0;
	if (n <= kTinyMax && ctx->tinyCalls && ctx->tinyHost && !ctx->counting)
		return traceHostTiny(ctx, n, starts, ends, mins, maxs, hull_stride, models, masks, mask_stride, origins, angles, boxmins, boxmaxs,
		                     results, hit_ids);

	const size_t N = (size_t)n;
	const size_t v3 = 3 * sizeof(float);
	// Big batches of the hot path: the pipeline that shares the records between the link and the host's threads, unless
	// the plain one has proven faster on this machine (slow or busy host cores).  Both are timed -- the shared one over
	// its first four calls (its share is still settling), the plain one over two; the first call of each (allocations,
	// thread start) does not count; the score is a running mean that halves the past with every call; the loser is tried again every 32nd call.
	int timedPipe = -1;
	const auto tPipe = std::chrono::steady_clock::now();
	if (hybridEligible(ctx, N, hull_stride, models, mask_stride, origins, boxmins)) {
		int pipe = 1;
		if (ctx->hostRecords < 0 && !isPageable(results)) {
			unsigned *calls = ctx->pipeCalls;
			if (calls[1] < 4) pipe = 1;
			else if (calls[0] < 2) pipe = 0;
			else {
				pipe = ctx->pipeNs[1] <= ctx->pipeNs[0] ? 1 : 0;
				if ((calls[0] + calls[1]) % 32 == 0) pipe ^= 1;
			}
			timedPipe = pipe;
			if (getenv("CMGPU_HYBRID_TRACE")) fprintf(stderr, "[pipe] %s pipeline; running means: shared %.3f ns/item (%u calls), plain %.3f (%u)\n", pipe ? "shared" : "plain",
			                                          ctx->pipeNs[1], calls[1], ctx->pipeNs[0], calls[0]);
		}
		if (pipe == 1 && !ctx->hybridUnavailable) {
			rc = traceHostHybrid(ctx, N, starts, ends, mins, maxs, masks[0], results, hit_ids);
			if (rc != CMGPU_ERR_UNSUPPORTED_INTERNAL) {
				if (timedPipe == 1 && rc == CMGPU_OK) {
					const double ns = std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now() - tPipe).count() / (double)N;
					if (ctx->pipeCalls[1]++ >= 1) ctx->pipeNs[1] = ctx->pipeCalls[1] == 2 ? ns : 0.5 * ctx->pipeNs[1] + 0.5 * ns;
				}
				return rc;
			}
			ctx->hybridUnavailable = true;               // the staging memory could not be pinned: the plain pipeline from here on
			timedPipe = -1;
		} else if (pipe == 1) timedPipe = -1;
	}
	if ((rc = reserve(ctx, ctx->bStarts, N * v3))) return rc;
	if ((rc = reserve(ctx, ctx->bEnds, N * v3))) return rc;
	if (!mins) hull_stride = 0;
	if (hull_stride) {
		if ((rc = reserve(ctx, ctx->bMins, N * v3))) return rc;
		if ((rc = reserve(ctx, ctx->bMaxs, N * v3))) return rc;
	}
	if (models && (rc = reserve(ctx, ctx->bModels, N * 4))) return rc;
	if (mask_stride && (rc = reserve(ctx, ctx->bMasks, N * 4))) return rc;
	if (origins) {
		if ((rc = reserve(ctx, ctx->bOrigins, N * v3))) return rc;
		if ((rc = reserve(ctx, ctx->bRot, N * 9 * sizeof(float)))) return rc;
		if ((rc = reserve(ctx, ctx->bRotIdx, N * 4))) return rc;
		ctx->hostRot.resize(N * 9);
		ctx->hostRotIdx.resize(N);
		cmgpu_rotation_matrices(n, angles, ctx->hostRot.data(), ctx->hostRotIdx.data());
	}
	if (boxmins) {
		if ((rc = reserve(ctx, ctx->bBoxMins, N * v3))) return rc;
		if ((rc = reserve(ctx, ctx->bBoxMaxs, N * v3))) return rc;
	}
	if ((rc = reserve(ctx, ctx->bOut, N * sizeof(cmgpu_trace_t)))) return rc;
	if (hit_ids && (rc = reserve(ctx, ctx->bHit, N * 4))) return rc;
	if (ctx->binning && N >= (size_t)ctx->binMinItems) {
		if ((rc = reserve(ctx, ctx->bKeys, N * 4))) return rc;
		if ((rc = reserve(ctx, ctx->bRecs, N * 32))) return rc;
		if ((rc = reserve(ctx, ctx->bCompact, N * 16))) return rc;
	}

	if (ctx->extBusy) {                              // an earlier device-pointer call may still use the shared workspace
		CUDA_TRY(ctx, cudaEventSynchronize(ctx->extDone));
		ctx->extBusy = false;
	}
	autoRegister(ctx, starts, N * v3);
	autoRegister(ctx, ends, N * v3);
	autoRegister(ctx, results, N * sizeof(cmgpu_trace_t));
	int chunks = (int)((N + kMinChunkItems - 1) / kMinChunkItems);
	if (chunks > ctx->pipeChunks) chunks = ctx->pipeChunks;
	if (chunks < 1) chunks = 1;
	// Piece boundaries: equal pieces, except that the first ones grow from an eighth of a piece -- the result
	// copies (the bottleneck on PCIe) cannot start before the first piece has been uploaded and traced, so that
	// piece is kept small.
	std::vector<size_t> bounds;
	bounds.push_back(0);
	{
		const size_t piece = (N + (size_t)chunks - 1) / (size_t)chunks;
		size_t first = piece / 8;
		if (first < kMinChunkItems) first = kMinChunkItems;
		size_t at = 0, step = first < piece && ctx->pipeRamp ? first : piece;
		while (at < N) {
			at = at + step < N ? ((at + step) & ~(size_t)255) : N;   // multiples of 256 items: every piece's rays and records start
			                                                          // on 1 KB boundaries (unaligned pieces copy 8 % slower over PCIe)
			bounds.push_back(at);
			if (step < piece) step = step * 2 < piece ? step * 2 : piece;
		}
	}
	chunks = (int)bounds.size() - 1;
	for (int c = 0; c < chunks; c++) {
		const size_t lo = bounds[c], hi = bounds[c + 1];
		const size_t cnt = hi - lo;
		if (!cnt) continue;
		cudaStream_t s = ctx->streams[c % kBinSlots];
#define H2D(buf, src, elem) CUDA_TRY(ctx, cudaMemcpyAsync((char *)(buf).p + lo * (elem), (const char *)(src) + lo * (elem), cnt * (elem), cudaMemcpyHostToDevice, s))
		H2D(ctx->bStarts, starts, v3);
		H2D(ctx->bEnds, ends, v3);
		if (hull_stride) { H2D(ctx->bMins, mins, v3); H2D(ctx->bMaxs, maxs, v3); }
		if (models) H2D(ctx->bModels, models, 4);
		if (mask_stride) H2D(ctx->bMasks, masks, 4);
		if (origins) {
			H2D(ctx->bOrigins, origins, v3);
			H2D(ctx->bRot, ctx->hostRot.data(), 9 * sizeof(float));
			H2D(ctx->bRotIdx, ctx->hostRotIdx.data(), 4);
		}
		if (boxmins) { H2D(ctx->bBoxMins, boxmins, v3); H2D(ctx->bBoxMaxs, boxmaxs, v3); }
#undef H2D
		TraceBatch q{};
		q.n = (int)cnt;
		q.starts = (const float *)ctx->bStarts.p + 3 * lo;
		q.ends = (const float *)ctx->bEnds.p + 3 * lo;
		if (hull_stride) {
			q.mins = (const float *)ctx->bMins.p + 3 * lo;
			q.maxs = (const float *)ctx->bMaxs.p + 3 * lo;
		} else if (mins) {
			memcpy(q.sharedMins, mins, 12);
			memcpy(q.sharedMaxs, maxs, 12);
		}
		q.models = models ? (const int *)ctx->bModels.p + lo : nullptr;
		if (mask_stride) q.masks = (const int *)ctx->bMasks.p + lo; else q.sharedMask = masks[0];
		if (origins) {
			q.origins = (const float *)ctx->bOrigins.p + 3 * lo;
			// rot_index values are absolute item numbers: keep the base pointer unshifted
			q.rotations = (const float *)ctx->bRot.p;
			q.rotIndex = (const int *)ctx->bRotIdx.p + lo;
		}
		if (boxmins) {
			q.boxmins = (const float *)ctx->bBoxMins.p + 3 * lo;
			q.boxmaxs = (const float *)ctx->bBoxMaxs.p + 3 * lo;
		}
		q.out = reinterpret_cast<uint2 *>((cmgpu_trace_t *)ctx->bOut.p + lo);
		q.hitIds = hit_ids ? (int *)ctx->bHit.p + lo : nullptr;
		q.status = ctx->dStatus;
		q.counters = ctx->dCounters;
		if ((rc = launchTrace(ctx, q, s, lo, c % kBinSlots))) return rc;
		CUDA_TRY(ctx, cudaMemcpyAsync(results + lo, (cmgpu_trace_t *)ctx->bOut.p + lo, cnt * sizeof(cmgpu_trace_t), cudaMemcpyDeviceToHost, s));
		if (hit_ids) CUDA_TRY(ctx, cudaMemcpyAsync(hit_ids + lo, (int *)ctx->bHit.p + lo, cnt * 4, cudaMemcpyDeviceToHost, s));
	}
	for (auto &s : ctx->streams) CUDA_TRY(ctx, cudaStreamSynchronize(s));
	rc = cmgpu_check_status(ctx, ctx->streams[0]);
	if (timedPipe == 0 && rc == CMGPU_OK) {
		const double ns = std::chrono::duration<double, std::nano>(std::chrono::steady_clock::now() - tPipe).count() / (double)N;
		if (ctx->pipeCalls[0]++ >= 1) ctx->pipeNs[0] = ctx->pipeCalls[0] == 2 ? ns : 0.5 * ctx->pipeNs[0] + 0.5 * ns;
	}
	return rc;
}

int cmgpu_box_trace_batch(cmgpu_ctx *ctx, int n, const float *starts, const float *ends,
                          const float *mins, const float *maxs, int hull_stride,
                          const int32_t *models, const int32_t *masks, int mask_stride,
                          const float *boxmins, const float *boxmaxs,
                          cmgpu_trace_t *results, int32_t *hit_ids) {
	return traceHost(ctx, n, starts, ends, mins, maxs, hull_stride, models, masks, mask_stride,
	                 nullptr, nullptr, boxmins, boxmaxs, results, hit_ids);
}

int cmgpu_transformed_box_trace_batch(cmgpu_ctx *ctx, int n, const float *starts, const float *ends,
                                      const float *mins, const float *maxs, int hull_stride,
                                      const int32_t *models, const int32_t *masks, int mask_stride,
                                      const float *origins, const float *angles,
                                      const float *boxmins, const float *boxmaxs,
                                      cmgpu_trace_t *results, int32_t *hit_ids) {
	if (ctx && (!origins || !angles)) return fail(ctx, CMGPU_ERR_INVALID, "transformed trace needs origins and angles");
	if (ctx && n > 0 && !models) return fail(ctx, CMGPU_ERR_INVALID, "transformed trace needs models");
	return traceHost(ctx, n, starts, ends, mins, maxs, hull_stride, models, masks, mask_stride,
	                 origins, angles, boxmins, boxmaxs, results, hit_ids);
}

int cmgpu_point_contents_batch(cmgpu_ctx *ctx, int n, const float *points,
                               const int32_t *models, const float *origins, const float *angles,
                               const float *boxmins, const float *boxmaxs, int32_t *contents) {
	if (!ctx) return CMGPU_ERR_INVALID;
	if (n < 0 || (n > 0 && (!points || !contents))) return fail(ctx, CMGPU_ERR_INVALID, "point contents batch: bad argument");
	if (origins && !angles) return fail(ctx, CMGPU_ERR_INVALID, "origins without angles");
	if (!ctx->hasMap) return fail(ctx, CMGPU_ERR_NO_MAP, "no map uploaded");
	if (n == 0) return CMGPU_OK;
	int rc = checkHandles(ctx, n, models);
	if (rc) return rc;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	if (!boxmins || !boxmaxs) boxmins = boxmaxs = nullptr;
	const size_t N = (size_t)n, v3 = 3 * sizeof(float);
	if ((rc = reserve(ctx, ctx->bStarts, N * v3))) return rc;
	if (models && (rc = reserve(ctx, ctx->bModels, N * 4))) return rc;
	if (origins) {
		if ((rc = reserve(ctx, ctx->bOrigins, N * v3))) return rc;
		if ((rc = reserve(ctx, ctx->bRot, N * 9 * sizeof(float)))) return rc;
		if ((rc = reserve(ctx, ctx->bRotIdx, N * 4))) return rc;
		ctx->hostRot.resize(N * 9);
		ctx->hostRotIdx.resize(N);
		cmgpu_rotation_matrices(n, angles, ctx->hostRot.data(), ctx->hostRotIdx.data());
	}
	if (boxmins) {
		if ((rc = reserve(ctx, ctx->bBoxMins, N * v3))) return rc;
		if ((rc = reserve(ctx, ctx->bBoxMaxs, N * v3))) return rc;
	}
	if ((rc = reserve(ctx, ctx->bHit, N * 4))) return rc;
	cudaStream_t s = ctx->streams[0];
	CUDA_TRY(ctx, cudaMemcpyAsync(ctx->bStarts.p, points, N * v3, cudaMemcpyHostToDevice, s));
	if (models) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->bModels.p, models, N * 4, cudaMemcpyHostToDevice, s));
	if (origins) {
		CUDA_TRY(ctx, cudaMemcpyAsync(ctx->bOrigins.p, origins, N * v3, cudaMemcpyHostToDevice, s));
		CUDA_TRY(ctx, cudaMemcpyAsync(ctx->bRot.p, ctx->hostRot.data(), N * 9 * sizeof(float), cudaMemcpyHostToDevice, s));
		CUDA_TRY(ctx, cudaMemcpyAsync(ctx->bRotIdx.p, ctx->hostRotIdx.data(), N * 4, cudaMemcpyHostToDevice, s));
	}
	if (boxmins) {
		CUDA_TRY(ctx, cudaMemcpyAsync(ctx->bBoxMins.p, boxmins, N * v3, cudaMemcpyHostToDevice, s));
		CUDA_TRY(ctx, cudaMemcpyAsync(ctx->bBoxMaxs.p, boxmaxs, N * v3, cudaMemcpyHostToDevice, s));
	}
	rc = cmgpu_point_contents_batch_device(ctx, n, (const float *)ctx->bStarts.p,
	                                       models ? (const int32_t *)ctx->bModels.p : nullptr,
	                                       origins ? (const float *)ctx->bOrigins.p : nullptr,
	                                       origins ? (const float *)ctx->bRot.p : nullptr,
	                                       origins ? (const int32_t *)ctx->bRotIdx.p : nullptr,
	                                       boxmins ? (const float *)ctx->bBoxMins.p : nullptr,
	                                       boxmins ? (const float *)ctx->bBoxMaxs.p : nullptr,
	                                       (int32_t *)ctx->bHit.p, s);
	if (rc) return rc;
	CUDA_TRY(ctx, cudaMemcpyAsync(contents, ctx->bHit.p, N * 4, cudaMemcpyDeviceToHost, s));
	CUDA_TRY(ctx, cudaStreamSynchronize(s));
	return CMGPU_OK;
}

}  // extern "C"

// ---- SV_Trace / SV_PointContents on the device (cm_sv_world.cuh) ---------------------------

namespace {

struct SectorNode { int axis; float dist; int child[2]; int parent, side, end; };

// SV_CreateworldSector (sv_world.c:81-113): depth 4, split the longer of x and y at its middle; children[0] is the
// upper half.  Nodes are numbered in creation order, which is depth first with children[0] before children[1].
int buildSectors(std::vector<SectorNode> &nodes, int depth, const float *mins, const float *maxs, int parent, int side) {
	const int me = (int)nodes.size();
	nodes.push_back(SectorNode{-1, 0.0f, {-1, -1}, parent, side, 0});
	if (depth < 4) {
		const float sx = maxs[0] - mins[0], sy = maxs[1] - mins[1];
		const int axis = sx > sy ? 0 : 1;
		const float dist = (float)(0.5 * (double)(maxs[axis] + mins[axis]));
		nodes[me].axis = axis;
		nodes[me].dist = dist;
		float mins1[3], maxs1[3], mins2[3], maxs2[3];
		memcpy(mins1, mins, 12); memcpy(mins2, mins, 12); memcpy(maxs1, maxs, 12); memcpy(maxs2, maxs, 12);
		maxs1[axis] = mins2[axis] = dist;
		const int c0 = buildSectors(nodes, depth + 1, mins2, maxs2, me, 0);
		const int c1 = buildSectors(nodes, depth + 1, mins1, maxs1, me, 1);
		nodes[me].child[0] = c0;
		nodes[me].child[1] = c1;
	}
	nodes[me].end = (int)nodes.size();
	return me;
}

template <typename T>
T *carve(std::vector<char> &host, size_t count, size_t &offset) {
	offset = (host.size() + 255) & ~(size_t)255;
	host.resize(offset + count * sizeof(T));
	return reinterpret_cast<T *>(host.data() + offset);
}

// exclusive prefix sum of counts[0..m) in place (the binning scan kernels)
int scanOffsets(cmgpu_ctx *ctx, unsigned *counts, int m, cudaStream_t s) {
	const int tiles = (m + kScanTile - 1) / kScanTile;
	int rc = reserve(ctx, ctx->bSvTiles, (size_t)tiles * 4);
	if (rc) return rc;
	unsigned *tileSums = (unsigned *)ctx->bSvTiles.p;
	binTileSumKernel<<<tiles, 256, 0, s>>>(counts, m, tileSums);
	binTileScanKernel<<<1, 256, 0, s>>>(tileSums, tiles);
	binScanKernel<<<tiles, 256, 0, s>>>(counts, m, tileSums);
	ctx->launches += 3;
	CUDA_TRY(ctx, cudaGetLastError());
	return CMGPU_OK;
}

int readTotal(cmgpu_ctx *ctx, const unsigned *offsets, int n, cudaStream_t s, unsigned *total) {
	if (!ctx->hSvTotal) CUDA_TRY(ctx, cudaMallocHost((void **)&ctx->hSvTotal, sizeof(unsigned)));
	CUDA_TRY(ctx, cudaMemcpyAsync(ctx->hSvTotal, offsets + n, sizeof(unsigned), cudaMemcpyDeviceToHost, s));
	CUDA_TRY(ctx, cudaStreamSynchronize(s));
	*total = *ctx->hSvTotal;
	if (*total > 0x7fffff00u) return fail(ctx, CMGPU_ERR_LIMIT, "SV batch: %u entity clips do not fit one launch", *total);
	return CMGPU_OK;
}

}  // namespace

extern "C" {

int cmgpu_sv_set_entities(cmgpu_ctx *ctx, int n, const cmgpu_sv_entity *ents, const int32_t *ownerNums, int numGentities) {
	if (!ctx) return CMGPU_ERR_INVALID;
	if (n < 0 || (n > 0 && !ents) || numGentities < 0 || (numGentities > 0 && !ownerNums))
		return fail(ctx, CMGPU_ERR_INVALID, "cmgpu_sv_set_entities: bad argument");
	if (!ctx->hasMap) return fail(ctx, CMGPU_ERR_NO_MAP, "no map uploaded");
	for (int i = 0; i < n; i++) {
		if (ents[i].bmodel && (ents[i].modelindex < 0 || ents[i].modelindex >= ctx->numModels))
			return fail(ctx, CMGPU_ERR_BAD_HANDLE, "CM_InlineModel: bad number %i (entity %d)", ents[i].modelindex, ents[i].number);
	}
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	std::vector<SectorNode> nodes;
	buildSectors(nodes, 0, ctx->worldBounds, ctx->worldBounds + 3, -1, -1);
	// SV_LinkEntity's descent (sv_world.c:308-320): the first node whose plane the box straddles, or a leaf
	std::vector<int> nodeOf((size_t)n), order((size_t)n);
	std::vector<int> perNode(nodes.size() + 1, 0);
	for (int i = 0; i < n; i++) {
		int k = 0;
		while (nodes[k].axis != -1) {
			if (ents[i].absmin[nodes[k].axis] > nodes[k].dist) k = nodes[k].child[0];
			else if (ents[i].absmax[nodes[k].axis] < nodes[k].dist) k = nodes[k].child[1];
			else break;
		}
		nodeOf[i] = k;
		perNode[k + 1]++;
	}
	for (size_t k = 0; k < nodes.size(); k++) perNode[k + 1] += perNode[k];
	{
		std::vector<int> next(perNode.begin(), perNode.end() - 1);
		for (int i = 0; i < n; i++) order[next[nodeOf[i]]++] = i;          // stable: the caller's order inside a node
	}
	const size_t E = (size_t)n, G = (size_t)numGentities;
	std::vector<char> host;
	// the two grids (cm_sv_world.cuh): per cell, the entities whose box reaches it
	const int gridCells[2] = {32, 8};
	float gridLo[2], gridInv[2][2];
	std::vector<int> cellStart[2], cellEnts[2];
	std::vector<int> firstCell((size_t)n, 0);
	for (int a = 0; a < 2; a++) gridLo[a] = ctx->worldBounds[a];
	for (int l = 0; l < 2; l++) {
		const int G = gridCells[l];
		for (int a = 0; a < 2; a++) {
			const float size = ctx->worldBounds[3 + a] - ctx->worldBounds[a];
			gridInv[l][a] = size > 0.0f ? (float)G / size : 0.0f;
		}
		std::vector<int> count((size_t)G * G + 1, 0);
		auto range = [&](const cmgpu_sv_entity &e, int *r) {
			r[0] = svCellOf(e.absmin[0], gridLo[0], gridInv[l][0], G); r[1] = svCellOf(e.absmax[0], gridLo[0], gridInv[l][0], G);
			r[2] = svCellOf(e.absmin[1], gridLo[1], gridInv[l][1], G); r[3] = svCellOf(e.absmax[1], gridLo[1], gridInv[l][1], G);
		};
		int r[4];
		for (int j = 0; j < n; j++) {
			range(ents[order[j]], r);
			for (int y = r[2]; y <= r[3]; y++) for (int x = r[0]; x <= r[1]; x++) count[(size_t)y * G + x + 1]++;
		}
		for (size_t c = 0; c < (size_t)G * G; c++) count[c + 1] += count[c];
		cellStart[l] = count;
		cellEnts[l].assign((size_t)count.back() + 1, 0);
		std::vector<int> next(count.begin(), count.end() - 1);
		for (int j = 0; j < n; j++) {
			range(ents[order[j]], r);
			firstCell[j] |= (r[0] | r[2] << 8) << (16 * l);
			for (int y = r[2]; y <= r[3]; y++) for (int x = r[0]; x <= r[1]; x++) cellEnts[l][next[(size_t)y * G + x]++] = j;
		}
	}
	size_t oNodes, oLo, oHi, oInfo, oOrg, oMins, oMaxs, oRot, oOwner, oCellStart[2], oCellEnts[2];
	carve<int4>(host, 2 * nodes.size(), oNodes);
	carve<float4>(host, E + 1, oLo);
	carve<float4>(host, E + 1, oHi);
	carve<int4>(host, E + 1, oInfo);
	carve<float>(host, 3 * E + 3, oOrg);
	carve<float>(host, 3 * E + 3, oMins);
	carve<float>(host, 3 * E + 3, oMaxs);
	carve<float>(host, 9 * E + 9, oRot);
	carve<int>(host, G + 1, oOwner);
	for (int l = 0; l < 2; l++) {
		int *p0 = carve<int>(host, cellStart[l].size(), oCellStart[l]);
		memcpy(p0, cellStart[l].data(), 4 * cellStart[l].size());
		int *p1 = carve<int>(host, cellEnts[l].size(), oCellEnts[l]);
		memcpy(p1, cellEnts[l].data(), 4 * cellEnts[l].size());
	}
	int4 *hNodes = reinterpret_cast<int4 *>(host.data() + oNodes);
	for (size_t k = 0; k < nodes.size(); k++) {
		const SectorNode &nd = nodes[k];
		int axis = 0, distBits = 0;
		if (nd.parent >= 0) { axis = nodes[nd.parent].axis; memcpy(&distBits, &nodes[nd.parent].dist, 4); }
		hNodes[2 * k] = make_int4(axis, distBits, nd.side, nd.end);
		hNodes[2 * k + 1] = make_int4(perNode[k], perNode[k + 1], 0, 0);
	}
	std::vector<float> angles(3 * E + 3, 0.0f);
	std::vector<int32_t> rotIdx(E + 1, -1);
	for (size_t j = 0; j < E; j++) {
		const cmgpu_sv_entity &e = ents[order[j]];
		float4 lo = make_float4(e.absmin[0], e.absmin[1], e.absmin[2], 0.0f), hi = make_float4(e.absmax[0], e.absmax[1], e.absmax[2], 0.0f);
		memcpy(&lo.w, &e.ownerNum, 4);
		memcpy(&hi.w, &e.contents, 4);
		reinterpret_cast<float4 *>(host.data() + oLo)[j] = lo;
		reinterpret_cast<float4 *>(host.data() + oHi)[j] = hi;
		memcpy(host.data() + oOrg + 12 * j, e.origin, 12);
		memcpy(host.data() + oMins + 12 * j, e.mins, 12);
		memcpy(host.data() + oMaxs + 12 * j, e.maxs, 12);
		memcpy(&angles[3 * j], e.angles, 12);
	}
	// rotation matrices with the host's libm, as for every transformed trace (cm_trace.c:776-779)
	cmgpu_rotation_matrices(n, angles.data(), reinterpret_cast<float *>(host.data() + oRot), rotIdx.data());
	for (size_t j = 0; j < E; j++) {
		const cmgpu_sv_entity &e = ents[order[j]];
		// SV_ClipHandleForEntity (sv_world.c:34-42)
		reinterpret_cast<int4 *>(host.data() + oInfo)[j] = make_int4(e.number, e.bmodel ? e.modelindex : CMGPU_BOX_MODEL_HANDLE, rotIdx[j], firstCell[j]);
	}
	if (G) memcpy(host.data() + oOwner, ownerNums, 4 * G);
	CUDA_TRY(ctx, cudaDeviceSynchronize());                                 // nothing may still be reading the old snapshot
	int rc = reserve(ctx, ctx->bSvWorld, host.size());
	if (rc) return rc;
	CUDA_TRY(ctx, cudaMemcpy(ctx->bSvWorld.p, host.data(), host.size(), cudaMemcpyHostToDevice));
	const char *base = (const char *)ctx->bSvWorld.p;
	SvWorld &w = ctx->sv;
	w.nodes = (const int4 *)(base + oNodes);
	w.entLo = (const float4 *)(base + oLo);
	w.entHi = (const float4 *)(base + oHi);
	w.entInfo = (const int4 *)(base + oInfo);
	w.entOrigin = (const float *)(base + oOrg);
	w.entMins = (const float *)(base + oMins);
	w.entMaxs = (const float *)(base + oMaxs);
	w.entRot = (const float *)(base + oRot);
	w.ownerByNum = (const int *)(base + oOwner);
	w.numNodes = (int)nodes.size();
	w.numEnts = n;
	w.numGentities = numGentities;
	for (int l = 0; l < 2; l++) {
		w.cellStart[l] = (const int *)(base + oCellStart[l]);
		w.cellEnts[l] = (const int *)(base + oCellEnts[l]);
		w.gridCells[l] = gridCells[l];
		w.gridInv[l][0] = gridInv[l][0]; w.gridInv[l][1] = gridInv[l][1];
	}
	w.gridLo[0] = gridLo[0]; w.gridLo[1] = gridLo[1];
	ctx->hasSvWorld = true;
	return CMGPU_OK;
}

unsigned long long cmgpu_sv_last_candidates(const cmgpu_ctx *ctx) { return ctx ? ctx->svCandidates : 0; }

int cmgpu_sv_trace_batch_device(cmgpu_ctx *ctx, int n, const float *starts, const float *ends,
                                const float *mins, const float *maxs, int hull_stride,
                                const int32_t *passEntityNums, int pass_stride,
                                const int32_t *contentmasks, int mask_stride, cmgpu_trace_t *results, void *stream) {
	if (!ctx) return CMGPU_ERR_INVALID;
	if (n < 0 || (n > 0 && (!starts || !ends || !passEntityNums || !contentmasks || !results)))
		return fail(ctx, CMGPU_ERR_INVALID, "cmgpu_sv_trace_batch_device: bad argument");
	if (!ctx->hasMap) return fail(ctx, CMGPU_ERR_NO_MAP, "no map uploaded");
	if (!ctx->hasSvWorld) return fail(ctx, CMGPU_ERR_INVALID, "cmgpu_sv_set_entities has not been called for this map");
	if ((mins == nullptr) != (maxs == nullptr)) return fail(ctx, CMGPU_ERR_INVALID, "mins and maxs must both be given or both be NULL");
	if (hull_stride && !mins) hull_stride = 0;
	ctx->svCandidates = 0;
	if (n == 0) return CMGPU_OK;
	cudaStream_t s = (cudaStream_t)stream;
	// 1. clip to world (sv_world.c:576)
	int rc = cmgpu_box_trace_batch_device(ctx, n, starts, ends, mins, maxs, hull_stride, nullptr, contentmasks, mask_stride,
	                                      nullptr, nullptr, nullptr, nullptr, nullptr, results, nullptr, stream);
	if (rc) return rc;
	// 2. count the entities every move has to be clipped against
	const size_t offBytes = (((size_t)n + 1) * 4 + 255) & ~(size_t)255;
	if ((rc = reserve(ctx, ctx->bSvOffsets, offBytes + (size_t)n * sizeof(int4)))) return rc;
	SvMoves q{};
	q.n = n; q.starts = starts; q.ends = ends;
	if (hull_stride) { q.mins = mins; q.maxs = maxs; }
	else if (mins) { memcpy(q.sharedMins, mins, 12); memcpy(q.sharedMaxs, maxs, 12); }
	if (pass_stride) q.passEnts = passEntityNums; else q.sharedPass = passEntityNums[0];
	if (mask_stride) q.masks = contentmasks; else q.sharedMask = contentmasks[0];
	q.results = reinterpret_cast<uint2 *>(results);
	q.offsets = (unsigned *)ctx->bSvOffsets.p;
	q.firstCands = (int4 *)((char *)ctx->bSvOffsets.p + offBytes);
	const int grid = (n + 127) / 128;
	CUDA_TRY(ctx, cudaMemsetAsync(q.offsets + n, 0, 4, s));
	svGatherKernel<false><<<grid, 128, 0, s>>>(ctx->sv, q, SvItems{});
	ctx->launches++;
	CUDA_TRY(ctx, cudaGetLastError());
	if ((rc = scanOffsets(ctx, q.offsets, n + 1, s))) return rc;
	unsigned total = 0;
	if ((rc = readTotal(ctx, q.offsets, n, s, &total))) return rc;
	ctx->svCandidates = total;
	if (total == 0) return CMGPU_OK;
	// 3. one transformed-trace item per (move, entity)
	const size_t T = total, v3 = 12;
	const size_t perItem = 5 * v3 + (hull_stride ? 2 * v3 : 0) + 4 /*model*/ + (mask_stride ? 4 : 0) + 4 /*rot*/ + 4 /*number*/;
	if ((rc = reserve(ctx, ctx->bSvItems, T * perItem + 16 * 256))) return rc;
	if ((rc = reserve(ctx, ctx->bSvCand, T * sizeof(cmgpu_trace_t)))) return rc;
	char *p = (char *)ctx->bSvItems.p;
	auto take = [&](size_t bytes) { char *r = p; p += (bytes + 255) & ~(size_t)255; return r; };
	SvItems it{};
	it.starts = (float *)take(T * v3); it.ends = (float *)take(T * v3);
	if (hull_stride) { it.mins = (float *)take(T * v3); it.maxs = (float *)take(T * v3); }
	it.models = (int *)take(T * 4);
	if (mask_stride) it.masks = (int *)take(T * 4);
	it.origins = (float *)take(T * v3);
	it.rotIndex = (int *)take(T * 4);
	it.boxmins = (float *)take(T * v3); it.boxmaxs = (float *)take(T * v3);
	it.entIndex = (int *)take(T * 4);
	svGatherKernel<true><<<grid, 128, 0, s>>>(ctx->sv, q, it);
	ctx->launches++;
	CUDA_TRY(ctx, cudaGetLastError());
	// 4. clip to the entities (sv_world.c:526-528), all moves in one launch
	const int32_t sharedMask = mask_stride ? 0 : contentmasks[0];
	rc = cmgpu_box_trace_batch_device(ctx, (int)total, it.starts, it.ends,
	                                  hull_stride ? it.mins : mins, hull_stride ? it.maxs : maxs, hull_stride,
	                                  it.models, mask_stride ? it.masks : &sharedMask, mask_stride,
	                                  it.origins, ctx->sv.entRot, it.rotIndex, it.boxmins, it.boxmaxs,
	                                  (cmgpu_trace_t *)ctx->bSvCand.p, nullptr, stream);
	if (rc) return rc;
	// 5. merge in list order
	svMergeKernel<<<grid, 128, 0, s>>>(n, q.results, q.offsets, (const uint2 *)ctx->bSvCand.p, it.entIndex, ctx->sv.entInfo);
	ctx->launches++;
	CUDA_TRY(ctx, cudaGetLastError());
	return extRecord(ctx, s);                    // the merge still reads the offsets and the candidates' records
}

int cmgpu_sv_trace_batch(cmgpu_ctx *ctx, int n, const float *starts, const float *ends,
                         const float *mins, const float *maxs, int hull_stride,
                         const int32_t *passEntityNums, int pass_stride,
                         const int32_t *contentmasks, int mask_stride, cmgpu_trace_t *results) {
	if (!ctx) return CMGPU_ERR_INVALID;
	if (n < 0 || (n > 0 && (!starts || !ends || !passEntityNums || !contentmasks || !results)))
		return fail(ctx, CMGPU_ERR_INVALID, "cmgpu_sv_trace_batch: bad argument");
	if (n == 0) return CMGPU_OK;
	if (hull_stride && !mins) hull_stride = 0;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	const size_t N = (size_t)n, v3 = 12;
	int rc;
	if ((rc = reserve(ctx, ctx->bStarts, N * v3))) return rc;
	if ((rc = reserve(ctx, ctx->bEnds, N * v3))) return rc;
	if (hull_stride) {
		if ((rc = reserve(ctx, ctx->bMins, N * v3))) return rc;
		if ((rc = reserve(ctx, ctx->bMaxs, N * v3))) return rc;
	}
	if (mask_stride && (rc = reserve(ctx, ctx->bMasks, N * 4))) return rc;
	if (pass_stride && (rc = reserve(ctx, ctx->bSvPass, N * 4))) return rc;
	if ((rc = reserve(ctx, ctx->bOut, N * sizeof(cmgpu_trace_t)))) return rc;
	cudaStream_t s = ctx->streams[0];
	CUDA_TRY(ctx, cudaMemcpyAsync(ctx->bStarts.p, starts, N * v3, cudaMemcpyHostToDevice, s));
	CUDA_TRY(ctx, cudaMemcpyAsync(ctx->bEnds.p, ends, N * v3, cudaMemcpyHostToDevice, s));
	if (hull_stride) {
		CUDA_TRY(ctx, cudaMemcpyAsync(ctx->bMins.p, mins, N * v3, cudaMemcpyHostToDevice, s));
		CUDA_TRY(ctx, cudaMemcpyAsync(ctx->bMaxs.p, maxs, N * v3, cudaMemcpyHostToDevice, s));
	}
	if (mask_stride) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->bMasks.p, contentmasks, N * 4, cudaMemcpyHostToDevice, s));
	if (pass_stride) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->bSvPass.p, passEntityNums, N * 4, cudaMemcpyHostToDevice, s));
	rc = cmgpu_sv_trace_batch_device(ctx, n, (const float *)ctx->bStarts.p, (const float *)ctx->bEnds.p,
	                                 hull_stride ? (const float *)ctx->bMins.p : mins, hull_stride ? (const float *)ctx->bMaxs.p : maxs, hull_stride,
	                                 pass_stride ? (const int32_t *)ctx->bSvPass.p : passEntityNums, pass_stride,
	                                 mask_stride ? (const int32_t *)ctx->bMasks.p : contentmasks, mask_stride,
	                                 (cmgpu_trace_t *)ctx->bOut.p, s);
	if (rc) return rc;
	CUDA_TRY(ctx, cudaMemcpyAsync(results, ctx->bOut.p, N * sizeof(cmgpu_trace_t), cudaMemcpyDeviceToHost, s));
	CUDA_TRY(ctx, cudaStreamSynchronize(s));
	return cmgpu_check_status(ctx, s);
}

int cmgpu_sv_point_contents_batch_device(cmgpu_ctx *ctx, int n, const float *points,
                                         const int32_t *passEntityNums, int pass_stride, int32_t *contents, void *stream) {
	if (!ctx) return CMGPU_ERR_INVALID;
	if (n < 0 || (n > 0 && (!points || !passEntityNums || !contents)))
		return fail(ctx, CMGPU_ERR_INVALID, "cmgpu_sv_point_contents_batch_device: bad argument");
	if (!ctx->hasMap) return fail(ctx, CMGPU_ERR_NO_MAP, "no map uploaded");
	if (!ctx->hasSvWorld) return fail(ctx, CMGPU_ERR_INVALID, "cmgpu_sv_set_entities has not been called for this map");
	ctx->svCandidates = 0;
	if (n == 0) return CMGPU_OK;
	cudaStream_t s = (cudaStream_t)stream;
	int rc = extWait(ctx, s);
	if (rc) return rc;
	// base contents from the world (sv_world.c:625)
	rc = cmgpu_point_contents_batch_device(ctx, n, points, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, contents, stream);
	if (rc) return rc;
	if ((rc = reserve(ctx, ctx->bSvOffsets, ((size_t)n + 1) * 4))) return rc;
	unsigned *offsets = (unsigned *)ctx->bSvOffsets.p;
	const int grid = (n + 127) / 128;
	const int *passDev = pass_stride ? passEntityNums : nullptr;
	const int sharedPass = pass_stride ? 0 : passEntityNums[0];
	CUDA_TRY(ctx, cudaMemsetAsync(offsets + n, 0, 4, s));
	svPointGatherKernel<false><<<grid, 128, 0, s>>>(ctx->sv, n, points, passDev, sharedPass, offsets, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
	ctx->launches++;
	CUDA_TRY(ctx, cudaGetLastError());
	if ((rc = scanOffsets(ctx, offsets, n + 1, s))) return rc;
	unsigned total = 0;
	if ((rc = readTotal(ctx, offsets, n, s, &total))) return rc;
	ctx->svCandidates = total;
	if (total == 0) return CMGPU_OK;
	const size_t T = total, v3 = 12;
	if ((rc = reserve(ctx, ctx->bSvItems, T * (4 * v3 + 8) + 8 * 256))) return rc;
	if ((rc = reserve(ctx, ctx->bSvCand, T * 4))) return rc;
	char *p = (char *)ctx->bSvItems.p;
	auto take = [&](size_t bytes) { char *r = p; p += (bytes + 255) & ~(size_t)255; return r; };
	float *itPoints = (float *)take(T * v3), *itOrigins = (float *)take(T * v3);
	float *itBoxMins = (float *)take(T * v3), *itBoxMaxs = (float *)take(T * v3);
	int *itModels = (int *)take(T * 4), *itRot = (int *)take(T * 4);
	svPointGatherKernel<true><<<grid, 128, 0, s>>>(ctx->sv, n, points, passDev, sharedPass, offsets, itPoints, itModels, itOrigins, itRot, itBoxMins, itBoxMaxs);
	ctx->launches++;
	CUDA_TRY(ctx, cudaGetLastError());
	// CM_TransformedPointContents of every candidate (sv_world.c:641)
	rc = cmgpu_point_contents_batch_device(ctx, (int)total, itPoints, itModels, itOrigins, ctx->sv.entRot, itRot, itBoxMins, itBoxMaxs,
	                                       (int32_t *)ctx->bSvCand.p, stream);
	if (rc) return rc;
	svPointMergeKernel<<<grid, 128, 0, s>>>(n, contents, offsets, (const int *)ctx->bSvCand.p);
	ctx->launches++;
	CUDA_TRY(ctx, cudaGetLastError());
	return extRecord(ctx, s);
}

int cmgpu_sv_point_contents_batch(cmgpu_ctx *ctx, int n, const float *points,
                                  const int32_t *passEntityNums, int pass_stride, int32_t *contents) {
	if (!ctx) return CMGPU_ERR_INVALID;
	if (n < 0 || (n > 0 && (!points || !passEntityNums || !contents)))
		return fail(ctx, CMGPU_ERR_INVALID, "cmgpu_sv_point_contents_batch: bad argument");
	if (n == 0) return CMGPU_OK;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	const size_t N = (size_t)n;
	int rc;
	if ((rc = reserve(ctx, ctx->bStarts, N * 12))) return rc;
	if (pass_stride && (rc = reserve(ctx, ctx->bSvPass, N * 4))) return rc;
	if ((rc = reserve(ctx, ctx->bHit, N * 4))) return rc;
	cudaStream_t s = ctx->streams[0];
	CUDA_TRY(ctx, cudaMemcpyAsync(ctx->bStarts.p, points, N * 12, cudaMemcpyHostToDevice, s));
	if (pass_stride) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->bSvPass.p, passEntityNums, N * 4, cudaMemcpyHostToDevice, s));
	rc = cmgpu_sv_point_contents_batch_device(ctx, n, (const float *)ctx->bStarts.p,
	                                          pass_stride ? (const int32_t *)ctx->bSvPass.p : passEntityNums, pass_stride,
	                                          (int32_t *)ctx->bHit.p, s);
	if (rc) return rc;
	CUDA_TRY(ctx, cudaMemcpyAsync(contents, ctx->bHit.p, N * 4, cudaMemcpyDeviceToHost, s));
	CUDA_TRY(ctx, cudaStreamSynchronize(s));
	return CMGPU_OK;
}

// ---- leaf queries (cm_leaf_queries.cuh) ------------------------------------------------------
int cmgpu_point_leafnum_batch_device(cmgpu_ctx *ctx, int n, const float *points, int32_t *leafnums, void *stream) {
	if (!ctx) return CMGPU_ERR_INVALID;
	if (n < 0 || (n > 0 && (!points || !leafnums))) return fail(ctx, CMGPU_ERR_INVALID, "cmgpu_point_leafnum_batch_device: bad argument");
	if (!ctx->hasMap) return fail(ctx, CMGPU_ERR_NO_MAP, "no map uploaded");
	if (n == 0) return CMGPU_OK;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	pointLeafnumKernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(ctx->map, n, points, leafnums);
	ctx->launches++;
	CUDA_TRY(ctx, cudaGetLastError());
	return CMGPU_OK;
}

int cmgpu_box_leafnums_batch_device(cmgpu_ctx *ctx, int n, const float *mins, const float *maxs, int listsize,
                                    int32_t *lists, int32_t *counts, int32_t *lastLeafs, void *stream) {
	if (!ctx) return CMGPU_ERR_INVALID;
	if (n < 0 || listsize < 0 || (n > 0 && (!mins || !maxs || !counts || !lastLeafs || (listsize > 0 && !lists))))
		return fail(ctx, CMGPU_ERR_INVALID, "cmgpu_box_leafnums_batch_device: bad argument");
	if (!ctx->hasMap) return fail(ctx, CMGPU_ERR_NO_MAP, "no map uploaded");
	if (n == 0) return CMGPU_OK;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	LeafListBatch q{n, mins, maxs, listsize, lists, counts, lastLeafs, ctx->leafCluster, ctx->dStatus};
	if (ctx->treeDepth > kMaxDepth) boxLeafnumsKernel<kDeepDepth><<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(ctx->map, q);
	else boxLeafnumsKernel<kMaxDepth><<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(ctx->map, q);
	ctx->launches++;
	CUDA_TRY(ctx, cudaGetLastError());
	return CMGPU_OK;
}

int cmgpu_point_leafnum_batch(cmgpu_ctx *ctx, int n, const float *points, int32_t *leafnums) {
	if (!ctx) return CMGPU_ERR_INVALID;
	if (n < 0 || (n > 0 && (!points || !leafnums))) return fail(ctx, CMGPU_ERR_INVALID, "cmgpu_point_leafnum_batch: bad argument");
	if (n == 0) return CMGPU_OK;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	const size_t N = (size_t)n;
	int rc;
	if ((rc = reserve(ctx, ctx->bStarts, N * 12))) return rc;
	if ((rc = reserve(ctx, ctx->bHit, N * 4))) return rc;
	cudaStream_t s = ctx->streams[0];
	CUDA_TRY(ctx, cudaMemcpyAsync(ctx->bStarts.p, points, N * 12, cudaMemcpyHostToDevice, s));
	if ((rc = cmgpu_point_leafnum_batch_device(ctx, n, (const float *)ctx->bStarts.p, (int32_t *)ctx->bHit.p, s))) return rc;
	CUDA_TRY(ctx, cudaMemcpyAsync(leafnums, ctx->bHit.p, N * 4, cudaMemcpyDeviceToHost, s));
	CUDA_TRY(ctx, cudaStreamSynchronize(s));
	return CMGPU_OK;
}

int cmgpu_box_leafnums_batch(cmgpu_ctx *ctx, int n, const float *mins, const float *maxs, int listsize,
                             int32_t *lists, int32_t *counts, int32_t *lastLeafs) {
	if (!ctx) return CMGPU_ERR_INVALID;
	if (n < 0 || listsize < 0 || (n > 0 && (!mins || !maxs || !counts || !lastLeafs || (listsize > 0 && !lists))))
		return fail(ctx, CMGPU_ERR_INVALID, "cmgpu_box_leafnums_batch: bad argument");
	if (n == 0) return CMGPU_OK;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	const size_t N = (size_t)n, L = N * (size_t)listsize * 4;
	int rc;
	if ((rc = reserve(ctx, ctx->bStarts, N * 12))) return rc;
	if ((rc = reserve(ctx, ctx->bEnds, N * 12))) return rc;
	if ((rc = reserve(ctx, ctx->bLeafLists, L + 16))) return rc;
	if ((rc = reserve(ctx, ctx->bLeafCounts, 2 * N * 4))) return rc;
	cudaStream_t s = ctx->streams[0];
	CUDA_TRY(ctx, cudaMemcpyAsync(ctx->bStarts.p, mins, N * 12, cudaMemcpyHostToDevice, s));
	CUDA_TRY(ctx, cudaMemcpyAsync(ctx->bEnds.p, maxs, N * 12, cudaMemcpyHostToDevice, s));
	if (L) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->bLeafLists.p, lists, L, cudaMemcpyHostToDevice, s));    // entries past a count stay the caller's
	int32_t *dCounts = (int32_t *)ctx->bLeafCounts.p, *dLast = dCounts + N;
	if ((rc = cmgpu_box_leafnums_batch_device(ctx, n, (const float *)ctx->bStarts.p, (const float *)ctx->bEnds.p, listsize,
	                                          (int32_t *)ctx->bLeafLists.p, dCounts, dLast, s))) return rc;
	if (L) CUDA_TRY(ctx, cudaMemcpyAsync(lists, ctx->bLeafLists.p, L, cudaMemcpyDeviceToHost, s));
	CUDA_TRY(ctx, cudaMemcpyAsync(counts, dCounts, N * 4, cudaMemcpyDeviceToHost, s));
	CUDA_TRY(ctx, cudaMemcpyAsync(lastLeafs, dLast, N * 4, cudaMemcpyDeviceToHost, s));
	CUDA_TRY(ctx, cudaStreamSynchronize(s));
	return cmgpu_check_status(ctx, s);
}

// ---- entity linking (PVS part), areas, snapshot culling (cm_leaf_queries.cuh) --------------------------------
int cmgpu_entity_clusters_batch_device(cmgpu_ctx *ctx, int n, const float *absmins, const float *absmaxs,
                                       cmgpu_entity_clusters_t *out, void *stream) {
	if (!ctx) return CMGPU_ERR_INVALID;
	if (n < 0 || (n > 0 && (!absmins || !absmaxs || !out))) return fail(ctx, CMGPU_ERR_INVALID, "cmgpu_entity_clusters_batch_device: bad argument");
	if (!ctx->hasMap) return fail(ctx, CMGPU_ERR_NO_MAP, "no map uploaded");
	if (n == 0) return CMGPU_OK;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	static_assert(sizeof(EntClusters) == sizeof(cmgpu_entity_clusters_t), "record layout");
	EntClusterBatch q{n, absmins, absmaxs, reinterpret_cast<EntClusters *>(out), ctx->leafCluster, ctx->dStatus};
	if (ctx->treeDepth > kMaxDepth) entityClustersKernel<kDeepDepth><<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(ctx->map, q);
	else entityClustersKernel<kMaxDepth><<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(ctx->map, q);
	ctx->launches++;
	CUDA_TRY(ctx, cudaGetLastError());
	return CMGPU_OK;
}

int cmgpu_entity_clusters_batch(cmgpu_ctx *ctx, int n, const float *absmins, const float *absmaxs,
                                cmgpu_entity_clusters_t *out) {
	if (!ctx) return CMGPU_ERR_INVALID;
	if (n < 0 || (n > 0 && (!absmins || !absmaxs || !out))) return fail(ctx, CMGPU_ERR_INVALID, "cmgpu_entity_clusters_batch: bad argument");
	if (n == 0) return CMGPU_OK;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	const size_t N = (size_t)n;
	int rc;
	if ((rc = reserve(ctx, ctx->bStarts, N * 12))) return rc;
	if ((rc = reserve(ctx, ctx->bEnds, N * 12))) return rc;
	if ((rc = reserve(ctx, ctx->bEntRecs, N * sizeof(cmgpu_entity_clusters_t)))) return rc;
	cudaStream_t s = ctx->streams[0];
	CUDA_TRY(ctx, cudaMemcpyAsync(ctx->bStarts.p, absmins, N * 12, cudaMemcpyHostToDevice, s));
	CUDA_TRY(ctx, cudaMemcpyAsync(ctx->bEnds.p, absmaxs, N * 12, cudaMemcpyHostToDevice, s));
	if ((rc = cmgpu_entity_clusters_batch_device(ctx, n, (const float *)ctx->bStarts.p, (const float *)ctx->bEnds.p,
	                                             (cmgpu_entity_clusters_t *)ctx->bEntRecs.p, s))) return rc;
	CUDA_TRY(ctx, cudaMemcpyAsync(out, ctx->bEntRecs.p, N * sizeof(cmgpu_entity_clusters_t), cudaMemcpyDeviceToHost, s));
	CUDA_TRY(ctx, cudaStreamSynchronize(s));
	return cmgpu_check_status(ctx, s);
}

int cmgpu_adjust_area_portal_state(cmgpu_ctx *ctx, int area1, int area2, int open) {      // cm_test.c:373-396
	if (!ctx) return CMGPU_ERR_INVALID;
	if (!ctx->hasMap) return fail(ctx, CMGPU_ERR_NO_MAP, "no map uploaded");
	if (area1 < 0 || area2 < 0) return CMGPU_OK;
	const int n = ctx->numAreas;
	if (area1 >= n || area2 >= n) return fail(ctx, CMGPU_ERR_INVALID, "CM_ChangeAreaPortalState: bad area number");
	int &a = ctx->areaPortals[(size_t)area1 * n + area2], &b = ctx->areaPortals[(size_t)area2 * n + area1];
	if (open) {
		a++;
		if (&a != &b) b++; else a++;                     // area1 == area2: the reference increments the same cell twice
	} else {
		a--;
		if (&a != &b) b--; else a--;
		if (b < 0) return fail(ctx, CMGPU_ERR_INVALID, "CM_AdjustAreaPortalState: negative reference count");
	}
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	return floodAreas(ctx);
}

int cmgpu_areas_connected(cmgpu_ctx *ctx, int area1, int area2) {                         // cm_test.c:404-424
	if (!ctx) return CMGPU_ERR_INVALID;
	if (ctx->noAreas) return 1;
	if (area1 < 0 || area2 < 0) return 0;
	if (area1 >= ctx->numAreas || area2 >= ctx->numAreas) return fail(ctx, CMGPU_ERR_INVALID, "area >= cm.numAreas");
	return ctx->floodnum[area1] == ctx->floodnum[area2] ? 1 : 0;
}

int cmgpu_write_area_bits(cmgpu_ctx *ctx, uint8_t *buffer, int area) {                   // cm_test.c:441-470
	if (!ctx || !buffer) return CMGPU_ERR_INVALID;
	const int n = ctx->numAreas, bytes = (n + 7) >> 3;
	if (ctx->noAreas || area == -1) {
		memset(buffer, 255, (size_t)bytes);
	} else {
		if (area < 0 || area >= n) return fail(ctx, CMGPU_ERR_INVALID, "CM_WriteAreaBits: bad area number");
		const int fl = ctx->floodnum[area];
		for (int i = 0; i < n; i++) {
			if (ctx->floodnum[i] == fl) buffer[i >> 3] |= (uint8_t)(1 << (i & 7));
		}
	}
	return bytes;
}

int cmgpu_set_no_areas(cmgpu_ctx *ctx, int noAreas) {
	if (!ctx) return CMGPU_ERR_INVALID;
	ctx->noAreas = noAreas ? 1 : 0;
	return CMGPU_OK;
}

int cmgpu_visible_from_points_batch_device(cmgpu_ctx *ctx, int nViews, const float *origins, int nEnts,
                                           const cmgpu_entity_clusters_t *ents, uint32_t *visible, int32_t *viewLeafs, void *stream) {
	if (!ctx) return CMGPU_ERR_INVALID;
	if (nViews < 0 || nEnts < 0 || (nViews > 0 && !origins) || (nViews > 0 && nEnts > 0 && (!ents || !visible)))
		return fail(ctx, CMGPU_ERR_INVALID, "cmgpu_visible_from_points_batch_device: bad argument");
	if (nViews > 65535) return fail(ctx, CMGPU_ERR_INVALID, "cmgpu_visible_from_points_batch_device: more than 65535 viewpoints in one call");
	if (!ctx->hasMap) return fail(ctx, CMGPU_ERR_NO_MAP, "no map uploaded");
	if (nViews == 0) return CMGPU_OK;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	cudaStream_t s = (cudaStream_t)stream;
	int rc;
	if (!viewLeafs) {
		if ((rc = reserve(ctx, ctx->bLeafCounts, (size_t)nViews * 4))) return rc;
		viewLeafs = (int32_t *)ctx->bLeafCounts.p;
	}
	if ((rc = cmgpu_point_leafnum_batch_device(ctx, nViews, origins, viewLeafs, stream))) return rc;
	if (nEnts == 0) return CMGPU_OK;
	VisBatch q{};
	q.nViews = nViews; q.nEnts = nEnts;
	q.viewLeafs = viewLeafs;
	q.ents = reinterpret_cast<const EntClusters *>(ents);
	q.visible = visible;
	q.leafCluster = ctx->leafCluster;
	q.floodnum = ctx->noAreas ? nullptr : (const int *)ctx->bFlood.p;
	q.numAreas = ctx->noAreas ? 0 : ctx->numAreas;
	if (!ctx->noAreas && ctx->numAreas == 0) { q.floodnum = (const int *)ctx->dStatus; q.numAreas = 0; }   // no areas: nothing is connected
	q.vis = ctx->vised ? (const unsigned char *)ctx->bVis.p : nullptr;
	q.numClusters = ctx->numClusters; q.clusterBytes = ctx->clusterBytes;
	dim3 grid((unsigned)((nEnts + 127) / 128), (unsigned)nViews);
	visibleFromPointsKernel<<<grid, 128, 0, s>>>(q);
	ctx->launches++;
	CUDA_TRY(ctx, cudaGetLastError());
	return CMGPU_OK;
}

int cmgpu_visible_from_points_batch(cmgpu_ctx *ctx, int nViews, const float *origins, int nEnts,
                                    const cmgpu_entity_clusters_t *ents, uint32_t *visible, int32_t *viewLeafs) {
	if (!ctx) return CMGPU_ERR_INVALID;
	if (nViews < 0 || nEnts < 0 || (nViews > 0 && !origins) || (nViews > 0 && nEnts > 0 && (!ents || !visible)))
		return fail(ctx, CMGPU_ERR_INVALID, "cmgpu_visible_from_points_batch: bad argument");
	if (nViews == 0) return CMGPU_OK;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	const size_t V = (size_t)nViews, E = (size_t)nEnts, words = (E + 31) / 32;
	int rc;
	if ((rc = reserve(ctx, ctx->bStarts, V * 12))) return rc;
	if ((rc = reserve(ctx, ctx->bHit, V * 4))) return rc;
	if ((rc = reserve(ctx, ctx->bEntRecs, E * sizeof(cmgpu_entity_clusters_t) + 4))) return rc;
	if ((rc = reserve(ctx, ctx->bVisible, V * words * 4 + 4))) return rc;
	cudaStream_t s = ctx->streams[0];
	CUDA_TRY(ctx, cudaMemcpyAsync(ctx->bStarts.p, origins, V * 12, cudaMemcpyHostToDevice, s));
	if (E) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->bEntRecs.p, ents, E * sizeof(cmgpu_entity_clusters_t), cudaMemcpyHostToDevice, s));
	if ((rc = cmgpu_visible_from_points_batch_device(ctx, nViews, (const float *)ctx->bStarts.p, nEnts,
	                                                 (const cmgpu_entity_clusters_t *)ctx->bEntRecs.p, (uint32_t *)ctx->bVisible.p,
	                                                 (int32_t *)ctx->bHit.p, s))) return rc;
	if (E) CUDA_TRY(ctx, cudaMemcpyAsync(visible, ctx->bVisible.p, V * words * 4, cudaMemcpyDeviceToHost, s));
	if (viewLeafs) CUDA_TRY(ctx, cudaMemcpyAsync(viewLeafs, ctx->bHit.p, V * 4, cudaMemcpyDeviceToHost, s));
	CUDA_TRY(ctx, cudaStreamSynchronize(s));
	return CMGPU_OK;
}

int cmgpu_measure_l2_read(cmgpu_ctx *ctx, size_t bytes, int passes, double *gbs) {
	if (!ctx || !gbs || bytes < 4096 || bytes > ((size_t)64 << 20) || passes < 1) return ctx ? fail(ctx, CMGPU_ERR_INVALID, "cmgpu_measure_l2_read: bad argument") : CMGPU_ERR_INVALID;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	void *buf = nullptr;
	CUDA_TRY(ctx, cudaMalloc(&buf, bytes + 16));
	cudaEvent_t a = nullptr, b = nullptr;
	cudaDeviceProp prop;
	bool ok = cudaMemset(buf, 0, bytes + 16) == cudaSuccess && cudaEventCreate(&a) == cudaSuccess && cudaEventCreate(&b) == cudaSuccess &&
	          cudaGetDeviceProperties(&prop, ctx->device) == cudaSuccess;
	float best = 0.0f;
	if (ok) {
		const size_t n16 = bytes / 16;
		uint4 *sink = (uint4 *)((char *)buf + n16 * 16);
		const int grid = prop.multiProcessorCount * 8;
		cudaStream_t s = ctx->streams[0];
		for (int rep = 0; rep < 6 && ok; rep++) {            // the first repetitions also bring the buffer into L2
			ok = cudaEventRecord(a, s) == cudaSuccess;
			l2ReadKernel<<<grid, 256, 0, s>>>((const uint4 *)buf, n16, passes, sink);
			ok = ok && cudaEventRecord(b, s) == cudaSuccess && cudaEventSynchronize(b) == cudaSuccess;
			float ms = 0.0f;
			ok = ok && cudaEventElapsedTime(&ms, a, b) == cudaSuccess;
			if (ok && rep >= 2 && ms > 0.0f) {
				const float rate = (float)((double)n16 * 16.0 * passes / (ms * 1e-3) / 1e9);
				if (rate > best) best = rate;
			}
		}
		ctx->launches += 6;
	}
	if (a) cudaEventDestroy(a);
	if (b) cudaEventDestroy(b);
	cudaFree(buf);
	if (!ok) return fail(ctx, CMGPU_ERR_CUDA, "cmgpu_measure_l2_read: %s", cudaGetErrorString(cudaGetLastError()));
	*gbs = best;
	return CMGPU_OK;
}

}  // extern "C"


// =====================================================================================
// botlib's AAS tracer (cm_aas.cuh; code/botlib/be_aas_sample.c)
// =====================================================================================
extern "C" {

int cmgpu_aas_upload(cmgpu_ctx *ctx, int numplanes, const float *planes, int numnodes, const int32_t *nodes, int numareas,
                     const int32_t *presencetypes, const int32_t *areacontents) {
	if (!ctx) return CMGPU_ERR_INVALID;
	if (numplanes < 2 || !planes || numnodes < 2 || !nodes || numareas < 1 || !presencetypes)
		return fail(ctx, CMGPU_ERR_BAD_MAP, "cmgpu_aas_upload: an AAS world has planes, a dummy node 0 and a root node 1, and a dummy area 0");
	if (numplanes & 1) return fail(ctx, CMGPU_ERR_BAD_MAP, "cmgpu_aas_upload: %d planes -- AAS planes come in opposite pairs (plane n ^ 1 is plane n flipped, be_aas_sample.c:469)", numplanes);
	// every index the kernels follow is checked here; the node graph must be a tree (a cycle would never end on the device)
	for (int i = 1; i < numnodes; i++) {
		const int32_t *nd = nodes + 3 * (size_t)i;
		if (nd[0] < 0 || nd[0] >= numplanes) return fail(ctx, CMGPU_ERR_BAD_MAP, "aas node %d: bad plane %d", i, nd[0]);
		for (int c = 1; c <= 2; c++) {
			if (nd[c] > 0 && nd[c] >= numnodes) return fail(ctx, CMGPU_ERR_BAD_MAP, "aas node %d: bad child %d", i, nd[c]);
			if (nd[c] < 0 && -(long long)nd[c] >= numareas) return fail(ctx, CMGPU_ERR_BAD_MAP, "aas node %d: bad area %d", i, -nd[c]);
		}
	}
	{
		std::vector<unsigned char> seen((size_t)numnodes, 0);
		std::vector<int> st;
		st.push_back(1);
		while (!st.empty()) {
			const int n = st.back();
			st.pop_back();
			if (seen[n]) return fail(ctx, CMGPU_ERR_BAD_MAP, "aas node %d reachable twice: not a tree", n);
			seen[n] = 1;
			for (int c = 1; c <= 2; c++) if (nodes[3 * (size_t)n + c] > 0) st.push_back(nodes[3 * (size_t)n + c]);
		}
	}
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	CUDA_TRY(ctx, cudaDeviceSynchronize());
	const size_t nNodes = (size_t)numnodes, nPlanes = (size_t)numplanes, nAreas = (size_t)numareas;
	std::vector<float4> nodePlane(nNodes), pl(nPlanes);
	std::vector<int4> nodeInfo(nNodes);
	for (size_t i = 0; i < nPlanes; i++) pl[i] = make_float4(planes[5 * i], planes[5 * i + 1], planes[5 * i + 2], planes[5 * i + 3]);
	nodePlane[0] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
	nodeInfo[0] = make_int4(0, 0, 0, 0);
	for (size_t i = 1; i < nNodes; i++) {
		nodePlane[i] = pl[(size_t)nodes[3 * i]];
		nodeInfo[i] = make_int4(nodes[3 * i], nodes[3 * i + 1], nodes[3 * i + 2], 0);
	}
	const size_t offInfo = nNodes * 16, offPlanes = offInfo + nNodes * 16, offPres = offPlanes + nPlanes * 16, offCont = offPres + nAreas * 4,
	             total = offCont + nAreas * 4;
	ctx->hasAas = false;
	int rc = reserve(ctx, ctx->bAas, total);
	if (rc) return rc;
	char *base = (char *)ctx->bAas.p;
	CUDA_TRY(ctx, cudaMemcpy(base, nodePlane.data(), nNodes * 16, cudaMemcpyHostToDevice));
	CUDA_TRY(ctx, cudaMemcpy(base + offInfo, nodeInfo.data(), nNodes * 16, cudaMemcpyHostToDevice));
	CUDA_TRY(ctx, cudaMemcpy(base + offPlanes, pl.data(), nPlanes * 16, cudaMemcpyHostToDevice));
	CUDA_TRY(ctx, cudaMemcpy(base + offPres, presencetypes, nAreas * 4, cudaMemcpyHostToDevice));
	if (areacontents) CUDA_TRY(ctx, cudaMemcpy(base + offCont, areacontents, nAreas * 4, cudaMemcpyHostToDevice));
	else CUDA_TRY(ctx, cudaMemset(base + offCont, 0, nAreas * 4));
	ctx->aas.nodePlane = (const float4 *)base;
	ctx->aas.nodeInfo = (const int4 *)(base + offInfo);
	ctx->aas.planes = (const float4 *)(base + offPlanes);
	ctx->aas.presence = (const int *)(base + offPres);
	ctx->aas.areaContents = (const int *)(base + offCont);
	ctx->aas.numNodes = numnodes; ctx->aas.numPlanes = numplanes; ctx->aas.numAreas = numareas;
	ctx->hasAas = true;
	return CMGPU_OK;
}

int cmgpu_aas_trace_client_bbox_batch_device(cmgpu_ctx *ctx, int n, const float *starts, const float *ends,
                                             const int32_t *presencetypes, int presence_stride, cmgpu_aas_trace_t *results, void *stream) {
	if (!ctx) return CMGPU_ERR_INVALID;
	if (n < 0 || (n > 0 && (!starts || !ends || !presencetypes || !results)))
		return fail(ctx, CMGPU_ERR_INVALID, "cmgpu_aas_trace_client_bbox_batch_device: bad argument");
	if (!ctx->hasAas) return fail(ctx, CMGPU_ERR_NO_MAP, "no AAS world uploaded (the reference returns a zeroed trace, be_aas_sample.c:414)");
	if (n == 0) return CMGPU_OK;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	AasTraceBatch q{};
	q.n = n; q.starts = starts; q.ends = ends;
	if (presence_stride) q.presencetypes = presencetypes; else q.sharedPresence = presencetypes[0];     // host [1], by value
	q.out = reinterpret_cast<int *>(results);
	aasTraceClientBBoxKernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(ctx->aas, q);
	ctx->launches++;
	CUDA_TRY(ctx, cudaGetLastError());
	return CMGPU_OK;
}

int cmgpu_aas_point_area_num_batch_device(cmgpu_ctx *ctx, int n, const float *points, int32_t *areas, void *stream) {
	if (!ctx) return CMGPU_ERR_INVALID;
	if (n < 0 || (n > 0 && (!points || !areas))) return fail(ctx, CMGPU_ERR_INVALID, "cmgpu_aas_point_area_num_batch_device: bad argument");
	if (!ctx->hasAas) return fail(ctx, CMGPU_ERR_NO_MAP, "no AAS world uploaded");
	if (n == 0) return CMGPU_OK;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	aasPointAreaNumKernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(ctx->aas, n, points, areas);
	ctx->launches++;
	CUDA_TRY(ctx, cudaGetLastError());
	return CMGPU_OK;
}

int cmgpu_aas_trace_client_bbox_batch(cmgpu_ctx *ctx, int n, const float *starts, const float *ends,
                                      const int32_t *presencetypes, int presence_stride, cmgpu_aas_trace_t *results) {
	if (!ctx) return CMGPU_ERR_INVALID;
	if (n < 0 || (n > 0 && (!starts || !ends || !presencetypes || !results)))
		return fail(ctx, CMGPU_ERR_INVALID, "cmgpu_aas_trace_client_bbox_batch: bad argument");
	if (!ctx->hasAas) return fail(ctx, CMGPU_ERR_NO_MAP, "no AAS world uploaded");
	if (n == 0) return CMGPU_OK;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	const size_t N = (size_t)n;
	int rc;
	if ((rc = reserve(ctx, ctx->bStarts, N * 12))) return rc;
	if ((rc = reserve(ctx, ctx->bEnds, N * 12))) return rc;
	if (presence_stride && (rc = reserve(ctx, ctx->bMasks, N * 4))) return rc;
	if ((rc = reserve(ctx, ctx->bOut, N * sizeof(cmgpu_aas_trace_t)))) return rc;
	cudaStream_t s = ctx->streams[0];
	CUDA_TRY(ctx, cudaMemcpyAsync(ctx->bStarts.p, starts, N * 12, cudaMemcpyHostToDevice, s));
	CUDA_TRY(ctx, cudaMemcpyAsync(ctx->bEnds.p, ends, N * 12, cudaMemcpyHostToDevice, s));
	if (presence_stride) CUDA_TRY(ctx, cudaMemcpyAsync(ctx->bMasks.p, presencetypes, N * 4, cudaMemcpyHostToDevice, s));
	rc = cmgpu_aas_trace_client_bbox_batch_device(ctx, n, (const float *)ctx->bStarts.p, (const float *)ctx->bEnds.p,
	                                              presence_stride ? (const int32_t *)ctx->bMasks.p : presencetypes, presence_stride,
	                                              (cmgpu_aas_trace_t *)ctx->bOut.p, s);
	if (rc) return rc;
	CUDA_TRY(ctx, cudaMemcpyAsync(results, ctx->bOut.p, N * sizeof(cmgpu_aas_trace_t), cudaMemcpyDeviceToHost, s));
	CUDA_TRY(ctx, cudaStreamSynchronize(s));
	return CMGPU_OK;
}

int cmgpu_aas_point_area_num_batch(cmgpu_ctx *ctx, int n, const float *points, int32_t *areas) {
	if (!ctx) return CMGPU_ERR_INVALID;
	if (n < 0 || (n > 0 && (!points || !areas))) return fail(ctx, CMGPU_ERR_INVALID, "cmgpu_aas_point_area_num_batch: bad argument");
	if (!ctx->hasAas) return fail(ctx, CMGPU_ERR_NO_MAP, "no AAS world uploaded");
	if (n == 0) return CMGPU_OK;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	const size_t N = (size_t)n;
	int rc;
	if ((rc = reserve(ctx, ctx->bStarts, N * 12))) return rc;
	if ((rc = reserve(ctx, ctx->bHit, N * 4))) return rc;
	cudaStream_t s = ctx->streams[0];
	CUDA_TRY(ctx, cudaMemcpyAsync(ctx->bStarts.p, points, N * 12, cudaMemcpyHostToDevice, s));
	rc = cmgpu_aas_point_area_num_batch_device(ctx, n, (const float *)ctx->bStarts.p, (int32_t *)ctx->bHit.p, s);
	if (rc) return rc;
	CUDA_TRY(ctx, cudaMemcpyAsync(areas, ctx->bHit.p, N * 4, cudaMemcpyDeviceToHost, s));
	CUDA_TRY(ctx, cudaStreamSynchronize(s));
	return CMGPU_OK;
}

void cmgpu_aas_default_settings(cmgpu_aas_settings *st) {          // AAS_InitSettings, be_aas_move.c:74-92 (the libvars' defaults)
	if (!st) return;
	const cmgpu_aas_settings d = {6.0f, 100.0f, 800.0f, 1.0f, 400.0f, 320.0f, 100.0f, 150.0f, 10.0f, 1.0f, 4.0f, 19.0f, 0.7f, 33.0f, 270.0f};
	*st = d;
}

int cmgpu_aas_set_settings(cmgpu_ctx *ctx, const cmgpu_aas_settings *st) {
	if (!ctx || !st) return CMGPU_ERR_INVALID;
	static_assert(sizeof(cmgpu_aas_settings) == sizeof(AasPhysics), "the two are the same fifteen floats");
	memcpy(&ctx->aasPhys, st, sizeof(AasPhysics));
	return CMGPU_OK;
}

int cmgpu_aas_predict_client_movement_batch_device(cmgpu_ctx *ctx, int n, const cmgpu_aas_moves *in, cmgpu_aas_clientmove_t *moves,
                                                   int32_t *results, void *stream) {
	if (!ctx) return CMGPU_ERR_INVALID;
	if (n < 0 || !in || (n > 0 && (!in->origins || !in->velocities || !in->cmdmoves || !in->frametimes || !in->presencetypes || !in->ongrounds ||
	                               !in->cmdframes || !in->maxframes || !in->stopevents || !in->stopareanums || !moves)))
		return fail(ctx, CMGPU_ERR_INVALID, "cmgpu_aas_predict_client_movement_batch_device: bad argument");
	if (!ctx->hasAas) return fail(ctx, CMGPU_ERR_NO_MAP, "no AAS world uploaded");
	if (n == 0) return CMGPU_OK;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	AasMoveBatch q{};
	q.n = n;
	q.origins = in->origins; q.velocities = in->velocities; q.cmdmoves = in->cmdmoves; q.frametimes = in->frametimes;
	q.presencetypes = in->presencetypes; q.ongrounds = in->ongrounds; q.cmdframes = in->cmdframes; q.maxframes = in->maxframes;
	q.stopevents = in->stopevents; q.stopareanums = in->stopareanums;
	q.out = reinterpret_cast<int *>(moves);
	q.results = results;
	ctx->map.noCurves = ctx->noCurves;
	ctx->map.playerCurveClip = ctx->playerCurveClip;
	// the world's contents (AAS_PointContents) come from the clip map of this context, when one is uploaded
	aasPredictClientMovementKernel<<<(n + 63) / 64, 64, 0, (cudaStream_t)stream>>>(ctx->aas, ctx->map, ctx->hasMap ? 1 : 0, ctx->aasPhys, q);
	ctx->launches++;
	CUDA_TRY(ctx, cudaGetLastError());
	return CMGPU_OK;
}

int cmgpu_aas_predict_client_movement_batch(cmgpu_ctx *ctx, int n, const cmgpu_aas_moves *in, cmgpu_aas_clientmove_t *moves, int32_t *results) {
	if (!ctx) return CMGPU_ERR_INVALID;
	if (n < 0 || !in || (n > 0 && (!in->origins || !in->velocities || !in->cmdmoves || !in->frametimes || !in->presencetypes || !in->ongrounds ||
	                               !in->cmdframes || !in->maxframes || !in->stopevents || !in->stopareanums || !moves)))
		return fail(ctx, CMGPU_ERR_INVALID, "cmgpu_aas_predict_client_movement_batch: bad argument");
	if (!ctx->hasAas) return fail(ctx, CMGPU_ERR_NO_MAP, "no AAS world uploaded");
	if (n == 0) return CMGPU_OK;
	CUDA_TRY(ctx, cudaSetDevice(ctx->device));
	const size_t N = (size_t)n;
	// one staging block: three [n][3] float arrays, then seven [n] arrays of 4-byte values
	const size_t bytesIn = N * (3 * 12 + 7 * 4);
	int rc;
	if ((rc = reserve(ctx, ctx->bAasIn, bytesIn))) return rc;
	if ((rc = reserve(ctx, ctx->bAasOut, N * (sizeof(cmgpu_aas_clientmove_t) + 4)))) return rc;
	cudaStream_t s = ctx->streams[0];
	char *d = (char *)ctx->bAasIn.p;
	cmgpu_aas_moves dv{};
	size_t off = 0;
	auto put = [&](const void *src, size_t bytes) -> const void * {
		const void *at = d + off;
		if (cudaMemcpyAsync(d + off, src, bytes, cudaMemcpyHostToDevice, s) != cudaSuccess) rc = CMGPU_ERR_CUDA;
		off += bytes;
		return at;
	};
	dv.origins = (const float *)put(in->origins, N * 12);
	dv.velocities = (const float *)put(in->velocities, N * 12);
	dv.cmdmoves = (const float *)put(in->cmdmoves, N * 12);
	dv.frametimes = (const float *)put(in->frametimes, N * 4);
	dv.presencetypes = (const int32_t *)put(in->presencetypes, N * 4);
	dv.ongrounds = (const int32_t *)put(in->ongrounds, N * 4);
	dv.cmdframes = (const int32_t *)put(in->cmdframes, N * 4);
	dv.maxframes = (const int32_t *)put(in->maxframes, N * 4);
	dv.stopevents = (const int32_t *)put(in->stopevents, N * 4);
	dv.stopareanums = (const int32_t *)put(in->stopareanums, N * 4);
	if (rc) return fail(ctx, CMGPU_ERR_CUDA, "cudaMemcpyAsync (AAS moves): %s", cudaGetErrorString(cudaGetLastError()));
	cmgpu_aas_clientmove_t *dMoves = (cmgpu_aas_clientmove_t *)ctx->bAasOut.p;
	int32_t *dRes = (int32_t *)((char *)ctx->bAasOut.p + N * sizeof(cmgpu_aas_clientmove_t));
	rc = cmgpu_aas_predict_client_movement_batch_device(ctx, n, &dv, dMoves, dRes, s);
	if (rc) return rc;
	CUDA_TRY(ctx, cudaMemcpyAsync(moves, dMoves, N * sizeof(cmgpu_aas_clientmove_t), cudaMemcpyDeviceToHost, s));
	if (results) CUDA_TRY(ctx, cudaMemcpyAsync(results, dRes, N * 4, cudaMemcpyDeviceToHost, s));
	CUDA_TRY(ctx, cudaStreamSynchronize(s));
	return CMGPU_OK;
}

}  // extern "C"


// =====================================================================================
// Bezier patches -> collision facets on the device (CM_GeneratePatchCollide, cm_patch.c:1137-1203)
// =====================================================================================
// One thread per patch runs the builder of cm_patchgen_core.h -- the code the host loader runs -- in its own scratch
// block in global memory.  The work of a patch is one long dependent chain (every new plane is compared with all the
// planes found so far, facets are validated against them in order), so there is nothing to spread over a warp without
// changing the order in which planes are numbered; patches are independent, and a map has hundreds.
namespace cmgpu {

__global__ void __launch_bounds__(32) patchCollideKernel(int n, const int *widths, const int *heights, const int *firstPoint,
                                                         const float *points, pg::PatchScratch *scratch) {
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	pg::Builder(scratch[i]).run(widths[i], heights[i], points + 3 * (size_t)firstPoint[i]);
}

bool generatePatchCollidesOnDevice(int device, int n, const int *widths, const int *heights, const int *firstPoint,
                                   const float *points, size_t numPoints, PatchCollide *out, char *err, size_t errlen) {
	auto failWith = [&](const char *what, cudaError_t e) {
		if (err && errlen) snprintf(err, errlen, "%s: %s", what, cudaGetErrorString(e));
		return false;
	};
	if (n <= 0) return true;
	cudaError_t e = cudaSetDevice(device);
	if (e != cudaSuccess) return failWith("cudaSetDevice", e);
	int *dMeta = nullptr;
	float *dPoints = nullptr;
	pg::PatchScratch *dScratch = nullptr;
	// patches are built in waves so that the scratch stays within 1 GiB whatever the map
	const int wave = (int)std::min<size_t>((size_t)n, ((size_t)1 << 30) / sizeof(pg::PatchScratch));
	bool ok = true;
	std::vector<int> meta(3 * (size_t)n);
	for (int i = 0; i < n; i++) { meta[(size_t)i] = widths[i]; meta[(size_t)n + i] = heights[i]; meta[2 * (size_t)n + i] = firstPoint[i]; }
	std::unique_ptr<pg::PatchScratch> host(new (std::nothrow) pg::PatchScratch);
	if (!host) { if (err && errlen) snprintf(err, errlen, "out of memory"); return false; }
	if ((e = cudaMalloc(&dMeta, meta.size() * sizeof(int))) != cudaSuccess) ok = failWith("cudaMalloc", e);
	if (ok && (e = cudaMalloc(&dPoints, (numPoints ? numPoints : 1) * 12)) != cudaSuccess) ok = failWith("cudaMalloc", e);
	if (ok && (e = cudaMalloc(&dScratch, (size_t)wave * sizeof(pg::PatchScratch))) != cudaSuccess) ok = failWith("cudaMalloc (patch scratch)", e);
	if (ok && (e = cudaMemcpy(dMeta, meta.data(), meta.size() * sizeof(int), cudaMemcpyHostToDevice)) != cudaSuccess) ok = failWith("cudaMemcpy", e);
	if (ok && (e = cudaMemcpy(dPoints, points, numPoints * 12, cudaMemcpyHostToDevice)) != cudaSuccess) ok = failWith("cudaMemcpy", e);
	for (int base = 0; ok && base < n; base += wave) {
		const int cnt = std::min(wave, n - base);
		patchCollideKernel<<<(cnt + 31) / 32, 32>>>(cnt, dMeta + base, dMeta + n + base, dMeta + 2 * (size_t)n + base, dPoints, dScratch);
		if ((e = cudaGetLastError()) != cudaSuccess || (e = cudaDeviceSynchronize()) != cudaSuccess) { ok = failWith("patchCollideKernel", e); break; }
		for (int i = 0; i < cnt && ok; i++) {
			// results only: the tail of the block (planes, facets, bounds, counts, error), not the 330 KB of grid scratch
			const size_t off = offsetof(pg::PatchScratch, planes);
			if ((e = cudaMemcpy((char *)host.get() + off, (const char *)(dScratch + i) + off, sizeof(pg::PatchScratch) - off, cudaMemcpyDeviceToHost)) != cudaSuccess) {
				ok = failWith("cudaMemcpy (patch results)", e);
				break;
			}
			if (host->error) {
				if (err && errlen) snprintf(err, errlen, "%s", pg::errorText(host->error));
				ok = false;
				break;
			}
			patchCollideFromScratch(*host, &out[base + i]);
		}
	}
	cudaFree(dMeta); cudaFree(dPoints); cudaFree(dScratch);
	return ok;
}

}  // namespace cmgpu
