// Host half of the result path of a host-pointer batch (cm_gpu.cu: traceHostHybrid).
//
// A trace_t is 56 bytes, of which the walk decides 16: fraction, the two solid flags, which brush side clipped the sweep
// and which brush was hit.  The other 40 -- endpos, the plane, surface flags, contents (cm_trace.c:671-686 and the plane
// copies at cm_trace.c:352-356) -- follow from those 16, the caller's own start / end arrays and three tables of the
// map.  Over PCIe the 56-byte records are what a host batch waits for, so a share of every batch
// comes back as the 16-byte results and is written out HERE, by worker threads of the calling process, with the same
// float operations in the same order as expandRecordsKernel.  Nothing in this file walks a tree or tests a brush: it
// is the record formatter of the device's results, not a second implementation of the trace.
#pragma once
#include <cstddef>
#include <cstdint>

namespace hostrec {

// what a record needs from the clip side that produced it: cplane_t's normal and dist, type | signbits << 8, surfaceFlags
struct SideRec {
	float plane[4];
	uint32_t typeSignbits;   // trace_t word 10
	uint32_t surfaceFlags;   // trace_t word 11
	uint32_t pad[2];
};
static_assert(sizeof(SideRec) == 32, "two per cache line");

struct Tables {
	const SideRec *sides = nullptr;
	const int32_t *brushContents = nullptr;
};

// records [lo, hi) of one batch: compact[i] = {fraction bits, allsolid | startsolid << 1, side or -1, brush or -1}
void expand(const Tables &t, const uint32_t *compact, const float *starts, const float *ends, void *records, int32_t *hitIds, size_t lo, size_t hi);

class Pool;                  // worker threads; a job is a contiguous range of one batch, cut into slices
Pool *poolCreate(int threads);
void poolDestroy(Pool *p);
int poolThreads(const Pool *p);

struct Job {
	Tables tables;
	const uint32_t *compact;   // indexed like the records: compact + 4 * i
	const float *starts, *ends;
	void *records;
	int32_t *hitIds;
	size_t lo, hi;
	// held: the job's 16-byte results are still on their way when it is queued; poolRelease(ticket) makes it runnable
	// (called from a CUDA host function behind the copy: the workers themselves never touch the CUDA API)
	bool held;
	// A staging job instead: items [lo, hi) of the caller's pageable start / end arrays are copied into pinned memory
	// (stageStarts / stageEnds non-null; tables, compact, records unused), for the upload that follows.
	float *stageStarts, *stageEnds;
};

// Queue a job (returns its ticket); released jobs are taken up in the order they were queued.
int poolSubmit(Pool *p, const Job &job);
void poolRelease(Pool *p, int ticket);
// block until this job has been done
void poolWaitJob(Pool *p, int ticket);
// records queued or being written, over all jobs
size_t poolPendingRecords(const Pool *p);
// measured rate of the whole pool, records per second (0 until a job has run)
double poolRate(const Pool *p);
// block until every submitted job has been written
void poolWaitAll(Pool *p);
// time zero of poolLastDoneMs (and of the CMGPU_HYBRID_TRACE log lines): the start of a call
void poolSetEpoch(Pool *p);
// when the pool last ran out of work, ms after the epoch
double poolLastDoneMs(Pool *p);
// time since the epoch during which at least one worker was writing records (call it when the pool is idle)
double poolBusyMs(Pool *p);

}  // namespace hostrec
