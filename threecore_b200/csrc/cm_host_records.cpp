// cm_host_records.cpp -- see cm_host_records.h: the record formatter of a host batch's 16-byte results, and its threads.
#include "cm_host_records.h"

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <thread>
#include <vector>

#include <emmintrin.h>

namespace hostrec {

// One record, cm_trace.c:671-686 (endpos) + the plane copy of the clip side (cm_trace.c:352-356), written exactly like
// expandRecordsKernel (cm_trace_fast.cuh) writes it: endpos = end when fraction == 1, else start + fraction * (end -
// start) with a separately rounded subtraction, product and sum per component (SSE2 mulps / addps / subps round each
// lane like the scalar operations; this file is built with -ffp-contract=off and without -mfma).  Branch-free: whether
// a sweep hit anything is a coin toss per record, and a mispredicted branch costs more than the record.
// `sides` and `brushContents` have a zero entry in front: index = side + 1, brush + 1 (-1 = none).
template <bool NT>
static inline void oneRecord(const SideRec *sides, const int32_t *brushContents, const uint32_t *compact, const __m128 s, const __m128 e, char *o,
                             int32_t *hitId) {
	const __m128i c = _mm_loadu_si128(reinterpret_cast<const __m128i *>(compact));   // fraction, flags, side, brush
	const uint32_t flags = (uint32_t)_mm_cvtsi128_si32(_mm_shuffle_epi32(c, 1));
	const int32_t side = _mm_cvtsi128_si32(_mm_shuffle_epi32(c, 2));
	const int32_t brush = _mm_cvtsi128_si32(_mm_shuffle_epi32(c, 3));
	const SideRec *sr = sides + (side + 1);
	const uint32_t contents = (uint32_t)brushContents[brush + 1];
	const __m128 f = _mm_castsi128_ps(_mm_shuffle_epi32(c, 0));
	const __m128 mid = _mm_add_ps(s, _mm_mul_ps(f, _mm_sub_ps(e, s)));
	const __m128 whole = _mm_cmpeq_ps(f, _mm_set1_ps(1.0f));
	const __m128i p = _mm_and_si128(_mm_castps_si128(_mm_or_ps(_mm_and_ps(whole, e), _mm_andnot_ps(whole, mid))), _mm_setr_epi32(-1, -1, -1, 0));
	// words 0-3: allsolid, startsolid, fraction, endpos.x | 4-7: endpos.y, endpos.z, normal.x, normal.y |
	// 8-11: normal.z, dist, type | signbits << 8, surfaceFlags | 12-13: contents, entityNum = 0
	const __m128i fl = _mm_setr_epi32((int)(flags & 1u), (int)((flags >> 1) & 1u), 0, 0);
	const __m128i fr = _mm_and_si128(_mm_shuffle_epi32(c, _MM_SHUFFLE(3, 0, 3, 3)), _mm_setr_epi32(0, 0, -1, 0));
	const __m128i r0 = _mm_or_si128(_mm_or_si128(fl, fr), _mm_slli_si128(p, 12));
	const __m128i pl = _mm_loadu_si128(reinterpret_cast<const __m128i *>(sr->plane));
	const __m128i r1 = _mm_or_si128(_mm_srli_si128(p, 4), _mm_slli_si128(pl, 8));
	const __m128i tf = _mm_loadl_epi64(reinterpret_cast<const __m128i *>(&sr->typeSignbits));
	const __m128i r2 = _mm_or_si128(_mm_srli_si128(pl, 8), _mm_slli_si128(tf, 8));
	if (NT) {
		long long *q = reinterpret_cast<long long *>(o);
		_mm_stream_si64(q + 0, _mm_cvtsi128_si64(r0)); _mm_stream_si64(q + 1, _mm_cvtsi128_si64(_mm_srli_si128(r0, 8)));
		_mm_stream_si64(q + 2, _mm_cvtsi128_si64(r1)); _mm_stream_si64(q + 3, _mm_cvtsi128_si64(_mm_srli_si128(r1, 8)));
		_mm_stream_si64(q + 4, _mm_cvtsi128_si64(r2)); _mm_stream_si64(q + 5, _mm_cvtsi128_si64(_mm_srli_si128(r2, 8)));
		_mm_stream_si64(q + 6, (long long)contents);
	} else {
		_mm_storeu_si128(reinterpret_cast<__m128i *>(o), r0);
		_mm_storeu_si128(reinterpret_cast<__m128i *>(o + 16), r1);
		_mm_storeu_si128(reinterpret_cast<__m128i *>(o + 32), r2);
		const long long last = (long long)contents;
		memcpy(o + 48, &last, 8);
	}
	if (hitId) *hitId = brush;
}

template <bool NT>
static void expandT(const Tables &t, const uint32_t *compact, const float *starts, const float *ends, char *out, int32_t *hitIds, size_t lo, size_t hi) {
	if (lo >= hi) return;
	// all but the last record read their start and end with one 16-byte load each (the fourth lane is the next record's
	// first coordinate and is masked out of endpos); the last one must not read past the caller's arrays
	for (size_t i = lo; i + 1 < hi; i++) {
		const __m128 s = _mm_loadu_ps(starts + 3 * i), e = _mm_loadu_ps(ends + 3 * i);
		oneRecord<NT>(t.sides, t.brushContents, compact + 4 * i, s, e, out + 56 * i, hitIds ? hitIds + i : nullptr);
	}
	{
		const size_t i = hi - 1;
		const __m128 s = _mm_setr_ps(starts[3 * i], starts[3 * i + 1], starts[3 * i + 2], 0.0f);
		const __m128 e = _mm_setr_ps(ends[3 * i], ends[3 * i + 1], ends[3 * i + 2], 0.0f);
		oneRecord<NT>(t.sides, t.brushContents, compact + 4 * i, s, e, out + 56 * i, hitIds ? hitIds + i : nullptr);
	}
	if (NT) _mm_sfence();
}

void expand(const Tables &t, const uint32_t *compact, const float *starts, const float *ends, void *records, int32_t *hitIds, size_t lo, size_t hi) {
	// streaming stores need 8-byte aligned records (every array of trace_t that came from an allocator is)
	if ((reinterpret_cast<uintptr_t>(records) & 7u) == 0) expandT<true>(t, compact, starts, ends, static_cast<char *>(records), hitIds, lo, hi);
	else expandT<false>(t, compact, starts, ends, static_cast<char *>(records), hitIds, lo, hi);
}

// ---- worker threads ----------------------------------------------------------------------------------------------
// Workers take slices of released jobs front to back and sleep on a condition variable otherwise.  The pool measures
// its own rate (records per second of wall time while at least one slice is in progress) for the dispatcher's split.
constexpr size_t kSlice = 16384;

class Pool {
public:
	explicit Pool(int threads) {
		if (threads < 1) threads = 1;
		for (int i = 0; i < threads; i++) workers_.emplace_back([this] { run(); });
	}
	~Pool() {
		{
			std::lock_guard<std::mutex> g(m_);
			quit_ = true;
		}
		cv_.notify_all();
		for (auto &w : workers_) w.join();
	}
	int threads() const { return (int)workers_.size(); }

	int submit(const Job &job) {
		int ticket;
		{
			std::lock_guard<std::mutex> g(m_);
			Entry e;
			e.job = job;
			e.next = job.lo;
			e.left = job.hi > job.lo ? (job.hi - job.lo + kSlice - 1) / kSlice : 0;
			ticket = e.ticket = nextTicket_++;
			e.released = !job.held;
			if (e.left == 0) return ticket;
			pending_.fetch_add(job.hi - job.lo, std::memory_order_relaxed);
			jobs_.push_back(e);
			unfinished_++;
		}
		if (!job.held) cv_.notify_all();
		return ticket;
	}
	void release(int ticket) {
		{
			std::lock_guard<std::mutex> g(m_);
			for (auto &e : jobs_) if (e.ticket == ticket) {
				e.released = true;
				if (trace_) fprintf(stderr, "[hostrec] job %zu..%zu arrived at %.2f ms\n", e.job.lo, e.job.hi, sinceEpochMs());
				break;
			}
		}
		cv_.notify_all();
	}
	void waitJob(int ticket) {
		std::unique_lock<std::mutex> g(m_);
		done_.wait(g, [&] {
			for (auto &j : jobs_) if (j.ticket == ticket) return j.left == 0;
			return true;                                    // no longer queued: done
		});
	}
	size_t pendingRecords() const { return pending_.load(std::memory_order_relaxed); }
	double rate() const { return rate_.load(std::memory_order_relaxed); }
	void waitAll() {
		std::unique_lock<std::mutex> g(m_);
		done_.wait(g, [this] { return unfinished_ == 0; });
	}

private:
	struct Entry {
		Job job;
		size_t next = 0, left = 0;      // first record not yet handed out; slices not yet finished
		int ticket = 0;
		bool released = false;
	};
	void finishLocked(Entry &) {
		unfinished_--;
		while (!jobs_.empty() && jobs_.front().left == 0 && jobs_.front().released) jobs_.pop_front();
		if (unfinished_ == 0) lastDoneMs_ = sinceEpochMs();
		done_.notify_all();
	}
	void run() {
		std::unique_lock<std::mutex> g(m_);
		for (;;) {
			Entry *e = nullptr;
			for (auto &j : jobs_) if (j.released && j.next < j.job.hi) { e = &j; break; }
			if (!e) {
				if (quit_) return;
				cv_.wait(g);
				continue;
			}
			const size_t lo = e->next, hi = lo + kSlice < e->job.hi ? lo + kSlice : e->job.hi;
			e->next = hi;
			const Job job = e->job;
			const int ticket = e->ticket;
			if (busy_++ == 0) busySince_ = std::chrono::steady_clock::now();
			g.unlock();
			if (job.stageStarts) {
				memcpy(job.stageStarts + 3 * lo, job.starts + 3 * lo, 12 * (hi - lo));
				memcpy(job.stageEnds + 3 * lo, job.ends + 3 * lo, 12 * (hi - lo));
			} else {
				expand(job.tables, job.compact, job.starts, job.ends, job.records, job.hitIds, lo, hi);
			}
			pending_.fetch_sub(hi - lo, std::memory_order_relaxed);
			g.lock();
			busyRecords_ += hi - lo;
			if (--busy_ == 0) {
				// a stretch of uninterrupted work ends: fold its rate into the estimate (short stretches say little)
				const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - busySince_).count();
				busyMs_ += 1e3 * secs;
				if (busyRecords_ >= 4 * kSlice && secs > 0) {
					const double r = (double)busyRecords_ / secs, old = rate_.load(std::memory_order_relaxed);
					rate_.store(old > 0 ? 0.5 * old + 0.5 * r : r, std::memory_order_relaxed);
				}
				busyRecords_ = 0;
			}
			for (auto &j : jobs_) if (j.ticket == ticket) {
				if (--j.left == 0) {
					if (trace_) fprintf(stderr, "[hostrec] job %zu..%zu written at %.2f ms\n", j.job.lo, j.job.hi, sinceEpochMs());
					finishLocked(j);
				}
				break;
			}
		}
	}

public:
	void setEpoch() { std::lock_guard<std::mutex> g(m_); epoch_ = std::chrono::steady_clock::now(); lastDoneMs_ = 0.0; busyMs_ = 0.0; }
	double busyMs() { std::lock_guard<std::mutex> g(m_); return busyMs_; }
	double lastDoneMs() { std::lock_guard<std::mutex> g(m_); return lastDoneMs_; }
private:
	double sinceEpochMs() const { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - epoch_).count(); }
	std::chrono::steady_clock::time_point epoch_ = std::chrono::steady_clock::now();
	double lastDoneMs_ = 0.0, busyMs_ = 0.0;    // since the epoch: when the pool last ran dry; time with at least one slice in progress
	const bool trace_ = getenv("CMGPU_HYBRID_TRACE") != nullptr;
	std::mutex m_;
	std::condition_variable cv_, done_;
	std::deque<Entry> jobs_;
	std::vector<std::thread> workers_;
	std::atomic<size_t> pending_{0};
	std::atomic<double> rate_{0.0};
	std::chrono::steady_clock::time_point busySince_;
	size_t busyRecords_ = 0;
	int busy_ = 0, nextTicket_ = 1, unfinished_ = 0;
	bool quit_ = false;
};

Pool *poolCreate(int threads) { return new Pool(threads); }
void poolDestroy(Pool *p) { delete p; }
int poolThreads(const Pool *p) { return p->threads(); }
int poolSubmit(Pool *p, const Job &job) { return p->submit(job); }
void poolRelease(Pool *p, int ticket) { p->release(ticket); }
void poolWaitJob(Pool *p, int ticket) { p->waitJob(ticket); }
size_t poolPendingRecords(const Pool *p) { return p->pendingRecords(); }
double poolRate(const Pool *p) { return p->rate(); }
void poolWaitAll(Pool *p) { p->waitAll(); }
void poolSetEpoch(Pool *p) { p->setEpoch(); }
double poolLastDoneMs(Pool *p) { return p->lastDoneMs(); }
double poolBusyMs(Pool *p) { return p->busyMs(); }

}  // namespace hostrec
