// Thread- and address-sanitizer driver of the worker pool (cm_host_records.cpp): staged sweeps, held jobs released from another
// thread in reverse order, 1 / 4 / 9 workers; records compared with the single-threaded formatter.
// g++ -std=c++17 -O1 -g -fsanitize=thread -Ithreecore_b200/csrc scripts/host_records_sanitize.cpp threecore_b200/csrc/cm_host_records.cpp -pthread
// (and -fsanitize=address,undefined): both clean at the end of round 2.
#include "cm_host_records.h"
#include <vector>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <thread>
using namespace hostrec;
int main() {
	const size_t N = 600001; const int nSides = 5000, nBrushes = 700;
	std::vector<SideRec> sides(nSides + 1); std::vector<int32_t> cont(nBrushes + 1, 3);
	for (int i = 1; i <= nSides; i++) { sides[i].plane[0] = i; sides[i].typeSignbits = i & 0xff; sides[i].surfaceFlags = i * 3; }
	std::vector<uint32_t> c(4 * N); std::vector<float> st(3 * N + 1), en(3 * N + 1), sst(3 * N + 1), sen(3 * N + 1);
	srand(2);
	for (size_t i = 0; i < N; i++) { bool hit = rand() & 1; float f = hit ? (rand() % 1000) / 1000.0f : 1.0f; memcpy(&c[4 * i], &f, 4); c[4 * i + 1] = rand() & 3; c[4 * i + 2] = hit ? rand() % nSides : -1; c[4 * i + 3] = hit ? rand() % nBrushes : -1; }
	for (size_t i = 0; i < 3 * N; i++) { st[i] = rand() % 4096; en[i] = rand() % 4096; }
	std::vector<char> ref(56 * N), out(56 * N);
	Tables t{sides.data(), cont.data()};
	expand(t, c.data(), st.data(), en.data(), ref.data(), nullptr, 0, N);
	for (int threads : {1, 4, 9}) {
		for (int rep = 0; rep < 3; rep++) {
			memset(out.data(), 0, out.size());
			Pool *p = poolCreate(threads);
			poolSetEpoch(p);
			std::vector<int> tickets;
			for (size_t lo = 0; lo < N; lo += 70000) {
				size_t hi = lo + 70000 < N ? lo + 70000 : N;
				Job s{}; s.starts = st.data(); s.ends = en.data(); s.stageStarts = sst.data(); s.stageEnds = sen.data(); s.lo = lo; s.hi = hi;
				int ts = poolSubmit(p, s);
				if (rep) poolWaitJob(p, ts);
				Job j{}; j.tables = t; j.compact = c.data(); j.starts = rep ? sst.data() : st.data(); j.ends = rep ? sen.data() : en.data(); j.records = out.data(); j.lo = lo; j.hi = hi; j.held = true;
				tickets.push_back(poolSubmit(p, j));
			}
			// release from another thread, interleaved with queries
			std::thread rel([&] { for (size_t k = tickets.size(); k-- > 0;) { poolRelease(p, tickets[k]); (void)poolPendingRecords(p); (void)poolRate(p); } });
			rel.join();
			poolWaitAll(p);
			(void)poolBusyMs(p); (void)poolLastDoneMs(p);
			poolDestroy(p);
			if (memcmp(out.data(), ref.data(), out.size())) { printf("MISMATCH threads=%d rep=%d\n", threads, rep); return 1; }
		}
	}
	printf("ok\n");
	return 0;
}
