// microbench: host-side expansion of 16-byte results into 56-byte trace_t records
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cstdint>
#include <thread>
#include <vector>
#include <chrono>
#include <emmintrin.h>
struct C16 { float fraction; unsigned flags; int planeRef; int hitId; };
static void expandRange(size_t lo, size_t hi, const C16 *c, const float *starts, const float *ends, const float *planes, const int *meta, const int *contents, char *out, int nt) {
	for (size_t i = lo; i < hi; i++) {
		if (nt & 2) { const size_t j = i + 12 < hi ? i + 12 : i; const int pr = c[j].planeRef; if (pr >= 0) { __builtin_prefetch(planes + 4 * (size_t)pr); __builtin_prefetch(meta + 2 * (size_t)pr); } }
		const C16 r = c[i];
		const float *s = starts + 3 * i, *e = ends + 3 * i;
		float pl[4] = {0, 0, 0, 0}; int m0 = 0, m1 = 0;
		if (r.planeRef >= 0) { memcpy(pl, planes + 4 * (size_t)r.planeRef, 16); m0 = meta[2 * (size_t)r.planeRef]; m1 = meta[2 * (size_t)r.planeRef + 1]; }
		const int cont = r.hitId >= 0 ? contents[r.hitId] : 0;
		float p[3];
		if (r.fraction == 1.0f) { p[0] = e[0]; p[1] = e[1]; p[2] = e[2]; }
		else for (int k = 0; k < 3; k++) p[k] = s[k] + r.fraction * (e[k] - s[k]);
		uint32_t w[14];
		w[0] = r.flags & 1; w[1] = (r.flags >> 1) & 1; memcpy(&w[2], &r.fraction, 4); memcpy(&w[3], p, 12); memcpy(&w[6], pl, 16);
		w[10] = (unsigned)m1 & 0xffffu; w[11] = (unsigned)m0; w[12] = (unsigned)cont; w[13] = 0;
		char *o = out + 56 * i;
		if (nt & 1) { for (int k = 0; k < 7; k++) { long long v; memcpy(&v, &w[2 * k], 8); _mm_stream_si64((long long *)(o + 8 * k), v); } }
		else memcpy(o, w, 56);
	}
	if (nt & 1) _mm_sfence();
}
int main(int argc, char **argv) {
	size_t n = argc > 1 ? atol(argv[1]) : 4u << 20;
	int T = argc > 2 ? atoi(argv[2]) : 8, nt = argc > 3 ? atoi(argv[3]) : 1;
	const int nSides = 130000, nBrushes = 20000;
	std::vector<C16> c(n); std::vector<float> st(3 * n), en(3 * n), planes(4 * nSides); std::vector<int> meta(2 * nSides), cont(nBrushes);
	char *out = (char *)aligned_alloc(64, 56 * n);
	srand(1);
	for (size_t i = 0; i < n; i++) { bool hit = rand() % 100 < 48; c[i] = {hit ? 0.5f : 1.0f, 0, hit ? rand() % nSides : -1, hit ? rand() % nBrushes : -1}; }
	for (auto &x : st) x = 1.5f; for (auto &x : en) x = 3.25f;
	memset(out, 0, 56 * n);
	for (int rep = 0; rep < 4; rep++) {
		auto t0 = std::chrono::steady_clock::now();
		std::vector<std::thread> th;
		for (int t = 0; t < T; t++) th.emplace_back(expandRange, n * t / T, n * (t + 1) / T, c.data(), st.data(), en.data(), planes.data(), meta.data(), cont.data(), out, nt);
		for (auto &x : th) x.join();
		double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
		printf("n=%zu T=%d nt=%d: %.2f ms  %.0f Mrec/s  %.1f GB/s written\n", n, T, nt, ms, n / ms / 1e3, 56.0 * n / ms / 1e6);
	}
}
