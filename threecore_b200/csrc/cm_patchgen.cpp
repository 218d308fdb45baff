// cm_patchgen.cpp -- load-time Bezier patch -> collision facets on the HOST: the shared builder of cm_patchgen_core.h
// (one implementation for host and device; the algorithm and its citations are described there) run in a heap-allocated
// scratch block, the result copied into growable arrays for the loader (cm_load.cpp).
// Float/double evaluation widths follow the reference expression by expression (the reference is built without FMA
// contraction; this file is compiled with -ffp-contract=off).
#include "cm_patchgen.h"

#include "cm_patchgen_core.h"

#include <stdio.h>

#include <memory>
#include <new>

namespace cmgpu {

static_assert(kMaxFacetBorders == pg::kMaxBorders, "facet_t holds 4 + 6 + 16 borders (cm_patch.h:31-37)");

void patchCollideFromScratch(const pg::PatchScratch &s, PatchCollide *out) {
	for (int k = 0; k < 6; k++) out->bounds[k] = s.bounds[k];
	out->planes.clear();
	out->signbits.clear();
	out->planes.reserve(4 * (size_t)s.numPlanes);
	for (int i = 0; i < s.numPlanes; i++) {
		out->planes.insert(out->planes.end(), s.planes[i].p, s.planes[i].p + 4);
		out->signbits.push_back(s.planes[i].signbits);
	}
	out->facets.resize((size_t)s.numFacets);
	for (int i = 0; i < s.numFacets; i++) {
		const pg::Facet &f = s.facets[i];
		PatchFacet &o = out->facets[(size_t)i];
		o.surfacePlane = f.surfacePlane;
		o.numBorders = f.numBorders;
		for (int k = 0; k < kMaxFacetBorders; k++) {
			o.borderPlanes[k] = f.borderPlanes[k];
			o.borderInward[k] = f.borderInward[k];
			o.borderNoAdjust[k] = f.borderNoAdjust[k];
		}
	}
}

bool generatePatchCollide(int width, int height, const float *points, PatchCollide *out, char *err, size_t errlen) {
	std::unique_ptr<pg::PatchScratch> s(new (std::nothrow) pg::PatchScratch);
	if (!s) {
		if (err && errlen) snprintf(err, errlen, "CM_GeneratePatchCollide: out of memory");
		return false;
	}
	pg::Builder(*s).run(width, height, points);
	if (s->error) {
		if (err && errlen) snprintf(err, errlen, "%s", pg::errorText(s->error));
		return false;
	}
	patchCollideFromScratch(*s, out);
	return true;
}

}  // namespace cmgpu
