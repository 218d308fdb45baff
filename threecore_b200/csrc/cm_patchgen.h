// cm_patchgen.h -- load-time Bezier patch -> collision facets: the loader's view (growable arrays).
// Counterpart of CM_GeneratePatchCollide, code/qcommon/cm_patch.c:1136-1203.  The builder itself is
// cm_patchgen_core.h (host and device); generatePatchCollide runs it on the host, generatePatchCollidesOnDevice
// (cm_gpu.cu) runs it as one GPU thread per patch.
#pragma once

#include <stddef.h>
#include <stdint.h>

#include <vector>

namespace cmgpu {

constexpr int kMaxFacetBorders = 4 + 6 + 16;   // facet_t, cm_patch.h:31-37

struct PatchFacet {
	int32_t surfacePlane;
	int32_t numBorders;
	int32_t borderPlanes[kMaxFacetBorders];
	int32_t borderInward[kMaxFacetBorders];
	int32_t borderNoAdjust[kMaxFacetBorders];
};

struct PatchCollide {
	float bounds[6];                 // mins, maxs (already spread by 1, cm_patch.c:1190-1196)
	std::vector<float> planes;       // [n][4]
	std::vector<int32_t> signbits;   // [n]
	std::vector<PatchFacet> facets;
};

namespace pg { struct PatchScratch; }
void patchCollideFromScratch(const pg::PatchScratch &s, PatchCollide *out);

// points: [height*width][3] control grid, row-major as stored in the BSP (cm_load.c:567-572).
// Returns false with a message in err for the conditions the reference raises ERR_DROP on.
bool generatePatchCollide(int width, int height, const float *points, PatchCollide *out, char *err, size_t errlen);

// The same for n patches at once on CUDA device `device`, one thread per patch (patchCollideKernel).  widths / heights:
// [n]; firstPoint[i]: index of patch i's first control point in `points` ([*][3] floats).  out: n results.  Returns false
// with a message for the reference's ERR_DROP conditions (the first failing patch) and for CUDA failures.
bool generatePatchCollidesOnDevice(int device, int n, const int *widths, const int *heights, const int *firstPoint,
                                   const float *points, size_t numPoints, PatchCollide *out, char *err, size_t errlen);

}  // namespace cmgpu
