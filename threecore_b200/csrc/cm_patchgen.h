// cm_patchgen.h -- load-time Bezier patch -> collision facets (host only).
// Counterpart of CM_GeneratePatchCollide, code/qcommon/cm_patch.c:1136-1203.
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

// points: [height*width][3] control grid, row-major as stored in the BSP (cm_load.c:567-572).
// Returns false with a message in err for the conditions the reference raises ERR_DROP on.
bool generatePatchCollide(int width, int height, const float *points, PatchCollide *out, char *err, size_t errlen);

}  // namespace cmgpu
