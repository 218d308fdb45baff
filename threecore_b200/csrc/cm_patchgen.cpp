// cm_patchgen.cpp -- placeholder until the facet generator lands.
#include "cm_patchgen.h"

#include <stdio.h>

namespace cmgpu {

bool generatePatchCollide(int, int, const float *, PatchCollide *, char *err, size_t errlen) {
	if (err && errlen) snprintf(err, errlen, "CM_GeneratePatchCollide: patch surfaces are not supported by this build yet");
	return false;
}

}  // namespace cmgpu
