// cm_load.cpp -- host-side map loader: IBSP v46 bytes -> cmgpu_flatmap.
//
// Does for the batched engine what CM_LoadMap does for the reference
// (code/qcommon/cm_load.c:606-745): checks the header and lump table, reads the
// twelve lumps the clip map consumes, derives per-plane type/signbits and
// per-brush bounds, builds the inline-model pseudo leafs and appends the
// box-hull extras of CM_InitBoxHull (cm_load.c:873-916).  Instead of a pointer
// graph on the hunk it produces the index arrays of include/cm_gpu.h, ready for
// cmgpu_upload.  Errors the reference raises with Com_Error(ERR_DROP) become
// CMGPU_ERR_BAD_MAP plus a message.
#include "../../include/cm_gpu.h"
#include "cm_patchgen.h"

#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <new>
#include <stdexcept>
#include <stdint.h>
#include <vector>

namespace {

// on-disk records, code/qcommon/qfiles.h:173-299
constexpr int kBspVersion = 46;
enum {
	LUMP_ENTITIES, LUMP_SHADERS, LUMP_PLANES, LUMP_NODES, LUMP_LEAFS, LUMP_LEAFSURFACES, LUMP_LEAFBRUSHES,
	LUMP_MODELS, LUMP_BRUSHES, LUMP_BRUSHSIDES, LUMP_DRAWVERTS, LUMP_DRAWINDEXES, LUMP_FOGS, LUMP_SURFACES,
	LUMP_LIGHTMAPS, LUMP_LIGHTGRID, LUMP_VISIBILITY, HEADER_LUMPS
};
struct Lump { int32_t ofs, len; };
struct DHeader { int32_t ident, version; Lump lumps[HEADER_LUMPS]; };
struct DShader { char name[64]; int32_t surfaceFlags, contentFlags; };
struct DPlane { float normal[3], dist; };
struct DNode { int32_t plane, children[2], mins[3], maxs[3]; };
struct DLeaf { int32_t cluster, area, mins[3], maxs[3], firstLeafSurface, numLeafSurfaces, firstLeafBrush, numLeafBrushes; };
struct DBrushSide { int32_t plane, shader; };
struct DBrush { int32_t firstSide, numSides, shader; };
struct DModel { float mins[3], maxs[3]; int32_t firstSurface, numSurfaces, firstBrush, numBrushes; };
struct DrawVert { float xyz[3], st[2], lightmap[2], normal[3]; uint8_t color[4]; };
struct DSurface {
	int32_t shader, fog, surfaceType, firstVert, numVerts, firstIndex, numIndexes, lightmapNum, lmX, lmY, lmW, lmH;
	float lmOrigin[3], lmVecs[3][3];
	int32_t patchWidth, patchHeight;
};
static_assert(sizeof(DHeader) == 144 && sizeof(DShader) == 72 && sizeof(DPlane) == 16 && sizeof(DNode) == 36 &&
              sizeof(DLeaf) == 48 && sizeof(DBrushSide) == 8 && sizeof(DBrush) == 12 && sizeof(DModel) == 40 &&
              sizeof(DrawVert) == 44 && sizeof(DSurface) == 104, "IBSP record sizes");

constexpr int MST_PATCH = 2;
constexpr int CONTENTS_BODY = 0x2000000;
constexpr int kMaxPatchVerts = 1024;   // MAX_PATCH_VERTS, cm_load.c:513

struct OwnedMap {
	cmgpu_flatmap c;          // must stay first: cmgpu_flatmap_free casts back
	std::vector<float> planes;
	std::vector<uint8_t> planeType, planeSignbits;
	std::vector<int32_t> nodes, leafs, leafClusterArea, leafbrushes, leafsurfaces, brushes, brushsides, modelLeafs,
	                     surf2patch, patches, patchPlaneSignbits, facets, borders;
	std::vector<float> brushBounds, modelBounds, patchBounds, patchPlanes;
	std::vector<uint8_t> visibility;
};

struct Fail {
	char *err; size_t errlen;
	int operator()(const char *fmt, ...) const {
		if (err && errlen) {
			va_list ap;
			va_start(ap, fmt);
			vsnprintf(err, errlen, fmt, ap);
			va_end(ap);
		}
		return CMGPU_ERR_BAD_MAP;
	}
};

template <typename T>
int lumpView(const uint8_t *base, const Lump &l, const char *what, const Fail &fail, const T **out, int *count, int minCount) {
	if (l.len % (int)sizeof(T)) return fail("%s: funny lump size", what);
	*count = l.len / (int)sizeof(T);
	if (*count < minCount) return fail("%s: map with no %s", what, what);
	*out = reinterpret_cast<const T *>(base + l.ofs);
	return CMGPU_OK;
}

}  // namespace

static int flatmapFromBsp(const void *bsp, size_t len, cmgpu_flatmap **out, char *err, size_t errlen, int tessDevice) {
	Fail fail{err, errlen};
	if (err && errlen) err[0] = 0;
	if (!out) return CMGPU_ERR_INVALID;
	*out = nullptr;
	if (!bsp) return fail("CM_LoadMap: NULL buffer");
	if (len < sizeof(DHeader)) return fail("CM_LoadMap: truncated header");
	const uint8_t *base = static_cast<const uint8_t *>(bsp);
	DHeader h;
	memcpy(&h, base, sizeof(h));
	if (h.version != kBspVersion) return fail("CM_LoadMap: wrong version number (%i should be %i)", h.version, kBspVersion);
	for (int i = 0; i < HEADER_LUMPS; i++) {
		const int64_t o = h.lumps[i].ofs, l = h.lumps[i].len;
		if (o < 0 || l < 0 || o + l > (int64_t)len) return fail("CM_LoadMap: wrong lump[%i] size/offset", i);
		// the records are read in place as 4-byte words (the reference casts the same way, cm_load.c:104-580, and the
		// tools that write .bsp files pad every lump to 4 bytes)
		if (l > 0 && ((o & 3) || (reinterpret_cast<uintptr_t>(base) & 3))) return fail("CM_LoadMap: lump[%i] is not 4-byte aligned", i);
	}

	const DShader *shaders = nullptr; const DLeaf *dleafs = nullptr; const int32_t *dleafbrushes = nullptr, *dleafsurfaces = nullptr; const DPlane *dplanes = nullptr;
	const DBrushSide *dsides = nullptr; const DBrush *dbrushes = nullptr; const DModel *dmodels = nullptr; const DNode *dnodes = nullptr;
	const DSurface *dsurfs = nullptr; const DrawVert *dverts = nullptr;
	int nShaders = 0, nLeafs = 0, nLeafBrushes = 0, nLeafSurfaces = 0, nPlanes = 0, nSides = 0, nBrushes = 0, nModels = 0, nNodes = 0, nSurfs = 0, nVerts = 0;
	int rc;
	if ((rc = lumpView(base, h.lumps[LUMP_SHADERS], "shaders", fail, &shaders, &nShaders, 1))) return rc;
	if ((rc = lumpView(base, h.lumps[LUMP_LEAFS], "leafs", fail, &dleafs, &nLeafs, 1))) return rc;
	if ((rc = lumpView(base, h.lumps[LUMP_LEAFBRUSHES], "leafbrushes", fail, &dleafbrushes, &nLeafBrushes, 0))) return rc;
	if ((rc = lumpView(base, h.lumps[LUMP_LEAFSURFACES], "leafsurfaces", fail, &dleafsurfaces, &nLeafSurfaces, 0))) return rc;
	if ((rc = lumpView(base, h.lumps[LUMP_PLANES], "planes", fail, &dplanes, &nPlanes, 1))) return rc;
	if ((rc = lumpView(base, h.lumps[LUMP_BRUSHSIDES], "brushsides", fail, &dsides, &nSides, 0))) return rc;
	if ((rc = lumpView(base, h.lumps[LUMP_BRUSHES], "brushes", fail, &dbrushes, &nBrushes, 0))) return rc;
	if ((rc = lumpView(base, h.lumps[LUMP_MODELS], "models", fail, &dmodels, &nModels, 1))) return rc;
	if ((rc = lumpView(base, h.lumps[LUMP_NODES], "nodes", fail, &dnodes, &nNodes, 1))) return rc;
	if ((rc = lumpView(base, h.lumps[LUMP_SURFACES], "surfaces", fail, &dsurfs, &nSurfs, 0))) return rc;
	if ((rc = lumpView(base, h.lumps[LUMP_DRAWVERTS], "drawverts", fail, &dverts, &nVerts, 0))) return rc;
	if (nModels > CMGPU_MAX_SUBMODELS) return fail("CMod_LoadSubmodels: MAX_SUBMODELS exceeded");

	OwnedMap *m = new OwnedMap();
	memset(&m->c, 0, sizeof(m->c));
	struct Guard { OwnedMap *p; ~Guard() { delete p; } } guard{m};

	// planes (+12 for the box hull), CMod_LoadPlanes cm_load.c:330-365
	m->planes.assign(4 * (size_t)(nPlanes + 12), 0.0f);
	m->planeType.assign((size_t)nPlanes + 12, 0);
	m->planeSignbits.assign((size_t)nPlanes + 12, 0);
	for (int i = 0; i < nPlanes; i++) {
		const float *n = dplanes[i].normal;
		int bits = 0;
		for (int j = 0; j < 3; j++) {
			m->planes[4 * (size_t)i + j] = n[j];
			if (n[j] < 0) bits |= 1 << j;
		}
		m->planes[4 * (size_t)i + 3] = dplanes[i].dist;
		// PlaneTypeForNormal, q_shared.h:963: axial only for a component exactly 1.0
		m->planeType[i] = n[0] == 1.0f ? 0 : (n[1] == 1.0f ? 1 : (n[2] == 1.0f ? 2 : 3));
		m->planeSignbits[i] = (uint8_t)bits;
	}

	// brush sides (+6), CMod_LoadBrushSides cm_load.c:449-476
	m->brushsides.assign(2 * (size_t)(nSides + 6), 0);
	for (int i = 0; i < nSides; i++) {
		const int pl = dsides[i].plane, sh = dsides[i].shader;
		if (pl < 0 || pl >= nPlanes) return fail("CMod_LoadBrushSides: bad planeNum: %i", pl);
		if (sh < 0 || sh >= nShaders) return fail("CMod_LoadBrushSides: bad shaderNum: %i", sh);
		m->brushsides[2 * (size_t)i] = pl;
		m->brushsides[2 * (size_t)i + 1] = shaders[sh].surfaceFlags;
	}

	// brushes (+1), CMod_LoadBrushes cm_load.c:249-281 and CM_BoundBrush :230-239
	m->brushes.assign(3 * (size_t)(nBrushes + 1), 0);
	m->brushBounds.assign(6 * (size_t)(nBrushes + 1), 0.0f);
	for (int i = 0; i < nBrushes; i++) {
		const DBrush &b = dbrushes[i];
		if (b.shader < 0 || b.shader >= nShaders) return fail("CMod_LoadBrushes: bad shaderNum: %i", b.shader);
		if (b.numSides < 0 || b.firstSide < 0 || (int64_t)b.firstSide + b.numSides > nSides)
			return fail("CMod_LoadBrushes: brush %i: bad side range", i);
		if (b.numSides < 6) return fail("CMod_LoadBrushes: brush %i has %i sides (the six axial sides must come first)", i, b.numSides);
		m->brushes[3 * (size_t)i] = shaders[b.shader].contentFlags;
		m->brushes[3 * (size_t)i + 1] = b.firstSide;
		m->brushes[3 * (size_t)i + 2] = b.numSides;
		float *bb = &m->brushBounds[6 * (size_t)i];
		auto sideDist = [&](int s) { return m->planes[4 * (size_t)dsides[b.firstSide + s].plane + 3]; };
		bb[0] = -sideDist(0); bb[3] = sideDist(1);
		bb[1] = -sideDist(2); bb[4] = sideDist(3);
		bb[2] = -sideDist(4); bb[5] = sideDist(5);
	}

	// leafs, CMod_LoadLeafs cm_load.c:289-322
	m->leafs.resize(4 * (size_t)nLeafs);
	m->leafClusterArea.resize(2 * (size_t)nLeafs);
	for (int i = 0; i < nLeafs; i++) {
		m->leafs[4 * (size_t)i] = dleafs[i].firstLeafBrush;
		m->leafs[4 * (size_t)i + 1] = dleafs[i].numLeafBrushes;
		m->leafs[4 * (size_t)i + 2] = dleafs[i].firstLeafSurface;
		m->leafs[4 * (size_t)i + 3] = dleafs[i].numLeafSurfaces;
		m->leafClusterArea[2 * (size_t)i] = dleafs[i].cluster;
		m->leafClusterArea[2 * (size_t)i + 1] = dleafs[i].area;
		if (dleafs[i].cluster >= m->c.numClusters) m->c.numClusters = dleafs[i].cluster + 1;   // cm_load.c:313-316
		if (dleafs[i].area >= m->c.numAreas) m->c.numAreas = dleafs[i].area + 1;
	}
	if (m->c.numAreas > CMGPU_MAX_AREAS) return fail("CMod_LoadLeafs: more than %i areas", CMGPU_MAX_AREAS);

	// visibility, CMod_LoadVisibility cm_load.c:495-515.  Without a lump every cluster sees every cluster
	// (a single row of 0xff, :503-505), which the device expresses as vised = 0.
	{
		const Lump &vl = h.lumps[LUMP_VISIBILITY];
		if (vl.len) {
			if (vl.len < 8) return fail("CMod_LoadVisibility: truncated lump");
			int32_t hdr[2];
			memcpy(hdr, base + vl.ofs, 8);
			if (hdr[0] < 0 || hdr[1] < 0 || (int64_t)hdr[0] * hdr[1] > (int64_t)vl.len - 8)
				return fail("CMod_LoadVisibility: %i clusters of %i bytes do not fit the lump", hdr[0], hdr[1]);
			m->c.vised = 1;
			m->c.numClusters = hdr[0];
			m->c.clusterBytes = hdr[1];
			m->visibility.assign(base + vl.ofs + 8, base + vl.ofs + vl.len);
		} else {
			m->c.clusterBytes = (m->c.numClusters + 31) & ~31;
		}
	}

	// leafbrushes (+1 box entry, then the submodel lists), cm_load.c:373-395 and CMod_CheckLeafBrushes :428-440
	m->leafbrushes.assign(dleafbrushes, dleafbrushes + nLeafBrushes);
	for (int i = 0; i < nLeafBrushes; i++) {
		if (m->leafbrushes[i] < 0 || m->leafbrushes[i] >= nBrushes) m->leafbrushes[i] = 0;
	}
	m->leafbrushes.push_back(nBrushes);           // box_model.leaf, cm_load.c:888-891
	m->leafsurfaces.assign(dleafsurfaces, dleafsurfaces + nLeafSurfaces);

	// inline models, CMod_LoadSubmodels cm_load.c:133-184
	m->modelLeafs.assign(4 * (size_t)nModels, 0);
	m->modelBounds.assign(6 * (size_t)nModels, 0.0f);
	for (int i = 0; i < nModels; i++) {
		const DModel &dm = dmodels[i];
		for (int j = 0; j < 3; j++) {
			m->modelBounds[6 * (size_t)i + j] = dm.mins[j] - 1;
			m->modelBounds[6 * (size_t)i + 3 + j] = dm.maxs[j] + 1;
		}
		if (i == 0) continue;
		if (dm.numBrushes < 0 || dm.numSurfaces < 0) return fail("CMod_LoadSubmodels: model %i: negative count", i);
		if (dm.firstBrush < 0 || (int64_t)dm.firstBrush + dm.numBrushes > nBrushes)
			return fail("CMod_LoadSubmodels: model %i: brushes %i..+%i outside the map's %i", i, dm.firstBrush, dm.numBrushes, nBrushes);
		if (dm.firstSurface < 0 || (int64_t)dm.firstSurface + dm.numSurfaces > nSurfs)
			return fail("CMod_LoadSubmodels: model %i: surfaces %i..+%i outside the map's %i", i, dm.firstSurface, dm.numSurfaces, nSurfs);
		m->modelLeafs[4 * (size_t)i] = (int32_t)m->leafbrushes.size();
		m->modelLeafs[4 * (size_t)i + 1] = dm.numBrushes;
		for (int j = 0; j < dm.numBrushes; j++) m->leafbrushes.push_back(dm.firstBrush + j);
		m->modelLeafs[4 * (size_t)i + 2] = (int32_t)m->leafsurfaces.size();
		m->modelLeafs[4 * (size_t)i + 3] = dm.numSurfaces;
		for (int j = 0; j < dm.numSurfaces; j++) m->leafsurfaces.push_back(dm.firstSurface + j);
	}

	// nodes, CMod_LoadNodes cm_load.c:193-221
	m->nodes.resize(3 * (size_t)nNodes);
	for (int i = 0; i < nNodes; i++) {
		m->nodes[3 * (size_t)i] = dnodes[i].plane;
		m->nodes[3 * (size_t)i + 1] = dnodes[i].children[0];
		m->nodes[3 * (size_t)i + 2] = dnodes[i].children[1];
	}

	// patches, CMod_LoadPatches cm_load.c:514-580.  The facets of every MST_PATCH surface are generated on the host, one
	// patch after the other, or -- tessDevice >= 0 -- on that CUDA device, one thread per patch (cm_gpu.cu); either way by
	// the builder of cm_patchgen_core.h, and appended in surface order.
	m->surf2patch.assign((size_t)nSurfs, -1);
	std::vector<int> patchSurf, pw, ph, pfirst;
	for (int i = 0; i < nSurfs; i++) {
		const DSurface &s = dsurfs[i];
		if (s.surfaceType != MST_PATCH) continue;
		const int w = s.patchWidth, hgt = s.patchHeight;
		const int64_t c = (int64_t)w * hgt;
		if (c > kMaxPatchVerts) return fail("CMod_LoadPatches: MAX_PATCH_VERTS");
		if (w < 0 || hgt < 0 || s.firstVert < 0 || (int64_t)s.firstVert + c > nVerts) return fail("CMod_LoadPatches: surface %i: bad vertex range", i);
		if (s.shader < 0 || s.shader >= nShaders) return fail("CMod_LoadPatches: bad shaderNum: %i", s.shader);
		patchSurf.push_back(i); pw.push_back(w); ph.push_back(hgt); pfirst.push_back(s.firstVert);
	}
	std::vector<cmgpu::PatchCollide> pcs(patchSurf.size());
	char perr[256];
	if (tessDevice >= 0 && !patchSurf.empty()) {
		// the parameter checks of CM_GeneratePatchCollide come first, in surface order, like on the host
		std::vector<float> allPts(3 * (size_t)nVerts);
		for (int v = 0; v < nVerts; v++) memcpy(&allPts[3 * (size_t)v], dverts[v].xyz, 12);
		if (!cmgpu::generatePatchCollidesOnDevice(tessDevice, (int)patchSurf.size(), pw.data(), ph.data(), pfirst.data(), allPts.data(),
		                                          (size_t)nVerts, pcs.data(), perr, sizeof(perr))) return fail("%s", perr);
	} else {
		for (size_t k = 0; k < patchSurf.size(); k++) {
			const int64_t c = (int64_t)pw[k] * ph[k];
			std::vector<float> pts(3 * (size_t)c);
			for (int64_t j = 0; j < c; j++) memcpy(&pts[3 * (size_t)j], dverts[pfirst[k] + j].xyz, 12);
			if (!cmgpu::generatePatchCollide(pw[k], ph[k], pts.data(), &pcs[k], perr, sizeof(perr))) return fail("%s", perr);
		}
	}
	for (size_t k = 0; k < patchSurf.size(); k++) {
		const int i = patchSurf[k];
		const DSurface &s = dsurfs[i];
		const cmgpu::PatchCollide &pc = pcs[k];
		const int ord = (int)(m->patches.size() / 6);
		m->surf2patch[i] = ord;
		const int firstPlane = (int)(m->patchPlanes.size() / 4), firstFacet = (int)(m->facets.size() / 3);
		m->patches.insert(m->patches.end(), {shaders[s.shader].surfaceFlags, shaders[s.shader].contentFlags,
		                                      firstPlane, (int32_t)pc.planes.size() / 4, firstFacet, (int32_t)pc.facets.size()});
		m->patchBounds.insert(m->patchBounds.end(), pc.bounds, pc.bounds + 6);
		m->patchPlanes.insert(m->patchPlanes.end(), pc.planes.begin(), pc.planes.end());
		m->patchPlaneSignbits.insert(m->patchPlaneSignbits.end(), pc.signbits.begin(), pc.signbits.end());
		for (const cmgpu::PatchFacet &f : pc.facets) {
			m->facets.insert(m->facets.end(), {f.surfacePlane, f.numBorders, (int32_t)(m->borders.size() / 2)});
			for (int j = 0; j < f.numBorders; j++) {
				m->borders.push_back(f.borderPlanes[j]);
				m->borders.push_back(f.borderInward[j]);
			}
		}
	}

	// box hull, CM_InitBoxHull cm_load.c:873-916
	for (int i = 0; i < 6; i++) {
		const int side = i & 1, axis = i >> 1;
		m->brushsides[2 * (size_t)(nSides + i)] = nPlanes + i * 2 + side;
		m->brushsides[2 * (size_t)(nSides + i) + 1] = 0;
		const size_t p0 = (size_t)nPlanes + (size_t)i * 2, p1 = p0 + 1;
		m->planes[4 * p0 + axis] = 1.0f;
		m->planeType[p0] = (uint8_t)axis;
		m->planeSignbits[p0] = 0;
		m->planes[4 * p1 + axis] = -1.0f;
		m->planeType[p1] = (uint8_t)(3 + axis);
		m->planeSignbits[p1] = (uint8_t)(1 << axis);
	}
	m->brushes[3 * (size_t)nBrushes] = CONTENTS_BODY;
	m->brushes[3 * (size_t)nBrushes + 1] = nSides;
	m->brushes[3 * (size_t)nBrushes + 2] = 6;

	cmgpu_flatmap &c = m->c;
	c.numPlanes = nPlanes + 12; c.planes = m->planes.data(); c.planeType = m->planeType.data(); c.planeSignbits = m->planeSignbits.data();
	c.numNodes = nNodes; c.nodes = m->nodes.data();
	c.numLeafs = nLeafs; c.leafs = m->leafs.data(); c.leafClusterArea = m->leafClusterArea.data();
	c.numLeafBrushes = (int32_t)m->leafbrushes.size(); c.leafbrushes = m->leafbrushes.data();
	c.numLeafSurfaces = (int32_t)m->leafsurfaces.size(); c.leafsurfaces = m->leafsurfaces.data();
	c.numBrushes = nBrushes + 1; c.brushes = m->brushes.data(); c.brushBounds = m->brushBounds.data();
	c.numBrushSides = nSides + 6; c.brushsides = m->brushsides.data();
	c.numModels = nModels; c.modelLeafs = m->modelLeafs.data(); c.modelBounds = m->modelBounds.data();
	c.numSurfaces = nSurfs; c.surf2patch = m->surf2patch.data();
	c.numPatches = (int32_t)(m->patches.size() / 6); c.patches = m->patches.data(); c.patchBounds = m->patchBounds.data();
	c.numPatchPlanes = (int32_t)(m->patchPlanes.size() / 4); c.patchPlanes = m->patchPlanes.data();
	c.patchPlaneSignbits = m->patchPlaneSignbits.data();
	c.numFacets = (int32_t)(m->facets.size() / 3); c.facets = m->facets.data();
	c.numBorders = (int32_t)(m->borders.size() / 2); c.borders = m->borders.data();
	c.visibility = m->visibility.empty() ? nullptr : m->visibility.data();
	c.boxBrush = nBrushes; c.boxFirstPlane = nPlanes; c.boxFirstSide = nSides; c.boxLeafBrush = nLeafBrushes;

	guard.p = nullptr;
	*out = &m->c;
	return CMGPU_OK;
}

// nothing may leave an extern "C" function by exception: a corrupt file that asks for absurd sizes is an error code
static int flatmapFromBspGuarded(const void *bsp, size_t len, cmgpu_flatmap **out, char *err, size_t errlen, int tessDevice) {
	try {
		return flatmapFromBsp(bsp, len, out, err, errlen, tessDevice);
	} catch (const std::bad_alloc &) {
		if (out) *out = nullptr;
		if (err && errlen) snprintf(err, errlen, "CM_LoadMap: out of memory");
		return CMGPU_ERR_NOMEM;
	} catch (const std::exception &e) {
		if (out) *out = nullptr;
		if (err && errlen) snprintf(err, errlen, "CM_LoadMap: %s", e.what());
		return CMGPU_ERR_BAD_MAP;
	}
}

extern "C" int cmgpu_flatmap_from_bsp(const void *bsp, size_t len, cmgpu_flatmap **out, char *err, size_t errlen) {
	return flatmapFromBspGuarded(bsp, len, out, err, errlen, -1);
}

extern "C" int cmgpu_flatmap_from_bsp_device(int device, const void *bsp, size_t len, cmgpu_flatmap **out, char *err, size_t errlen) {
	if (device < 0) {
		if (err && errlen) snprintf(err, errlen, "cmgpu_flatmap_from_bsp_device: bad device");
		return CMGPU_ERR_INVALID;
	}
	return flatmapFromBspGuarded(bsp, len, out, err, errlen, device);
}

extern "C" void cmgpu_flatmap_free(cmgpu_flatmap *map) {
	if (map) delete reinterpret_cast<OwnedMap *>(map);
}
