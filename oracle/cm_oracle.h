/*
 * oracle/cm_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Interface of the CPU restatement of the reference clip-map trace
 * (code/qcommon/cm_trace.c, cm_test.c, cm_patch.c run-time part).
 * Plain C, index-based arrays, no global state.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs
 * may use it.
 */
#ifndef CM_ORACLE_H
#define CM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CMO_BOX_MODEL_HANDLE 1023   /* cm_local.h:28 */

/* The loaded clip map as index-based arrays (the pointer graph of clipMap_t,
 * cm_local.h:120-172, with every pointer replaced by an index).  The arrays
 * INCLUDE the box-hull extras the reference appends (cm_load.c:45-50,873-916):
 * 12 planes, 6 sides, 1 brush, 1 leafbrush entry. */
typedef struct {
	int numPlanes;        const float *planes;        /* [n][4] normal, dist */
	                      const uint8_t *planeType;   /* cplane_t.type */
	                      const uint8_t *planeSignbits;
	int numNodes;         const int *nodes;           /* [n][3] plane, child0, child1 */
	int numLeafs;         const int *leafs;           /* [n][4] firstLeafBrush, numLeafBrushes, firstLeafSurface, numLeafSurfaces */
	int numLeafBrushes;   const int *leafbrushes;
	int numLeafSurfaces;  const int *leafsurfaces;
	int numBrushes;       const int *brushes;         /* [n][3] contents, firstSide, numSides */
	                      const float *brushBounds;   /* [n][6] mins, maxs */
	int numBrushSides;    const int *brushsides;      /* [n][2] plane, surfaceFlags */
	int numModels;        const int *modelLeafs;      /* [n][4] like leafs[]; entry 0 unused */
	int numSurfaces;      const int *surf2patch;      /* [n] patch ordinal or -1 (cm.surfaces[i] == NULL) */
	int numPatches;       const int *patches;         /* [n][6] surfaceFlags, contents, firstPlane, numPlanes, firstFacet, numFacets */
	                      const float *patchBounds;   /* [n][6] */
	int numPatchPlanes;   const float *patchPlanes;   /* [n][4] */
	                      const int *patchPlaneSignbits;
	int numFacets;        const int *facets;          /* [n][3] surfacePlane, numBorders, firstBorder */
	int numBorders;       const int *borders;         /* [n][2] plane (patch-local), inward */
	int boxBrush;         /* index of the temp-box brush (= numBrushes-1) */
	int boxFirstPlane;    /* index of its 12 planes (= numPlanes-12) */
	int boxFirstSide;     /* index of its 6 sides   (= numBrushSides-6) */
	int boxLeafBrush;     /* leafbrushes[] slot that names it */
	int noCurves;         /* cvar cm_noCurves */
	int playerCurveClip;  /* cvar cm_playerCurveClip */
} cmo_map_t;

/* trace_t, q_shared.h:976-986 (56 bytes) */
typedef struct {
	int   allsolid;
	int   startsolid;
	float fraction;
	float endpos[3];
	float plane_normal[3];
	float plane_dist;
	uint8_t plane_type;
	uint8_t plane_signbits;
	uint8_t plane_pad[2];
	int   surfaceFlags;
	int   contents;
	int   entityNum;
} cmo_trace_t;

/* what the reference cannot tell us: which brush/patch produced the hit and
 * how much of the map one trace touched (feeds the L2 roofline model). */
typedef struct {
	int hitBrush;        /* brush index of the recorded fraction / allsolid, -1 none */
	int hitPatch;        /* patch ordinal, -1 none */
	int nodes;           /* nodes visited */
	int leafs;           /* leafs visited */
	int brushRefs;       /* leafbrush entries read */
	int brushBounds;     /* brushes that reached the bounds test */
	int brushSides;      /* brush sides evaluated */
	int crossings;       /* sides with an enter/leave division */
	int splits;          /* nodes where the segment was split */
	int patches;         /* patches that passed the bounds test */
	int patchPlanes;     /* patch planes evaluated */
	int facets;          /* facets evaluated */
} cmo_info_t;

/* one work item; models==0 traces the world tree.  box* only read for model 1023. */
void cmo_box_trace(const cmo_map_t *m, cmo_trace_t *out, cmo_info_t *info,
                   const float *start, const float *end, const float *mins, const float *maxs,
                   int model, int mask, const float *boxmins, const float *boxmaxs);

void cmo_transformed_box_trace(const cmo_map_t *m, cmo_trace_t *out, cmo_info_t *info,
                               const float *start, const float *end, const float *mins, const float *maxs,
                               int model, int mask, const float *origin, const float *angles,
                               const float *boxmins, const float *boxmaxs);

int cmo_point_contents(const cmo_map_t *m, const float *p, int model,
                       const float *boxmins, const float *boxmaxs);
int cmo_transformed_point_contents(const cmo_map_t *m, const float *p, int model,
                                   const float *origin, const float *angles,
                                   const float *boxmins, const float *boxmaxs);
int cmo_point_leafnum(const cmo_map_t *m, const float *p);

/* leaf lists, entity linking (PVS part), areas, visibility -- see cm_oracle.c */
#define CMO_ENT_CLUSTER_INTS 21     /* numLeafs, areanum, areanum2, numClusters, lastCluster, clusternums[16] */
int  cmo_box_leafnums(const cmo_map_t *m, const int *leafClusterArea, const float *mins, const float *maxs,
                      int *list, int listsize, int *lastLeaf);
void cmo_box_leafnums_batch(const cmo_map_t *m, const int *leafClusterArea, int n, const float *mins, const float *maxs,
                            int listsize, int *lists, int *counts, int *lastLeafs);
void cmo_entity_clusters(const cmo_map_t *m, const int *leafClusterArea, const float *absmin, const float *absmax, int *out);
void cmo_entity_clusters_batch(const cmo_map_t *m, const int *leafClusterArea, int n, const float *absmins,
                               const float *absmaxs, int *out);
void cmo_flood_areas(int numAreas, const int *areaPortals, int *floodnum);
int  cmo_entity_visible(int numAreas, const int *floodnum, int numClusters, int clusterBytes, const uint8_t *vis,
                        int clientArea, int clientCluster, const int *ent);

/* batch wrappers (hull_stride / mask_stride: 0 shared, 1 per item) */
void cmo_box_trace_batch(const cmo_map_t *m, int n, const float *starts, const float *ends,
                         const float *mins, const float *maxs, int hull_stride,
                         const int *models, const int *masks, int mask_stride,
                         const float *boxmins, const float *boxmaxs,
                         cmo_trace_t *out, cmo_info_t *info);
void cmo_transformed_box_trace_batch(const cmo_map_t *m, int n, const float *starts, const float *ends,
                                     const float *mins, const float *maxs, int hull_stride,
                                     const int *models, const int *masks, int mask_stride,
                                     const float *origins, const float *angles,
                                     const float *boxmins, const float *boxmaxs,
                                     cmo_trace_t *out, cmo_info_t *info);
void cmo_point_contents_batch(const cmo_map_t *m, int n, const float *pts, const int *models,
                              const float *origins, const float *angles,
                              const float *boxmins, const float *boxmaxs, int *out);
void cmo_angle_vectors(const float *angles, float *forward, float *right, float *up);

#ifdef __cplusplus
}
#endif
#endif
