/* oracle/aas_shim.c -- TEST INFRASTRUCTURE.
 *
 * Lets the UNMODIFIED code/botlib/be_aas_sample.c (AAS_TraceClientBBox :399-665, AAS_PointAreaNum :212-258) run outside
 * the engine: defines the globals it reads (aasworld, botimport, botDeveloper), stubs the functions its OTHER
 * functions call (never reached from the two we call with passent < 0), and adds a batch API for the tests.  Compiled
 * by oracle/Makefile with the reference's headers; linked into oracle/_ref/libcmref.so next to be_aas_sample.o.
 * An AAS world is handed over as plain arrays (the generator is threecore_b200/aasgen.py: the reference ships neither
 * .aas files nor the BSPC tool that makes them).
 */
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../qcommon/q_shared.h"
#include "l_memory.h"
#include "l_script.h"
#include "l_precomp.h"
#include "l_struct.h"
#include "l_libvar.h"
#include "aasfile.h"
#include "botlib.h"
#include "be_aas.h"
#include "be_interface.h"
#include "be_aas_funcs.h"
#include "be_aas_def.h"

#define EXPORT __attribute__((visibility("default")))

aas_t aasworld;
botlib_import_t botimport;
int botDeveloper;

static void QDECL shimPrint(int type, const char *fmt, ...) { (void)type; (void)fmt; }

/* referenced by functions of be_aas_sample.c that the tests never call */
int AAS_AreaReachability(int areanum) { (void)areanum; return 0; }
qboolean AAS_EntityCollision(int entnum, vec3_t start, vec3_t boxmins, vec3_t boxmaxs, vec3_t end, int contentmask, bsp_trace_t *trace) {
	(void)entnum; (void)start; (void)boxmins; (void)boxmaxs; (void)end; (void)contentmask; (void)trace;
	return qfalse;
}
void FreeMemory(void *ptr) { free(ptr); }
void *GetClearedHunkMemory(unsigned long size) { return calloc(1, size); }
void *GetHunkMemory(unsigned long size) { return malloc(size); }
float LibVarValue(const char *var_name, const char *value) { (void)var_name; return (float)atof(value); }   /* the default: no libvar is ever set */
/* be_aas_bspq3.c:138: botimport.PointContents, i.e. SV_PointContents(p, -1) -- the world's contents on a server without
 * entities in the way.  The clip map is the one ref_shim.c loaded (CM_PointContents answers 0 when there is none). */
int CM_PointContents(const vec3_t p, clipHandle_t model);
int AAS_PointContents(vec3_t point) { return CM_PointContents(point, 0); }
bsp_trace_t AAS_Trace(vec3_t start, vec3_t mins, vec3_t maxs, vec3_t end, int passent, int contentmask) {
	bsp_trace_t t;
	(void)start; (void)mins; (void)maxs; (void)end; (void)passent; (void)contentmask;
	memset(&t, 0, sizeof(t));
	t.fraction = 1.0f;
	return t;
}
void AAS_InitSettings(void);
int AAS_PredictClientMovement(struct aas_clientmove_s *move, int entnum, const vec3_t origin, int presencetype, int onground,
                              const vec3_t velocity, const vec3_t cmdmove, int cmdframes, int maxframes, float frametime,
                              int stopevent, int stopareanum, int visualize);

/* planes: [n][5] floats normal.xyz, dist, type-as-float; nodes: [n][3] ints planenum, children[0], children[1] (node 0 is the
 * dummy of be_aas_sample.c:225); presencetypes: [numareasettings] (area 0 is the dummy) */
EXPORT int aasref_load(int numplanes, const float *planes, int numnodes, const int *nodes, int numareasettings, const int *presencetypes,
                       const int *areacontents) {
	int i;
	free(aasworld.planes); free(aasworld.nodes); free(aasworld.areasettings);
	memset(&aasworld, 0, sizeof(aasworld));
	botimport.Print = shimPrint;
	aasworld.numplanes = numplanes;
	aasworld.planes = calloc(numplanes > 0 ? numplanes : 1, sizeof(aas_plane_t));
	for (i = 0; i < numplanes; i++) {
		aasworld.planes[i].normal[0] = planes[5 * i]; aasworld.planes[i].normal[1] = planes[5 * i + 1]; aasworld.planes[i].normal[2] = planes[5 * i + 2];
		aasworld.planes[i].dist = planes[5 * i + 3];
		aasworld.planes[i].type = (int)planes[5 * i + 4];
	}
	aasworld.numnodes = numnodes;
	aasworld.nodes = calloc(numnodes > 0 ? numnodes : 1, sizeof(aas_node_t));
	for (i = 0; i < numnodes; i++) {
		aasworld.nodes[i].planenum = nodes[3 * i];
		aasworld.nodes[i].children[0] = nodes[3 * i + 1];
		aasworld.nodes[i].children[1] = nodes[3 * i + 2];
	}
	aasworld.numareasettings = numareasettings;
	aasworld.areasettings = calloc(numareasettings > 0 ? numareasettings : 1, sizeof(aas_areasettings_t));
	for (i = 0; i < numareasettings; i++) {
		aasworld.areasettings[i].presencetype = presencetypes[i];
		aasworld.areasettings[i].contents = areacontents ? areacontents[i] : 0;
	}
	aasworld.loaded = qtrue;
	AAS_InitSettings();
	return 0;
}

EXPORT int aasref_sizeof_trace(void) { return (int)sizeof(aas_trace_t); }

/* n x AAS_TraceClientBBox(start, end, presencetype, -1) */
EXPORT void aasref_trace_client_bbox_batch(int n, const float *starts, const float *ends, const int *presencetypes, int presence_stride, aas_trace_t *out) {
	int i;
	for (i = 0; i < n; i++) {
		vec3_t s, e;
		VectorCopy(starts + 3 * i, s);
		VectorCopy(ends + 3 * i, e);
		out[i] = AAS_TraceClientBBox(s, e, presencetypes[presence_stride ? i : 0], -1);
	}
}

EXPORT void aasref_point_area_num_batch(int n, const float *points, int *out) {
	int i;
	for (i = 0; i < n; i++) {
		vec3_t p;
		VectorCopy(points + 3 * i, p);
		out[i] = AAS_PointAreaNum(p);
	}
}

EXPORT int aasref_sizeof_clientmove(void) { return (int)sizeof(aas_clientmove_t); }

/* n x AAS_PredictClientMovement(&move, -1, origin, presencetype, onground, velocity, cmdmove, cmdframes, maxframes, frametime,
 * stopevent, stopareanum, qfalse) (be_aas_move.c:965-979) */
EXPORT void aasref_predict_client_movement_batch(int n, const float *origins, const int *presencetypes, const int *ongrounds,
                                                 const float *velocities, const float *cmdmoves, const int *cmdframes, const int *maxframes,
                                                 const float *frametimes, const int *stopevents, const int *stopareanums,
                                                 aas_clientmove_t *moves, int *results) {
	int i;
	for (i = 0; i < n; i++) {
		vec3_t o, v, c;
		VectorCopy(origins + 3 * i, o);
		VectorCopy(velocities + 3 * i, v);
		VectorCopy(cmdmoves + 3 * i, c);
		results[i] = AAS_PredictClientMovement(&moves[i], -1, o, presencetypes[i], ongrounds[i], v, c, cmdframes[i], maxframes[i],
		                                       frametimes[i], stopevents[i], stopareanums[i], qfalse);
	}
}
