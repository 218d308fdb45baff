"""tests/golden/aas_trace.npz: lines through the generated AAS world of aasgen.make_world(n_boxes=120, seed=0xAA5002),
answered by the UNMODIFIED reference (botlib/be_aas_sample.c in oracle/_ref/libcmref.so).  Run where /root/reference exists:
    python tests/golden/make_golden_aas.py"""
import os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import cmref
from threecore_b200 import aasgen

w = aasgen.make_world(n_boxes=120, seed=0xAA5002, extent=2048.0, height=512.0)
r = cmref.RefCM().aas_load(w)
s, e = aasgen.make_lines(w, 6000, 0xAA11, length=(4.0, 900.0))
pt = np.random.default_rng(3).choice([aasgen.PRESENCE_NORMAL, aasgen.PRESENCE_CROUCH, aasgen.PRESENCE_NORMAL | aasgen.PRESENCE_CROUCH,
                                      aasgen.PRESENCE_NONE], size=len(s)).astype(np.int32)
tr = r.aas_trace_client_bbox(s, e, pt)
areas = r.aas_point_area_num(s)
# movement predictions on the same AAS world, the contents of the world from a small generated clip map
from threecore_b200 import bspgen
m = bspgen.arena_map(n_brushes=400, seed=0xC0FFEE21, size_xy=2048.0, size_z=512.0)
r.load(m.data)
mv_in = aasgen.make_moves(w, 4000, 0xAA12)
mv, res = r.aas_predict_client_movement(mv_in)
np.savez_compressed(os.path.join(HERE, "aas_trace.npz"), planes=w.planes, nodes=w.nodes, presence=w.presence, contents=w.contents, starts=s, ends=e,
                    presencetypes=pt, traces=np.ascontiguousarray(tr).view(np.uint8).reshape(-1, 36), areas=areas,
                    bsp=np.frombuffer(m.data, np.uint8), moves=np.ascontiguousarray(mv).view(np.uint8).reshape(-1, 84), move_results=res,
                    **{"mv_" + k: v for k, v in mv_in.items()})
print("moves:", sorted({int(x): int((mv["stopevent"] == x).sum()) for x in np.unique(mv["stopevent"])}.items()), "qfalse", int((res == 0).sum()))
print("wrote", len(s), "lines;", w.stats, "hit", float((tr["fraction"] < 1).mean()))
