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
np.savez_compressed(os.path.join(HERE, "aas_trace.npz"), planes=w.planes, nodes=w.nodes, presence=w.presence, starts=s, ends=e,
                    presencetypes=pt, traces=np.ascontiguousarray(tr).view(np.uint8).reshape(-1, 36), areas=areas)
print("wrote", len(s), "lines;", w.stats, "hit", float((tr["fraction"] < 1).mean()))
