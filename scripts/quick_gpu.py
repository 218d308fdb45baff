import sys, time, numpy as np
import os; R=os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R); sys.path.insert(0, os.path.join(R,'tests'))
from helpers import *
from threecore_b200.engine import Engine, FlatMap
e = Engine(0)
for kind in ("tiny", "room", "arena_small"):
    m = get_map(kind)
    fm = e.load_bsp(m.data)
    flat = fm.to_dict()
    o = oracle_for(flat)
    s, en = bspgen.make_sweeps(m, 200000, 7, len_short=(1, 300), len_long=(300, 3000))
    for hull in (None, (bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS)):
        mn, mx = hull if hull else (None, None)
        want, info = o.box_trace(s, en, mn, mx, mask=bspgen.MASK_PLAYERSOLID, want_info=True)
        t = time.time(); got, hit = e.box_trace(s, en, mn, mx, mask=bspgen.MASK_PLAYERSOLID, want_hit=True); dt = time.time() - t
        assert_traces_equal(got, want, f"{kind} hull={hull is not None}")
        wanthit = np.where(info['hitPatch'] >= 0, -2 - info['hitPatch'], info['hitBrush'])
        assert np.array_equal(hit, wanthit), "hit ids"
        print(kind, hull is not None, "OK", f"{dt*1e3:.1f} ms")
print("launches", e.launch_count)
