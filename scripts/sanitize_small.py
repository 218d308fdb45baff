"""Small launches of every kernel path added in round 1's last stretch (pool kernel variants with node tables and two-phase
results at an odd size, leaf queries, entity clusters, visibility), all variants compared with each other.  Written for
`compute-sanitizer --tool memcheck python scripts/sanitize_small.py`; the tool is closed on this pool, so it serves as a
quick all-paths check."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, R); sys.path.insert(0, os.path.join(R, "tests"))
import numpy as np, torch
import leaf_cases as lc
from helpers import get_map
from threecore_b200 import bspgen
from threecore_b200.engine import Engine

eng = Engine(0)
dev = torch.device("cuda", 0)
m = lc.vis_map("arena_small_vis")
eng.upload(eng.load_bsp(m.data))
# pool kernel with node tables, two-phase results (by item and by slot), odd size; then the plain variants
n = 70001
s, e = bspgen.make_sweeps(m, n, 9)
ds, de = torch.from_numpy(s).to(dev), torch.from_numpy(e).to(dev)
outs = []
for opts in (dict(pool_min_items=1, simple_max_items=0, bin_min_items=1), dict(compact_by_item=0), dict(node_thresholds=0), dict(two_phase=0),
             dict(pool=0, two_phase_small=1), dict(simple_max_items=1 << 30, two_phase_small=1), dict(bin=0, pool=1)):
    for k, v in opts.items():
        eng.set_option(k, v)
    o = torch.zeros((n, 56), dtype=torch.uint8, device=dev)
    hit = torch.zeros(n, dtype=torch.int32, device=dev)
    eng.box_trace_device(ds, de, bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, o, hit_ids=hit)
    torch.cuda.synchronize(); eng.check_status()
    outs.append(o)
assert all(torch.equal(outs[0], o) for o in outs[1:])
# leaf queries, entity clusters, visibility
pts = lc.viewpoints(m, 5000, 1)
eng.point_leafnum(pts)
mins, maxs = lc.query_boxes(m, 3000, 2)
for ls in (0, 1, 7, 128):
    eng.box_leafnums(mins, maxs, ls)
ents = eng.entity_clusters(mins, maxs)
eng.adjust_area_portal_state(0, 1, True)
eng.visible_from_points(pts[:77], ents[:333])
eng.visible_from_points(pts[:3], ents[:1])
print("l2 read", eng.measure_l2_read(1 << 20, 2))
print("sanitize_small ok")
