"""Generates tests/golden/leaf_queries.npz with the UNMODIFIED reference (oracle/_ref/libcmref.so): CM_PointLeafnum,
CM_BoxLeafnums, what SV_LinkEntity stores (clusters, areas), CM_AdjustAreaPortalState + the PVS stage of the snapshot
code.  Run where /root/reference exists:  python tests/golden/make_golden_leaf.py"""
import os
import sys
import zlib

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import leaf_cases as lc  # noqa: E402
from helpers import cmref  # noqa: E402

KIND = "arena_small_vis"
LISTSIZE = 48


def main():
    m = lc.vis_map(KIND)
    r = cmref.RefCM().load(m.data)
    pts = lc.viewpoints(m, 3000, 101)
    mins, maxs = lc.query_boxes(m, 1500, 102)
    lists, counts, last = r.box_leafnums(mins, maxs, LISTSIZE)
    emins, emaxs, origin = lc.entities(m, 300, 103)
    r.sv_clear_world()
    n = len(origin)
    absmin, absmax, recs = np.zeros((n, 3), np.float32), np.zeros((n, 3), np.float32), np.zeros((n, 20), np.int32)
    zero = np.zeros(3, np.float32)
    for i in range(n):
        r.sv_link_entity(i, False, 0, emins[i], emaxs[i], origin[i], zero, 1)
        absmin[i], absmax[i], _ = r.sv_entity_bounds(i)
        recs[i], _ = r.sv_entity_clusters(i)
    nA = m.stats["areas"]
    portals = np.zeros((nA, nA), np.int32)
    for p, q in ((0, 1), (1, 3)):
        r.adjust_area_portal_state(p, q, True)
        portals[p, q] += 1
        portals[q, p] += 1
    views = lc.viewpoints(m, 96, 104)
    visible = r.sv_visible_from_points(views, np.arange(n))
    np.savez_compressed(os.path.join(HERE, "leaf_queries.npz"), kind=KIND, bsp_crc32=np.uint32(zlib.crc32(m.data)),
                        listsize=LISTSIZE, pts=pts, leafnums=r.point_leafnum(pts), mins=mins, maxs=maxs, lists=lists,
                        counts=counts, lastLeafs=last, absmin=absmin, absmax=absmax, ent_records=recs, portals=portals,
                        views=views, visible=visible)
    print("leaf_queries.npz:", len(pts), "points,", len(mins), "boxes,", n, "entities,", visible.sum(), "of", visible.size,
          "pairs visible")


if __name__ == "__main__":
    main()
