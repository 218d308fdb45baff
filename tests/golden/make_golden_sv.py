"""Golden vectors for the device SV_Trace / SV_PointContents: made HERE by the reference's own server/sv_world.c
(compiled verbatim into oracle/_ref/libcmref.so, oracle/sv_shim.c) on the scene of tests/sv_scene.py.

    python tests/golden/make_golden_sv.py      # needs /root/reference (or a prebuilt oracle/_ref/libcmref.so)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))          # tests/
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import sv_scene                                      # noqa: E402
from helpers import cmref, get_map                   # noqa: E402


def main():
    m = get_map("arena_models")
    ref = cmref.RefCM().load(m.data)
    ents = sv_scene.make_entities(m, 300, seed=71)
    sv_scene.link_all_ref(ref, ents)
    s, e, mins, maxs, passent, masks = sv_scene.make_moves(m, ents, 4000, seed=76)
    traces = ref.sv_trace(s, e, mins, maxs, passent, masks)
    rng = np.random.default_rng(77)
    n = 4000
    tgt = rng.integers(0, len(ents["number"]), size=n)
    pts = sv_scene.f3(ents["origin"][tgt] + rng.uniform(-60, 60, size=(n, 3)))
    ppass = np.where(rng.random(n) < 0.3, ents["number"][tgt], sv_scene.ENTITYNUM_NONE).astype(np.int32)
    contents = ref.sv_point_contents(pts, ppass)
    out = os.path.join(HERE, "sv_world.npz")
    np.savez_compressed(out, starts=s, ends=e, mins=mins, maxs=maxs, passent=passent, masks=masks,
                        traces=np.ascontiguousarray(traces).view(np.uint8).reshape(-1, 56),
                        points=pts, point_pass=ppass, contents=contents, bsp_len=np.int64(len(m.data)))
    ent_hits = (traces["entityNum"] < sv_scene.ENTITYNUM_WORLD).mean()
    print(f"{out}: {len(traces)} traces ({ent_hits:.2f} hit entities), {n} point contents, {os.path.getsize(out) >> 10} KiB")


if __name__ == "__main__":
    main()
