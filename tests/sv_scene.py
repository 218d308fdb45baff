"""A deterministic server scene for the SV_Trace tests: entities to link (inline-model movers, boxes, missiles with
owners, triggers) and moves aimed at them.  Used by tests/test_sv_world.py and tests/golden/make_golden_sv.py."""
from __future__ import annotations

import ctypes as C

import numpy as np

from threecore_b200 import bspgen

ENTITYNUM_NONE, ENTITYNUM_WORLD, MAX_GENTITIES = 4095, 4094, 4096
V3 = C.c_float * 3


class SvEntity(C.Structure):       # sv_entity_t, include/sv_world_batch.h
    _fields_ = [("number", C.c_int), ("modelindex", C.c_int), ("linked", C.c_int), ("linkcount", C.c_int),
                ("bmodel", C.c_int), ("mins", V3), ("maxs", V3), ("contents", C.c_int), ("absmin", V3), ("absmax", V3),
                ("currentOrigin", V3), ("currentAngles", V3), ("ownerNum", C.c_int)]


def f3(a):
    return np.ascontiguousarray(a, np.float32)


def make_entities(m, n_ents=300, seed=71):
    """-> dict of arrays, one row per entity, in LINK order"""
    rng = np.random.default_rng(seed)
    nm = m.num_models
    E = n_ents
    number = rng.permutation(np.arange(1, 1500))[:E].astype(np.int32)
    bmodel = rng.random(E) < 0.4
    lo, hi = np.asarray(m.world_mins) + 200, np.asarray(m.world_maxs) - 200
    origin = f3(np.floor(rng.uniform(lo, hi, size=(E, 3))))
    # a cluster of entities around the world's centre planes, so that the upper sector nodes hold some
    mid = 0.5 * (np.asarray(m.world_mins) + np.asarray(m.world_maxs))
    k = E // 6
    origin[:k, 0] = np.floor(mid[0] + rng.uniform(-30, 30, size=k))
    origin[k:2 * k, 1] = np.floor(mid[1] + rng.uniform(-30, 30, size=k))
    angles = f3(np.where(rng.random((E, 1)) < 0.5, 0.0, rng.uniform(0, 360, size=(E, 3))))
    angles[rng.random(E) < 0.15, 0] = 0.0
    mins = f3(-np.floor(rng.uniform(8, 48, size=(E, 3))))
    maxs = f3(np.floor(rng.uniform(8, 64, size=(E, 3))))
    modelindex = rng.integers(1, nm, size=E).astype(np.int32)
    modelindex[~bmodel] = 0
    contents = np.where(bmodel, bspgen.CONTENTS_SOLID, bspgen.CONTENTS_BODY).astype(np.int32)
    trig = rng.random(E) < 0.1
    contents[trig] = bspgen.CONTENTS_TRIGGER                 # filtered out by MASK_PLAYERSOLID
    contents[rng.random(E) < 0.05] = bspgen.CONTENTS_PLAYERCLIP
    owner = np.full(E, ENTITYNUM_NONE, np.int32)
    # missiles: box entities owned by other entities (some owners are not linked at all)
    mis = np.flatnonzero(~bmodel)[: E // 8]
    owner[mis] = rng.choice(number, size=len(mis))
    owner[mis[:3]] = 2000                                     # an owner that is never linked
    return dict(number=number, bmodel=bmodel.astype(np.int32), modelindex=modelindex, mins=mins, maxs=maxs, origin=origin,
                angles=angles, contents=contents, owner=owner)


def make_moves(m, ents, n, seed=72):
    """moves aimed at entities: start/end, hull per move, pass entity per move, mask per move"""
    rng = np.random.default_rng(seed)
    E = len(ents["number"])
    tgt = rng.integers(0, E, size=n)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    reach = rng.uniform(40, 400, size=(n, 1))
    s = f3(ents["origin"][tgt] + rng.uniform(-40, 40, size=(n, 3)) - d * reach)
    e = f3(ents["origin"][tgt] + rng.uniform(-40, 40, size=(n, 3)) + d * reach * rng.uniform(-0.1, 1.0, size=(n, 1)))
    k = n // 40
    s[:k] = ents["origin"][tgt[:k]]                           # starts inside entities
    e[k:2 * k] = s[k:2 * k]                                   # position tests
    lw = rng.random(n) < 0.1                                  # long moves through many sectors
    e[lw] = f3(s[lw] + d[lw] * rng.uniform(1500, 5000, size=(int(lw.sum()), 1)))
    hull = rng.integers(0, 3, size=n)
    mins = f3(np.where(hull[:, None] == 0, 0.0, np.where(hull[:, None] == 1, bspgen.PLAYER_MINS, [-4.0, -4.0, -4.0])))
    maxs = f3(np.where(hull[:, None] == 0, 0.0, np.where(hull[:, None] == 1, bspgen.PLAYER_MAXS, [4.0, 4.0, 4.0])))
    passent = np.full(n, ENTITYNUM_NONE, np.int32)
    r = rng.random(n)
    passent[r < 0.25] = ents["number"][tgt[r < 0.25]]         # the target itself
    own = (r >= 0.25) & (r < 0.45)
    passent[own] = ents["owner"][tgt[own]]                    # the target's owner (own missiles are skipped)
    sib = (r >= 0.45) & (r < 0.55)
    passent[sib] = rng.choice(ents["number"], size=int(sib.sum()))
    passent[(r >= 0.55) & (r < 0.58)] = 2000
    masks = np.where(rng.random(n) < 0.7, bspgen.MASK_PLAYERSOLID, bspgen.MASK_SOLID).astype(np.int32)
    masks[rng.random(n) < 0.05] = -1
    return s, e, mins, maxs, passent, masks


def link_all_ref(ref, ents):
    ref.sv_clear_world()
    for i in range(len(ents["number"])):
        ref.sv_link_entity(ents["number"][i], ents["bmodel"][i], ents["modelindex"][i], ents["mins"][i], ents["maxs"][i],
                           ents["origin"][i], ents["angles"][i], ents["contents"][i], ents["owner"][i])
    ref.sv_set_owner(2000, int(ents["number"][0]))            # an unlinked gentity that has an owner


def new_game_array():
    arr = (SvEntity * MAX_GENTITIES)()
    for i in range(MAX_GENTITIES):
        arr[i].number = i
        arr[i].ownerNum = ENTITYNUM_NONE
    return arr


def fill_game_entity(g, ents, i):
    g.bmodel = int(ents["bmodel"][i]); g.modelindex = int(ents["modelindex"][i])
    g.mins = V3(*ents["mins"][i]); g.maxs = V3(*ents["maxs"][i])
    g.currentOrigin = V3(*ents["origin"][i]); g.currentAngles = V3(*ents["angles"][i])
    g.contents = int(ents["contents"][i]); g.ownerNum = int(ents["owner"][i])


def declare(lib):
    vp = C.c_void_p
    lib.SV_LocateGameData.argtypes = [vp, C.c_int, C.c_int]
    lib.SV_LocateGameData.restype = None
    lib.SV_ClearWorld.restype = None
    lib.SV_ClearWorldForBounds.argtypes = [vp, vp]
    lib.SV_ClearWorldForBounds.restype = None
    lib.SV_LinkEntity.argtypes = [vp]
    lib.SV_LinkEntity.restype = None
    lib.SV_UnlinkEntity.argtypes = [vp]
    lib.SV_UnlinkEntity.restype = None
    lib.SV_AreaEntities.argtypes = [vp, vp, vp, C.c_int]
    lib.SV_TraceBatch.argtypes = [vp, C.c_int, vp, vp, vp, C.c_int, vp, vp, C.c_int, vp, C.c_int]
    lib.SV_PointContentsWorldBatch.argtypes = [vp, C.c_int, vp, vp, C.c_int]
    lib.SV_LastBatchCandidates.restype = C.c_ulonglong
    lib.CM_LastError.restype = C.c_char_p
    return lib


def link_all_product(lib, arr, ents):
    for i in range(len(ents["number"])):
        g = arr[int(ents["number"][i])]
        fill_game_entity(g, ents, i)
        lib.SV_LinkEntity(C.byref(g))
    arr[2000].ownerNum = int(ents["number"][0])
