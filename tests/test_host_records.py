"""The host half of a big host batch's result path (cm_host_records.cpp, cmgpu_format_records; traceHostHybrid in cm_gpu.cu).

CPU: the formatter turns the 16 bytes a walk decides (fraction, solid flags, clip side, hit brush) back into the reference's
56-byte trace_t -- endpos (cm_trace.c:671-686), the clip side's plane copy (cm_trace.c:352-356), surface flags, contents --
for traces made by the unmodified reference (the restatement where the reference tree is absent).
GPU: however a batch's records are shared between the link and the host's threads, they are the bytes the device writes."""
import numpy as np
import pytest

from helpers import cmoracle, cmref, get_map, have_ref
from threecore_b200 import bspgen, engine


def side_tables(flat):
    pl = flat["brushsides"][:, 0]
    planes = flat["planes"][pl]
    ts = flat["plane_type"][pl].astype(np.int32) | (flat["plane_signbits"][pl].astype(np.int32) << 8)
    return np.ascontiguousarray(planes), ts, np.ascontiguousarray(flat["brushsides"][:, 1]), np.ascontiguousarray(flat["brushes"][:, 0])


def results16_from_records(rec, planes, ts, sflags, contents):
    """what the walk would have left for these records: any side with the record's plane bytes, any brush with its contents"""
    side_key = {}
    for i in range(len(ts) - 1, -1, -1):
        side_key[(planes[i].tobytes(), int(ts[i]) & 0xffff, int(sflags[i]))] = i
    brush_key = {}
    for i in range(len(contents) - 1, -1, -1):
        brush_key[int(contents[i])] = i
    raw = np.ascontiguousarray(rec).view(np.uint8).reshape(-1, 56)
    words = raw.view(np.uint32).reshape(-1, 14)
    out = np.zeros((len(rec), 4), dtype=np.uint32)
    out[:, 0] = words[:, 2]
    out[:, 1] = words[:, 0] | (words[:, 1] << 1)
    for i in range(len(rec)):
        key = (raw[i, 24:40].tobytes(), int(words[i, 10]) & 0xffff, int(words[i, 11]))
        if key == (b"\0" * 16, 0, 0):
            out[i, 2] = 0xffffffff
        else:
            out[i, 2] = side_key[key]
        out[i, 3] = brush_key[int(words[i, 12])] if words[i, 12] else 0xffffffff
    return out


@pytest.mark.parametrize("kind", ["tiny", "arena_small"])
def test_formatter_rebuilds_the_reference_records(kind):
    m = get_map(kind)
    flat = engine.FlatMap.from_bsp(m.data).to_dict()
    n = 6000
    starts, ends = bspgen.make_sweeps(m, n, seed=41)
    starts[:200] = ends[:200]                                   # position tests: startsolid / allsolid records
    chk = cmref.RefCM().load(m.data) if have_ref() else cmoracle.OracleCM(flat)
    want = chk.box_trace(starts, ends, bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, mask=bspgen.MASK_PLAYERSOLID)
    assert (want["fraction"] < 1).any() and (want["fraction"] == 1).any() and want["startsolid"].any()
    planes, ts, sflags, contents = side_tables(flat)
    r16 = results16_from_records(want, planes, ts, sflags, contents)
    got, hit = engine.format_records(r16, starts, ends, planes, ts, sflags, contents, want_hit=True)
    a = np.ascontiguousarray(got).view(np.uint8).reshape(-1, 56)
    b = np.ascontiguousarray(want).view(np.uint8).reshape(-1, 56)
    bad = np.flatnonzero((a != b).any(axis=1))
    assert bad.size == 0, (bad[:5], got[bad[:2]], want[bad[:2]])
    assert np.array_equal(hit, r16[:, 3].view(np.int32))
    # an unaligned record array takes the plain-store path: same bytes
    buf = np.zeros(56 * n + 4, dtype=np.uint8)
    lib = engine.load_library()
    f32 = lambda x: np.ascontiguousarray(x, dtype=np.float32)
    i32 = lambda x: np.ascontiguousarray(x, dtype=np.int32)
    s_, e_, p_, t_, f_, c_ = f32(starts), f32(ends), f32(planes), i32(ts), i32(sflags), i32(contents)
    rc = lib.cmgpu_format_records(n, r16.ctypes.data, s_.ctypes.data, e_.ctypes.data, len(t_), p_.ctypes.data, t_.ctypes.data, f_.ctypes.data,
                                  len(c_), c_.ctypes.data, buf.ctypes.data + 4, None)
    assert rc == 0 and np.array_equal(buf[4:].reshape(-1, 56), b)


def test_worker_pool_stages_and_formats_like_the_single_thread():
    """the pool of a big host batch without a device: staged sweeps, held jobs released out of order, 1 to 7 threads"""
    m = get_map("arena_small")
    flat = engine.FlatMap.from_bsp(m.data).to_dict()
    n = 230001
    starts, ends = bspgen.make_sweeps(m, n, seed=47)
    chk = cmoracle.OracleCM(flat)
    want = chk.box_trace(starts[:40000], ends[:40000], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, mask=bspgen.MASK_PLAYERSOLID)
    planes, ts, sflags, contents = side_tables(flat)
    r16 = np.tile(results16_from_records(want, planes, ts, sflags, contents), (6, 1))[:n]
    one = engine.format_records(r16, starts, ends, planes, ts, sflags, contents)
    assert np.array_equal(np.ascontiguousarray(one[:40000]).view(np.uint8), np.ascontiguousarray(want).view(np.uint8))
    for threads in (1, 3, 7):
        got, hit = engine.format_records(r16, starts, ends, planes, ts, sflags, contents, want_hit=True, threads=threads)
        assert np.array_equal(np.ascontiguousarray(got).view(np.uint8), np.ascontiguousarray(one).view(np.uint8)), threads
        assert np.array_equal(hit, r16[:, 3].view(np.int32))


def test_formatter_refuses_indices_outside_the_tables():
    r16 = np.array([[0x3f800000, 0, 5, 0xffffffff]], dtype=np.uint32)
    z = np.zeros((1, 3), np.float32)
    with pytest.raises(engine.CMGPUError):
        engine.format_records(r16, z, z, np.zeros((2, 4), np.float32), np.zeros(2, np.int32), np.zeros(2, np.int32), np.zeros(1, np.int32))


@pytest.mark.gpu
def test_shared_records_equal_the_device_records():
    """4 Mi + 1 sweeps through cmgpu_box_trace_batch with the records written by the device only (round-1 pipeline), by the
    host threads only, by the new pipeline's device path only, and shared by the measured rates -- pinned and pageable result
    arrays, with hit ids: always the bytes of the device-pointer call."""
    import torch
    m = bspgen.arena_map(n_brushes=6000, seed=0xC0FFEE31)
    eng = engine.Engine(0)
    eng.load_bsp(m.data)
    n = (1 << 22) + 1
    starts, ends = bspgen.make_sweeps(m, n, seed=43, z_range=(24.0, 2048.0))
    starts[::1000] = ends[::1000]
    dev = torch.device("cuda", 0)
    d_out = torch.empty((n, 56), dtype=torch.uint8, device=dev)
    eng.box_trace_device(torch.from_numpy(starts).to(dev), torch.from_numpy(ends).to(dev), bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS,
                         bspgen.MASK_PLAYERSOLID, d_out)
    torch.cuda.synchronize()
    want = d_out.cpu().numpy()
    assert (want[:, 8:12].view(np.float32) < 1).any()
    h_s, h_e = torch.from_numpy(starts).pin_memory(), torch.from_numpy(ends).pin_memory()
    h_out = torch.empty((n, 56), dtype=torch.uint8).pin_memory()
    for mode, threads in ((0, -1), (1, 3), (2, -1), (-1, -1), (-1, 2)):
        eng.set_option("host_records", mode)
        eng.set_option("host_threads", threads)
        for rep in range(5 if mode < 0 else 1):                 # the shared pipeline tries both ways before it settles
            h_out.zero_()
            eng.box_trace_hostptr(n, h_s.data_ptr(), h_e.data_ptr(), bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID,
                                  h_out.data_ptr())
            assert np.array_equal(h_out.numpy(), want), (mode, threads, rep)
        by_host, by_device = eng.last_batch_split()
        if mode == 1:
            assert (by_host, by_device) == (n, 0)
        if mode == 2:
            assert (by_host, by_device) == (0, n)
    # pageable arrays and hit ids through the numpy front end
    eng.set_option("host_records", -1)
    got, hit = eng.box_trace(starts, ends, bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, mask=bspgen.MASK_PLAYERSOLID, want_hit=True)
    assert np.array_equal(np.ascontiguousarray(got).view(np.uint8).reshape(-1, 56), want)
    assert eng.last_batch_split() == (n, 0)                     # a pageable result array is written by the threads
    eng.set_option("host_records", 0)
    got0, hit0 = eng.box_trace(starts, ends, bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, mask=bspgen.MASK_PLAYERSOLID, want_hit=True)
    assert np.array_equal(hit, hit0) and (hit >= 0).any()
    eng.close()
