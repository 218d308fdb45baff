#!/usr/bin/env python
"""bench.py -- Mtraces/s of batched CM_BoxTrace (BASELINE.json metric) on N B200s of one node.

  python bench.py --gpus N --steps K --warmup W            our CUDA engine
  python bench.py --impl reference --gpus N --steps K ...  the reference's own CPU code (oracle/_ref)

Workload (BASELINE.json configs[1]): 16 Mi player-hull sweeps (mins -15,-15,-24 /
maxs 15,15,32, MASK_PLAYERSOLID) through a synthetic 20k-brush arena BSP.  One
"step" = one pass of the hot path over the whole ray batch.  With N > 1 the map
is replicated and every rank traces its own 16 Mi-ray shard (weak scaling, no
data-path collective).

Numbers on the JSON line:
  value      Mtraces/s, whole job, rays resident in HBM, CUDA-event timed, max over ranks
  e2e        the same metric through the host-pointer C ABI call (cmgpu_box_trace_batch) with
             pinned host buffers: H2D of the rays and D2H of the 56-byte trace_t records inside
             the timed region
  roofline   algorithmic HBM bytes (80 B per trace) / kernel time vs the measured copy peak
  cpu_baseline  the reference's CM_BoxTrace on this box's host cores (bounded sample), N=1 rank 0 only
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from threecore_b200 import bspgen  # noqa: E402

N_RAYS = 1 << 24
N_BRUSHES = 20000
MAP_SEED = 0xC0FFEE02
RAY_SEED = 0xBADC0DE2
BYTES_PER_TRACE = 24 + 56          # ray record (start,end) + trace_t, shared hull and mask
METRIC = "Mtraces/s batched CM_BoxTrace"


def workload_config(n_rays, gpus):
    return {"workload": "configs[1]: 16Mi player-hull box traces (mins -15,-15,-24 maxs 15,15,32, MASK_PLAYERSOLID) "
                        "on a synthetic 20k-brush arena BSP",
            "rays_per_gpu": n_rays, "brushes": N_BRUSHES, "map_seed": hex(MAP_SEED), "ray_seed": hex(RAY_SEED),
            "parallelism": f"BSP replicated per GPU, rays sharded x{gpus} (the reference arm runs the same rays on the host cores)",
            "l2_policy": "inputs (384 MiB of rays + 896 MiB of results per step) are larger than the 126 MB L2; no flush needed"}


def make_workload(n_rays, rank=0):
    m = bspgen.arena_map(n_brushes=N_BRUSHES, seed=MAP_SEED)
    starts, ends = bspgen.make_sweeps(m, n_rays, RAY_SEED + 1000 * rank, z_range=(24.0, 2048.0))
    return m, starts, ends


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def parity_gate(eng, data, starts, ends, d_out, sample, seed):
    """no timing is accepted before the records agree with the reference.  A random sample of this rank's
    batch, traced by the unmodified reference (oracle/_ref, forked workers) or, when that library was not shipped, by
    the oracle port; all 56 bytes of every sampled record must be equal."""
    import torch
    from oracle import cmref
    n = len(starts)
    idx = np.sort(np.random.default_rng(seed).choice(n, size=min(sample, n), replace=False))
    got = d_out[torch.from_numpy(idx).to(d_out.device)].cpu().numpy()
    s, e = np.ascontiguousarray(starts[idx]), np.ascontiguousarray(ends[idx])
    if cmref.available():
        _, _, want = cmref.RefCM().load(data).bench_box_trace(s, e, bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID,
                                                              nworkers=host_cores(), want_results=True)
        against = "reference (oracle/_ref/libcmref.so)"
    else:
        from oracle import cmoracle
        from threecore_b200.engine import FlatMap
        want = cmoracle.OracleCM(FlatMap.from_bsp(data).to_dict()).box_trace(s, e, bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS,
                                                                             mask=bspgen.MASK_PLAYERSOLID)
        against = "oracle port (oracle/cm_oracle.c)"
    w = np.ascontiguousarray(want).view(np.uint8).reshape(-1, 56)
    bad = int((got != w).any(axis=1).sum())
    return {"checked": int(len(idx)), "mismatches": bad, "against": against,
            "what": "all 56 bytes of trace_t, random sample of the timed batch, before timing"}


def ncu_traffic():
    """Per-launch DRAM bytes of the trace kernel from the committed ncu capture, if one exists."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        with open(p) as f:
            return json.load(f)
    except Exception:
        return None


class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.QUERY}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_reference_run(data, starts, ends, sample, repeats=1):
    """Time the reference's own CM_BoxTrace (oracle/_ref) with one forked worker per host core."""
    from oracle import cmref
    cores = host_cores()
    if cmref.available():
        r = cmref.RefCM().load(data)
        secs, _, _ = r.bench_box_trace(starts[:sample], ends[:sample], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS,
                                       bspgen.MASK_PLAYERSOLID, nworkers=cores, repeats=repeats)
        return sample / secs / 1e6, cores, "reference", secs
    # the reference library was not shipped: time the C restatement instead (single process)
    from oracle import cmoracle
    from threecore_b200.engine import FlatMap
    o = cmoracle.OracleCM(FlatMap.from_bsp(data).to_dict())
    t0 = time.perf_counter()
    o.box_trace(starts[:sample], ends[:sample], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, mask=bspgen.MASK_PLAYERSOLID)
    secs = time.perf_counter() - t0
    return sample / secs / 1e6, 1, "port", secs


def single_core_run(data, starts, ends, sample):
    """the same reference code on ONE core (SURVEY 8d asks for both figures)"""
    from oracle import cmref
    if not cmref.available():
        return None
    r = cmref.RefCM().load(data)
    secs, _, _ = r.bench_box_trace(starts[:sample], ends[:sample], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS,
                                   bspgen.MASK_PLAYERSOLID, nworkers=1, repeats=1)
    return sample / secs / 1e6


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = args.ref_sample
    m, starts, ends = make_workload(max(sample, 1 << 16))
    vals = []
    for _ in range(args.warmup):
        cpu_reference_run(m.data, starts, ends, min(sample, 1 << 18))
    cores = kind = None
    t_all = 0.0
    for _ in range(args.steps):
        v, cores, kind, secs = cpu_reference_run(m.data, starts, ends, sample)
        vals.append(v)
        t_all += secs
    value = float(sample * args.steps / t_all / 1e6)
    sample_txt = f"{sample} rays per step drawn from the configs[1] distribution (same map, same generator), one forked worker per host core"
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Mtraces/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_all / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
            "config": workload_config(N_RAYS, args.gpus),
            "cpu_baseline": {"value": value, "unit": "Mtraces/s", "cores": cores, "kind": kind, "sample": sample_txt},
            "e2e": {"value": value, "unit": "Mtraces/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---- the other arms of the line (the multi-GPU data path, the host copy ceiling, pageable callers) ----
def _max_over_ranks(x, dev, world):
    import torch
    import torch.distributed as dist
    t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _barrier(world):
    import torch
    import torch.distributed as dist
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def arm_host_copy_ceiling(dev, world, n, h_starts, h_ends, h_out, d_starts, d_ends, d_out, reps=3):
    """What the box's PCIe path gives this job with NO kernel in the way: every rank copies one step's rays host->device
    and one step's records device->host at the same time (pinned memory, two streams), all ranks together.  The end to
    end arm cannot beat n / that time."""
    import torch
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def once():
        with torch.cuda.stream(s_in):
            d_starts.copy_(h_starts, non_blocking=True)
            d_ends.copy_(h_ends, non_blocking=True)
        with torch.cuda.stream(s_out):
            h_out.copy_(d_out, non_blocking=True)
    once()
    _barrier(world)
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    torch.cuda.synchronize()
    secs = _max_over_ranks(time.perf_counter() - t0, dev, world) / reps
    return {"secs_per_step": secs, "gbs": world * 80.0 * n / secs / 1e9, "copy_bound_mtraces": world * n / secs / 1e6,
            "what": f"{world} rank(s) at once, each: 24 B/trace host->device + 56 B/trace device->host, pinned, concurrent, no kernels"}


def arm_e2e_pageable(eng, dev, world, n, starts, ends, steps=2):
    """The host-pointer call on plain malloc'ed arrays (what a caller that never heard of CUDA passes): as they are (the
    context's worker threads copy the rays into pinned memory piece by piece and write every record themselves; round 1:
    the driver staged every copy and nothing overlapped), and after cmgpu_host_register of the same three arrays."""
    out = np.empty((n, 56), np.uint8)

    def step():
        eng.box_trace_hostptr(n, starts.ctypes.data, ends.ctypes.data, bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS,
                              bspgen.MASK_PLAYERSOLID, out.ctypes.data)

    def timed(k):
        step()
        _barrier(world)
        t0 = time.perf_counter()
        for _ in range(k):
            step()
        return world * n * k / _max_over_ranks(time.perf_counter() - t0, dev, world) / 1e6
    res = {"unit": "Mtraces/s", "unregistered": timed(1)}
    t0 = time.perf_counter()
    for a in (starts, ends, out):
        eng.host_register(a.ctypes.data, a.nbytes)
    res["register_ms_once"] = 1e3 * (time.perf_counter() - t0)
    res["registered"] = timed(steps)
    for a in (starts, ends, out):
        eng.host_unregister(a.ctypes.data)
    res["what"] = ("cmgpu_box_trace_batch on numpy (pageable) arrays: `unregistered` = as they are (rays staged and records written by the context's threads), `registered` = after "
                   "cmgpu_host_register of rays and results (one-time cost register_ms_once)")
    return res, out


def arm_gathered(eng, dev, world, rank, n, d_starts, d_ends, d_out, steps):
    """north_star's data path: every rank traces its shard and its kernels store the records straight into rank 0's HBM
    over NVLink (sharding.PeerResults); the timed region ends when all records have landed."""
    import torch
    import torch.distributed as dist
    from threecore_b200.sharding import PeerResults
    pr = PeerResults(eng, n * world)
    lo = rank * n

    def step():
        eng.box_trace_device(d_starts, d_ends, bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, pr.pointer(lo))
    for _ in range(2):
        step()
    _barrier(world)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        step()
    b.record()
    _barrier(world)
    ms = _max_over_ranks(a.elapsed_time(b), dev, world) / steps
    eng.check_status()
    # every rank's shard in rank 0's buffer must be the records that rank holds locally (64-bit word sums)
    mine = d_out.view(torch.int64).sum().reshape(1)
    sums = [torch.zeros_like(mine) for _ in range(world)]
    if world > 1:
        dist.all_gather(sums, mine)
    else:
        sums = [mine]
    ok = None
    if rank == 0:
        rec = pr.records()
        ok = all(int(rec[r * n:(r + 1) * n].view(torch.int64).sum().item()) == int(sums[r].item()) for r in range(world))
    pr.close()
    return {"value": world * n / ms / 1e3, "unit": "Mtraces/s", "ms_per_step": ms, "verified": ok,
            "rank0_ingress_gbs": (world - 1) * n * 56 / ms / 1e6,
            "what": f"{world} x {n} rays, all {world * n} trace_t records landing in rank 0's HBM inside the timed region "
                    "(peer stores of the trace kernels over NVLink, no gather pass)"}


def arm_strong(eng, dev, world, rank, n, starts, ends, d_out_rank0_full, steps):
    """Strong scaling: BASELINE's 16 Mi batch (rank 0's stream) split into `world` contiguous ranges, records kept on the
    ranks (`value`) and gathered into rank 0's HBM (`gathered`)."""
    import torch
    from threecore_b200.sharding import PeerResults, shard_range
    if rank != 0:
        _, starts, ends = make_workload(n, 0)          # the one global batch; rank 0 already holds it
    lo, hi = shard_range(n, rank, world)
    ds, de = torch.from_numpy(starts[lo:hi]).to(dev), torch.from_numpy(ends[lo:hi]).to(dev)
    local = torch.empty((hi - lo, 56), dtype=torch.uint8, device=dev)
    pr = PeerResults(eng, n)

    def timed(target):
        for _ in range(2):
            eng.box_trace_device(ds, de, bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, target)
        _barrier(world)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(steps):
            eng.box_trace_device(ds, de, bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, target)
        b.record()
        _barrier(world)
        return _max_over_ranks(a.elapsed_time(b), dev, world) / steps
    ms_local = timed(local)
    ms_gath = timed(pr.pointer(lo))
    eng.check_status()
    ok = bool(torch.equal(pr.records(), d_out_rank0_full)) if rank == 0 else None
    pr.close()
    return {"value": n / ms_local / 1e3, "unit": "Mtraces/s", "ms_per_step": ms_local, "rays_total": n, "rays_per_gpu": hi - lo,
            "gathered": {"value": n / ms_gath / 1e3, "ms_per_step": ms_gath, "equals_single_gpu_records": ok,
                         "rank0_ingress_gbs": (n - (n // world)) * 56 / ms_gath / 1e6},
            "limiter": "a launch of n/N rays still pays the five binning launches, the expansion launch and the walk of its slowest warp "
                       "(the per-launch floor is ~0.3 ms; see DESIGN.md 5)"}


def arm_rollout(local, dev, world, rank, ticks=40):
    """BASELINE config 5: 65 536 Pmove-shaped agents x 10 traces per tick on the 50k-brush arena, agents sharded by id,
    the tick replayed as a CUDA graph (threecore_b200/rollout.py)."""
    import torch
    from threecore_b200.engine import Engine
    from threecore_b200.rollout import Rollout, TRACES_PER_TICK
    from threecore_b200.sharding import shard_range
    agents = 65536
    m = bspgen.arena_map(n_brushes=50000, seed=0xC0FFEE05)
    e2 = Engine(local)
    try:
        e2.load_bsp(m.data)
        lo, hi = shard_range(agents, rank, world)
        ro = Rollout(e2, hi - lo, m.world_mins, m.world_maxs, seed=9, first_agent=lo).capture(warmup=3)
        ro.run(3)
        _barrier(world)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ro.run(ticks)
        b.record()
        _barrier(world)
        ms = _max_over_ranks(a.elapsed_time(b), dev, world) / ticks
        e2.check_status()
    finally:
        e2.close()
    return {"value": agents * TRACES_PER_TICK / ms / 1e3, "unit": "Mtraces/s", "ms_per_tick": ms, "agents": agents,
            "agents_per_gpu": hi - lo, "brushes": 50000, "scaling": "strong (agents sharded by id)",
            "limiter": "a tick is eight DEPENDENT trace launches; each waits for its slowest walk"}


def arm_other_configs(local, dev):
    """The other BASELINE configs and the botlib tracer on the driver's record (rank 0, one GPU): resident throughput, each
    with a sample of its records compared with the reference (oracle/_ref) when that library was shipped."""
    import torch
    from oracle import cmref
    from threecore_b200 import aasgen
    from threecore_b200.engine import Engine, rotation_matrices
    have = cmref.available()
    e2 = Engine(local)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    res = {}

    def timed(fn, n, reps=5):
        for _ in range(2):
            fn()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        e2.check_status()
        ms = a.elapsed_time(b) / reps
        return {"value": n / ms / 1e3, "unit": "M items/s", "ms": ms, "items": n}

    def sample_equal(got_u8, want):
        return bool(np.array_equal(got_u8, np.ascontiguousarray(want).view(np.uint8).reshape(got_u8.shape)))
    try:
        idx = np.arange(0, 1 << 20, 16)
        # config 1: 1 Mi point traces, 500-brush room
        m = bspgen.room_map(n_brushes=500, seed=0xC0FFEE01)
        e2.load_bsp(m.data)
        n = 1 << 20
        s, e = bspgen.make_point_sweeps(m, n, 0xBADC0DE1)
        ds, de, out = T(s), T(e), torch.empty((n, 56), dtype=torch.uint8, device=dev)
        res["config1_point_traces"] = timed(lambda: e2.box_trace_device(ds, de, None, None, bspgen.MASK_SOLID, out), n)
        if have:
            res["config1_point_traces"]["sample_equals_reference"] = sample_equal(
                out.cpu().numpy()[idx], cmref.RefCM().load(m.data).box_trace(s[idx], e[idx], None, None, mask=bspgen.MASK_SOLID))
        # config 3: 4 Mi point + hull traces, 2k brushes + 512 Bezier patches
        m = bspgen.patch_map(n_brushes=2000, n_patches=512, seed=0xC0FFEE03)
        e2.load_bsp(m.data)
        n = 1 << 22
        s, e, mins, maxs = bspgen.make_patch_sweeps(m, n, 0xBADC0DE3)
        ds, de, dmn, dmx, out = T(s), T(e), T(mins), T(maxs), torch.empty((n, 56), dtype=torch.uint8, device=dev)
        res["config3_patches"] = timed(lambda: e2.box_trace_device(ds, de, dmn, dmx, bspgen.MASK_ALL, out), n)
        if have:
            res["config3_patches"]["sample_equals_reference"] = sample_equal(
                out.cpu().numpy()[idx], cmref.RefCM().load(m.data).box_trace(s[idx], e[idx], mins[idx], maxs[idx], mask=bspgen.MASK_ALL))
        # config 4: transformed traces against 256 inline models, point contents
        m = bspgen.arena_map(n_brushes=N_BRUSHES, seed=MAP_SEED, n_models=256)
        e2.load_bsp(m.data)
        t = bspgen.make_entity_traces(m, n, 0xBADC0DE4)
        rot, ridx = rotation_matrices(t["angles"])
        d = {k: T(v) for k, v in t.items() if k != "angles"}
        drot, dridx = T(rot), T(ridx)
        res["config4_transformed_traces"] = timed(lambda: e2.box_trace_device(
            d["starts"], d["ends"], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_ALL, out, models=d["models"], origins=d["origins"],
            rotations=drot, rot_index=dridx, boxmins=d["boxmins"], boxmaxs=d["boxmaxs"]), n)
        if have:
            r = cmref.RefCM().load(m.data)
            res["config4_transformed_traces"]["sample_equals_reference"] = sample_equal(out.cpu().numpy()[idx], r.transformed_box_trace(
                t["starts"][idx], t["ends"][idx], bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, t["models"][idx], bspgen.MASK_ALL, t["origins"][idx],
                t["angles"][idx], t["boxmins"][idx], t["boxmaxs"][idx]))
        npts = 1 << 24
        pts = bspgen.make_points(m, npts, 0xBADC0DE5)
        dp, pc = T(pts), torch.empty(npts, dtype=torch.int32, device=dev)
        res["config4_point_contents"] = timed(lambda: e2.point_contents_device(dp, pc), npts)
        if have:
            res["config4_point_contents"]["sample_equals_reference"] = bool(np.array_equal(pc.cpu().numpy()[idx], r.point_contents(pts[idx])))
        del ds, de, dmn, dmx, out, dp, pc, d
        # botlib: AAS_TraceClientBBox and AAS_PredictClientMovement on a generated AAS world
        w = aasgen.make_world(n_boxes=2000, seed=0xAA5004, extent=8192.0, height=1024.0)
        e2.aas_upload(w)
        n = 1 << 24
        s, e = aasgen.make_lines(w, n, 11)
        ds, de, out = T(s), T(e), torch.empty((n, 36), dtype=torch.uint8, device=dev)
        res["aas_trace_client_bbox"] = timed(lambda: e2.aas_trace_client_bbox_device(ds, de, aasgen.PRESENCE_NORMAL, out), n)
        nm = 1 << 20
        p = aasgen.make_moves(w, nm, 13)
        dpm = {k: T(v) for k, v in p.items()}
        mo = torch.empty((nm, 84), dtype=torch.uint8, device=dev)
        res["aas_predict_client_movement"] = timed(lambda: e2.aas_predict_client_movement_device(dpm, mo), nm, reps=3)
        if have:
            r.aas_load(w)
            res["aas_trace_client_bbox"]["sample_equals_reference"] = sample_equal(
                out.cpu().numpy()[idx], r.aas_trace_client_bbox(s[idx], e[idx], aasgen.PRESENCE_NORMAL))
            k = idx[:16384]
            res["aas_predict_client_movement"]["sample_equals_reference"] = sample_equal(
                mo.cpu().numpy()[k], r.aas_predict_client_movement({kk: v[k] for kk, v in p.items()})[0])
    finally:
        e2.close()
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist

    from threecore_b200.engine import Engine

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    # stdout carries exactly one JSON line: whatever native libraries print while the job runs (NCCL's version banner
    # for one) goes to stderr
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n = args.rays
    m, starts, ends = make_workload(n, rank)
    eng = Engine(local)
    eng.load_bsp(m.data)

    # ---- device-resident arm -------------------------------------------------------------
    d_starts = torch.from_numpy(starts).to(dev)
    d_ends = torch.from_numpy(ends).to(dev)
    d_out = torch.empty((n, 56), dtype=torch.uint8, device=dev)

    def step():
        eng.box_trace_device(d_starts, d_ends, bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, d_out)

    for _ in range(args.warmup):
        step()
    eng.check_status()
    torch.cuda.synchronize()
    # parity gate before any timing: rank 0 checks a large sample, the other ranks (same host cores) a small one
    parity = parity_gate(eng, m.data, starts, ends, d_out, args.parity_sample if rank == 0 else 1 << 14, 77 + rank)
    if parity["mismatches"]:
        raise SystemExit(f"bench.py: parity gate failed on rank {rank}: {parity}")
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = eng.launch_count
    eng.kernel_timing(True)                     # CUDA events around the trace kernel itself, on its launching stream
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    torch.cuda.synchronize()
    evs[0].record()
    for k in range(args.steps):
        step()
        evs[k + 1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    launches = eng.launch_count - launches0
    trace_ms, trace_launches = eng.kernel_time_ms()
    eng.kernel_timing(False)
    total_ms = evs[0].elapsed_time(evs[-1])
    kern_ms = [evs[k].elapsed_time(evs[k + 1]) for k in range(args.steps)]
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    eng.check_status()
    value = world * n * args.steps / (total_ms * 1e-3) / 1e6
    hits = float((d_out[:, 8:12].view(torch.float32) < 1.0).float().mean().item())

    # ---- end-to-end arm: host pointers through the C ABI ------------------------------------
    h_starts = torch.from_numpy(starts).pin_memory()
    h_ends = torch.from_numpy(ends).pin_memory()
    h_out = torch.empty((n, 56), dtype=torch.uint8).pin_memory()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))

    def e2e_step():
        eng.box_trace_hostptr(n, h_starts.data_ptr(), h_ends.data_ptr(), bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS,
                              bspgen.MASK_PLAYERSOLID, h_out.data_ptr())

    def e2e_timed(warm):
        for _ in range(warm):
            e2e_step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        tt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return world * n * e2e_steps / float(tt.item()) / 1e6

    # The call as shipped: a big batch's records are shared between the link and the host's threads (cm_gpu.h,
    # "host_records" -1).  The call tries both of its pipelines before it settles (four calls and two, the first of each
    # uncounted), so it gets eight untimed calls here; the plain pipeline alone (every record over PCIe) is timed next to it.
    e2e_value = e2e_timed(max(8, min(args.warmup, 10)))
    split = eng.last_batch_split()
    e2e_split = {"records_by_host_threads": split[0], "records_by_device": split[1]}
    same_shared = bool(torch.equal(h_out[: 1 << 16], d_out[: 1 << 16].cpu()))
    eng.set_option("host_records", 0)
    h_out.zero_()
    e2e_plain = e2e_timed(2)
    eng.set_option("host_records", -1)
    clocks = sampler.stop() if rank == 0 else None
    # the e2e results must be the same records the resident arm produced
    same = same_shared and bool(torch.equal(h_out[: 1 << 16], d_out[: 1 << 16].cpu()))
    extra = {}
    if not args.no_extra_arms:
        def attempt(name, fn):
            try:
                extra[name] = fn()
            except Exception as ex:                       # an extra arm never takes the headline down with it
                extra[name] = {"error": f"{type(ex).__name__}: {ex}"[:300]}
                torch.cuda.synchronize()
        attempt("host_copy", lambda: arm_host_copy_ceiling(dev, world, n, h_starts, h_ends, h_out, d_starts, d_ends, d_out))
    if not args.no_extra_arms:
        def pageable():
            res, out = arm_e2e_pageable(eng, dev, world, n, starts, ends)
            res["matches_resident"] = bool(np.array_equal(out[: 1 << 16], d_out[: 1 << 16].cpu().numpy()))
            return res
        attempt("e2e_pageable", pageable)
        del h_starts, h_ends, h_out
        if world > 1:
            attempt("gathered", lambda: arm_gathered(eng, dev, world, rank, n, d_starts, d_ends, d_out, max(2, min(args.steps, 5))))
        attempt("strong", lambda: arm_strong(eng, dev, world, rank, n, starts, ends, d_out, max(2, min(args.steps, 5))))
        attempt("rollout", lambda: arm_rollout(local, dev, world, rank))
        if world == 1:
            attempt("other_configs", lambda: arm_other_configs(local, dev))

    # ---- L2 working-set roofline: visit counts of this very batch from the counting kernel variant (untimed),
    # priced with the reference's own record sizes, against an L2-resident copy measured here
    l2 = None
    if rank == 0:
        eng.enable_counters(True)
        eng.read_counters(reset=True)
        step()
        torch.cuda.synchronize()
        c = eng.read_counters(reset=True)
        eng.enable_counters(False)
        l2_bytes = (36 * c["nodes"] + 24 * c["leafs"] + 4 * c["brushRefs"] + 56 * c["brushBounds"] + 36 * c["brushSides"]
                    + 20 * c["patchPlanes"] + 320 * c["facets"])
        nb = 24 << 20                                   # two 24 MiB buffers: resident in the 126 MB L2
        a = torch.empty(nb, dtype=torch.uint8, device=dev)
        b = torch.empty(nb, dtype=torch.uint8, device=dev)
        best = 1e9
        for _ in range(30):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            b.copy_(a)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        l2_copy = 2 * nb / (best * 1e-3) / 1e9
        del a, b
        l2_peak = eng.measure_l2_read(nb, 50)           # the walk reads: its ceiling is the rate at which SMs read L2
        l2 = {"bytes_per_trace": l2_bytes / max(c["traces"], 1), "peak": l2_peak, "unit": "GB/s",
              "peak_source": "measured here: cmgpu_measure_l2_read, 50 sweeps of 16-byte loads over a 24 MiB buffer that stays in L2, "
                             "best of 4 timed launches",
              "l2_copy_gbs": l2_copy,   # a torch copy of the same buffer, read+write bytes (the earlier denominator)
              "visits_per_trace": {k: c[k] / max(c["traces"], 1) for k in ("nodes", "leafs", "brushRefs", "brushBounds", "brushSides", "splits")}}

    if rank == 0:
        peak, peak_src = measured_peak()
        kms = trace_ms / max(trace_launches, 1)        # the trace launch alone (walk + record expansion), without the binning passes
        l2["achieved"] = l2["bytes_per_trace"] * n / (kms * 1e-3) / 1e9
        l2["frac"] = l2["achieved"] / l2["peak"]
        achieved = BYTES_PER_TRACE * n / (kms * 1e-3) / 1e9
        tr = ncu_traffic()
        roof = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": (tr or {}).get("dram_bytes_per_launch") if n == N_RAYS else None,
                "traffic_source": ("NOT measured in this run: dram__bytes_read.sum + dram__bytes_write.sum of the two kernels from the "
                                   f"committed ncu capture {tr.get('source')} ({tr.get('captured', 'date not recorded')}), same command at 16 Mi rays")
                                  if tr and n == N_RAYS else None,
                "warp_efficiency": ({"lanes_per_instruction": tr.get("lanes_per_instruction"), "of": 32,
                                     "issue_active_pct": tr.get("issue_active_pct"),
                                     "source": f"smsp__thread_inst_executed_per_inst_executed.ratio of the walk kernel, committed ncu --set full capture {tr.get('full_source')}"}
                                    if tr and tr.get("lanes_per_instruction") else None),
                "peak_source": peak_src,
                "kernel": "cmgpu::tracePoolKernel<false, 3> + cmgpu::expandRecordsKernel (the two phases of one trace launch: "
                          "the walk stores 16 B per item, the expansion writes the 56-byte records; kernel_ms and traffic cover both)",
                "kernel_ms": kms, "kernel_launches": trace_launches,
                "step_ms": float(np.mean(kern_ms)),
                "algorithmic_bytes_per_trace": BYTES_PER_TRACE, "traces_per_launch": n, "l2": l2,
                "note": "irregular BSP walk: bound by instruction issue and dependent-load latency, not by HBM; the HBM floor "
                        "is reported because the metric asks for it (DESIGN.md 4.1)"}
        line = {"metric": METRIC, "value": value, "unit": "Mtraces/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64", "data": "synthetic",
                "config": workload_config(n, world),
                "parity": parity,
                "roofline": roof,
                "e2e": {"value": e2e_value, "unit": "Mtraces/s", "h2d_bytes_per_step": int(24 * n),
                        "d2h_bytes_per_step": int((16 * e2e_split["records_by_host_threads"] + 56 * e2e_split["records_by_device"])
                                                  if sum(split) else 56 * n),
                        "steps": e2e_steps, "matches_resident": same, "last_call": e2e_split,
                        "every_record_over_pcie": e2e_plain,
                        "host_copy_ceiling_gbs": (extra.get("host_copy") or {}).get("gbs"),
                        "copy_bound_mtraces": (extra.get("host_copy") or {}).get("copy_bound_mtraces"),
                        "frac_of_copy_bound": (e2e_value / extra["host_copy"]["copy_bound_mtraces"]
                                               if (extra.get("host_copy") or {}).get("copy_bound_mtraces") else None),
                        "host_copy": extra.get("host_copy")},
                "e2e_pageable": extra.get("e2e_pageable"), "gathered": extra.get("gathered"), "strong": extra.get("strong"),
                "rollout": extra.get("rollout"), "other_configs": extra.get("other_configs"),
                "gpu_launches": int(launches), "clocks": clocks, "hit_fraction": hits}
        if world == 1 and not args.no_cpu_baseline:
            v, cores, kind, secs = cpu_reference_run(m.data, starts, ends, min(args.ref_sample, n))
            one = single_core_run(m.data, starts, ends, min(1 << 18, n))
            line["cpu_baseline"] = {"value": v, "unit": "Mtraces/s", "cores": cores, "kind": kind, "single_core": one,
                                    "sample": f"first {min(args.ref_sample, n)} rays of the same stream, one forked "
                                              f"worker per host core, {secs:.2f} s wall"}
        sys.stdout.flush()
        os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rays", type=int, default=N_RAYS, help="rays per GPU (default: the 16 Mi of configs[1])")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--ref-sample", type=int, default=1 << 24,
                    help="rays per reference/cpu_baseline step (default: the whole 16 Mi workload, ~17 s of CPU work on 16 cores)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-arms", action="store_true", help="only value / e2e / roofline: skip host-copy ceiling, pageable, gathered, strong-scaling and rollout arms")
    ap.add_argument("--parity-sample", type=int, default=1 << 21, help="records checked against the reference before timing")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
