"""Multi-GPU sharding of a trace batch: replicated map, contiguous item ranges, optional result gather.

Every work item is independent and the clip map is read-only, so the data path needs no
collective: rank r traces items [lo_r, hi_r) on its own GPU against its own copy of the map.  The only
communication is an optional gather of the 56-byte trace_t records (all_gather over NCCL/NVLink on GPUs,
gloo on CPU in the tests) for consumers that need every result on every rank.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous range of rank `rank`; ranges differ in size by at most one item and tile [0, n)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    lo = (n * rank) // world
    hi = (n * (rank + 1)) // world
    return lo, hi


def shard_ranges(n: int, world: int) -> list[tuple[int, int]]:
    return [shard_range(n, r, world) for r in range(world)]


def gather_records(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """All-gather per-rank record tensors ([n_r, k], same dtype/k everywhere, ranges from shard_range)
    into the full [n_total, k] tensor on every rank.  Uneven shards are padded to the largest one."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    ranges = shard_ranges(n_total, world)
    lo, hi = ranges[rank]
    if local.shape[0] != hi - lo:
        raise ValueError(f"rank {rank}: {local.shape[0]} records for range [{lo},{hi})")
    width = max(h - l for l, h in ranges)
    pad = torch.zeros((width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: hi - lo] = local
    buf = torch.empty((world * width,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(buf, pad, group=group)
    out = torch.empty((n_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    for r, (l, h) in enumerate(ranges):
        out[l:h] = buf[r * width: r * width + (h - l)]
    return out


class ShardedTracer:
    """One rank of a replicated-map, sharded-rays job.  `trace_fn(starts, ends) -> [n,56] uint8 tensor`
    runs the local engine; tests inject a CPU function so the plumbing is exercised with gloo."""

    def __init__(self, trace_fn, group=None):
        self.trace_fn = trace_fn
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0

    def local_range(self, n_total: int):
        return shard_range(n_total, self.rank, self.world)

    def trace(self, starts: torch.Tensor, ends: torch.Tensor, gather: bool = False):
        """starts/ends are the GLOBAL arrays (every rank holds or can index them); returns this rank's
        records, or all records when gather=True."""
        n = starts.shape[0]
        lo, hi = self.local_range(n)
        local = self.trace_fn(starts[lo:hi], ends[lo:hi])
        if gather and self.world > 1:
            return gather_records(local, n, self.group)
        return local


class _DevicePointer:
    """wraps a raw device pointer for torch.as_tensor (CUDA array interface)"""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}


class PeerResults:
    """All ranks' trace_t records in ONE buffer in the root rank's HBM, filled by the other ranks' trace kernels
    themselves: every rank passes `pointer(lo)` as the result pointer of its launches, so its records travel over
    NVLink while its kernel runs (cmgpu_ipc_* in cm_gpu.h).  There is no gather pass; `finish()` is the stream
    synchronisation plus a barrier that makes the records visible to the root.

    The control plane (handle broadcast, barrier) is torch.distributed on whatever backend the job uses.
    """

    def __init__(self, engine, n_total: int, root: int = 0, group=None):
        self.eng = engine
        self.n = int(n_total)
        self.root = root
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        nbytes = max(self.n, 1) * 56
        box = [None]
        if self.rank == root:
            self.base, handle = engine.ipc_alloc(nbytes)
            box[0] = handle
        if self.world > 1:
            dist.broadcast_object_list(box, src=root, group=group)
        if self.rank != root:
            self.base = engine.ipc_open(box[0])
            # With three or more ranks the root's NVLink ingress is what limits a gathered step (7 x 0.94 GB per 16 Mi-ray
            # step at 8 ranks): the engine then traces in pieces and streams each piece's records while the next is walked
            # (cm_gpu.cu: tracePieces).  Measured: 8 ranks 14.2 -> 9.9 ms per step; 2 ranks 7.3 -> 7.6 ms, so not there.
            engine.set_option("remote_pipeline", 1 if self.world >= 3 else 0)
            engine.set_option("remote_expand_ctas", 148)
        self.nbytes = nbytes

    def local_range(self):
        return shard_range(self.n, self.rank, self.world)

    def pointer(self, first_item: int) -> int:
        return self.base + 56 * int(first_item)

    def finish(self):
        torch.cuda.synchronize(self.eng.device)
        if self.world > 1:
            dist.barrier(group=self.group)

    def records(self) -> torch.Tensor:
        """root only: the [n,56] uint8 view of the shared buffer"""
        if self.rank != self.root:
            raise RuntimeError("the records live on the root rank")
        t = torch.as_tensor(_DevicePointer(self.base, self.nbytes), device=torch.device("cuda", self.eng.device))
        return t[: self.n * 56].view(self.n, 56)

    def close(self):
        if self.world > 1:
            dist.barrier(group=self.group)
        if self.rank == self.root:
            self.eng.ipc_free(self.base)
        else:
            self.eng.ipc_close(self.base)
        self.base = 0
