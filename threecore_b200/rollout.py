"""BASELINE config 5: a rollout of many agents that move by issuing player-hull traces every tick.

The reference has no Pmove (code/game is headers only); the trace *stream* is modelled on the shape of
botlib's movement predictor (forward move trace be_aas_move.c:644, step up/down :772, gap check :923) and on stock
Quake 3's slide move: per tick and agent

  1 ground trace        0.25 units down from the current origin
  <= 4 slide traces     origin -> origin + velocity * dt, the velocity clipped against every plane that is hit
  3 step traces         18 units up, forward along the horizontal velocity, 18 units down
  2 probe traces        short sweeps at foot and eye height (what a game does with CM_PointContents)

= 10 traces per agent per tick.  Kinematics are deterministic torch elementwise ops on the GPU; they are OURS, not the
reference's -- parity is pinned where the reference has an interface, at CM_BoxTrace: every batch the rollout issues
can be recorded and replayed through the oracle (tests/test_rollout.py).

Agents are independent, so ranks take contiguous ranges of agent ids (sharding.shard_range): no collective on the
data path.
"""
from __future__ import annotations

import numpy as np
import torch

from . import bspgen

TRACES_PER_TICK = 10
STEP_HEIGHT = 18.0
GROUND_DIST = 0.25
OVERCLIP = 1.001


def _fields(rec: torch.Tensor):
    """views into [n,56] uint8 trace_t records: fraction, endpos, plane normal, startsolid"""
    f = rec[:, 8:12].view(torch.float32).reshape(-1)
    endpos = rec[:, 12:24].view(torch.float32).reshape(-1, 3)
    normal = rec[:, 24:36].view(torch.float32).reshape(-1, 3)
    startsolid = rec[:, 4:8].view(torch.int32).reshape(-1)
    return f, endpos, normal, startsolid


class Rollout:
    """n agents on one GPU.  `trace(starts, ends) -> [n,56] uint8 CUDA tensor` is the engine call; `record`, when set,
    receives (starts, ends, records) of every batch (numpy copies) for replay against the oracle."""

    def __init__(self, engine, n_agents: int, world_mins, world_maxs, seed: int = 5, device=None, dt: float = 0.05,
                 speed: float = 320.0, first_agent: int = 0, record=None):
        self.eng = engine
        self.n = int(n_agents)
        self.dev = device if device is not None else torch.device("cuda", engine.device)
        self.dt = float(dt)
        self.record = record
        # agent i's start state depends only on its global id, so any sharding gives the same agents
        # (two generators, each drawn in agent order: a prefix of the stream does not depend on how many agents follow)
        g, g_yaw = np.random.default_rng(seed), np.random.default_rng(seed + 7919)
        total = first_agent + self.n
        lo = np.asarray(world_mins, np.float64) + 160.0
        hi = np.asarray(world_maxs, np.float64) - 160.0
        hi[2] = min(hi[2], lo[2] + 1024.0)
        org = g.uniform(lo, hi, size=(total, 3))[first_agent:]
        yaw = g_yaw.uniform(0.0, 2 * np.pi, size=total)[first_agent:]
        vel = np.stack([np.cos(yaw) * speed, np.sin(yaw) * speed, np.full(self.n, -60.0)], axis=1)
        self.origin = torch.from_numpy(org.astype(np.float32)).to(self.dev)
        self.velocity = torch.from_numpy(vel.astype(np.float32)).to(self.dev)
        self.out = torch.empty((3 * self.n, 56), dtype=torch.uint8, device=self.dev)
        self.traces = 0
        self.launches = 0
        # constants of a tick, made once (torch.tensor(..., device=...) is a synchronous copy every time it is called)
        c = lambda *xyz: torch.tensor(list(xyz), dtype=torch.float32, device=self.dev)
        self._down, self._eye, self._flat, self._up = c(0.0, 0.0, -1.0), c(0.0, 0.0, 26.0), c(1.0, 1.0, 0.0), c(0.0, 0.0, STEP_HEIGHT)

    def _trace(self, starts: torch.Tensor, ends: torch.Tensor) -> torch.Tensor:
        starts = starts.contiguous()
        ends = ends.contiguous()
        out = self.out[: starts.shape[0]]
        self.eng.box_trace_device(starts, ends, bspgen.PLAYER_MINS, bspgen.PLAYER_MAXS, bspgen.MASK_PLAYERSOLID, out)
        self.traces += starts.shape[0]
        self.launches += 1
        if self.record is not None:
            torch.cuda.synchronize(self.dev)
            self.record(starts.cpu().numpy().copy(), ends.cpu().numpy().copy(), out.cpu().numpy().copy())
        return out

    def tick(self):
        n, dt = self.n, self.dt
        o, v = self.origin, self.velocity
        down = self._down
        # ---- ground trace + the two probes: independent of each other, one launch of 3n traces
        eye = o + self._eye
        starts = torch.cat([o, o, eye])
        ends = torch.cat([o + down * GROUND_DIST, o + down * 2.0, eye + v * (dt * 0.25)])
        rec = self._trace(starts, ends)
        gf, _, gn, _ = _fields(rec[:n])
        on_ground = (gf < 1.0) & (gn[:, 2] > 0.7)
        v = torch.where(on_ground[:, None], v * self._flat, v + down * (800.0 * dt))
        # ---- slide move: 4 bumps, every agent traces every bump (a finished agent sweeps zero distance: position test)
        time_left = torch.full((n,), dt, device=self.dev)
        for _ in range(4):
            end = o + v * time_left[:, None]
            rec = self._trace(o, end)
            f, endpos, nrm, _ = _fields(rec)
            o = endpos.clone()
            hit = f < 1.0
            time_left = torch.where(hit, time_left * (1.0 - f), torch.zeros_like(time_left))
            back = (v * nrm).sum(dim=1)
            back = torch.where(back < 0, back * OVERCLIP, back / OVERCLIP)
            v = torch.where(hit[:, None], v - nrm * back[:, None], v)
        # ---- step: up, forward, down
        up = self._up
        rec = self._trace(o, o + up)
        _, top, _, _ = _fields(rec)
        top = top.clone()
        fwd = v * self._flat * dt
        rec = self._trace(top, top + fwd)
        _, ahead, _, _ = _fields(rec)
        ahead = ahead.clone()
        rec = self._trace(ahead, ahead - up)
        f, landed, nrm, ss = _fields(rec)
        take = (f < 1.0) & (nrm[:, 2] > 0.7) & (ss == 0) & on_ground
        o = torch.where(take[:, None], landed, o)
        # in place: origin and velocity are the static tensors a captured tick (capture()) reads and writes
        self.origin.copy_(o)
        self.velocity.copy_(v)

    def capture(self, warmup: int = 3):
        """Record one tick -- its eight trace launches and the kinematics between them, some seventy kernels -- as a CUDA
        graph; run() then replays it.  A tick is a chain of small dependent launches, so issued one by one it is bound
        by launch latency, not by the GPU.  `warmup` eager ticks come first (they size the engine's workspaces: nothing
        may be allocated while the stream is being captured) and advance the agents like any other tick."""
        assert self.record is None, "recording needs the eager path"
        cur = torch.cuda.current_stream(self.dev)
        side = torch.cuda.Stream(self.dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self.tick()
        cur.wait_stream(side)
        torch.cuda.synchronize(self.dev)
        t0, l0 = self.traces, self.launches
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            self.tick()
        self._tick_counts = (self.traces - t0, self.launches - l0)
        self.traces, self.launches = t0, l0              # capturing executed nothing
        self._graph = g
        return self

    def run(self, ticks: int):
        g = getattr(self, "_graph", None)
        for _ in range(ticks):
            if g is None:
                self.tick()
            else:
                g.replay()
                self.traces += self._tick_counts[0]
                self.launches += self._tick_counts[1]
        return self
