"""One large map across the GPUs of a box (BASELINE.json configs[3]): exact row-block sharding.

Every rank holds a plan of the SAME problem (same seed, so the same relabelling and edge buckets) and
a full replica of the positions.  The tiles are cut into 2R mega-blocks (R = world size).  One
iteration is a round-robin tournament over the mega-blocks (circle method, 2R-1 rounds): in a round
rank r updates every pair BETWEEN its two mega-blocks (a bipartite job, `Plan.run_job(kind=1)`); the
last round updates the pairs INSIDE mega-blocks 2r and 2r+1 (kind 0).  Mega-blocks of one round are
disjoint, so the jobs of a round commute; after each round the ranks all-gather the mega-blocks they
changed (NCCL over NVLink, 2/R of the 6.4 MB position array per rank at cfg4) and every replica is
current again.  Cooling, the edge MAE and the convergence controller then run on every rank's
identical replica (`Plan.end_iteration`).  The iteration as a whole is one sequential order of the
reference's pair loop (src/optimization.cpp:199-282) - `ShardedMap.enumerate` lists it - exactly as on
one GPU.

`emulate=True` runs the jobs of all ranks one after another on a single GPU (no process group): the
result is bit-identical to the multi-GPU run and is what the single-GPU parity tests check.
"""
from __future__ import annotations

import numpy as np

from . import _lib


def circle_pair(m, rr, q):
    """q-th pair of round rr in a round-robin tournament of m (even) teams (csrc/schedule.h)."""
    big = m - 1
    if q == 0:
        return big, rr
    return (rr + q) % big, (rr - q + big) % big


class _DeviceArray:
    """Minimal __cuda_array_interface__ holder so torch can wrap the plan's position buffer."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


class ShardedMap:
    def __init__(self, initial_positions, degrees, edge_i, edge_j, edge_dist, edge_thresh, n_iter, k0, cooling_rate,
                 c_repulsion, relative_epsilon=1e-4, convergence_window=5, convergence_check_freq=3, *, world_size,
                 rank=0, device=0, precision=_lib.PREC_F32, seed=0, tile_points=0, emulate=False, max_ctas=0):
        self.R, self.rank, self.emulate = int(world_size), int(rank), bool(emulate)
        if self.R < 2:
            raise ValueError("ShardedMap needs world_size >= 2 (use Plan / fit for one GPU)")
        self.seed = int(seed)
        self.plan = _lib.Plan(initial_positions, degrees, edge_i, edge_j, edge_dist, edge_thresh, n_iter, k0,
                              cooling_rate, c_repulsion, relative_epsilon, convergence_window, convergence_check_freq,
                              precision=precision, seed=seed, device=device, tile_points=tile_points,
                              n_shards=self.R, max_ctas=max_ctas)
        lay = self.plan.layout()
        self.M = 2 * self.R
        self.Tm = lay["total_tiles"] // self.M
        assert self.Tm * self.M == lay["total_tiles"]
        self.block_elems = self.Tm * lay["tile_points"] * lay["ndim"]
        self.iteration = 0
        self.n_iter = int(n_iter)
        self.check_freq = int(convergence_check_freq) if int(convergence_check_freq) >= 1 else 10
        self.stopped = False
        self._pos = None
        if not self.emulate:
            import torch
            dt = "<f8" if lay["element_bytes"] == 8 else "<f4"
            arr = _DeviceArray(self.plan.positions_ptr(), (self.M, self.block_elems), dt)
            self._pos = torch.as_tensor(arr, device=f"cuda:{device}")
            self._stage = torch.empty((self.R, 2, self.block_elems), dtype=self._pos.dtype, device=self._pos.device)

    # ---- the GPU-level schedule (identical on every rank) -------------------------------------
    def round_order(self, it):
        return np.random.default_rng([self.seed, 911, int(it)]).permutation(self.M - 1)

    def jobs(self, it):
        """[[(kind, t0, tc, y0, yc) per rank] per round] of iteration `it`."""
        rounds = []
        for rr in self.round_order(it):
            row = []
            for r in range(self.R):
                a, b = circle_pair(self.M, int(rr), r)
                row.append((1, a * self.Tm, self.Tm, b * self.Tm, self.Tm))
            rounds.append(row)
        return rounds

    def enumerate(self, it):
        """The sequential pair order iteration `it` is equivalent to ([pairs][2] original point ids)."""
        out = []
        for row in self.jobs(it):
            for job in row:
                out.append(self.plan.enumerate_job(it, *job))
        for r in range(self.R):
            for mb in (2 * r, 2 * r + 1):
                out.append(self.plan.enumerate_job(it, 0, mb * self.Tm, self.Tm))
        return np.concatenate(out)

    # ---- execution -------------------------------------------------------------------------------
    def _exchange(self, blocks):
        """All ranks changed blocks[r] = (a, b): make every replica current."""
        import torch
        import torch.distributed as dist
        a, b = blocks[self.rank]
        mine = torch.stack([self._pos[a], self._pos[b]])
        dist.all_gather_into_tensor(self._stage.view(-1), mine.view(-1))
        ia = torch.tensor([x[0] for x in blocks], device=self._pos.device)
        ib = torch.tensor([x[1] for x in blocks], device=self._pos.device)
        self._pos.index_copy_(0, ia, self._stage[:, 0])
        self._pos.index_copy_(0, ib, self._stage[:, 1])

    def step(self, n_iters=1):
        """Run `n_iters` iterations (all ranks call this together).  Returns False once the fit stopped."""
        stream = None
        if not self.emulate:
            import torch
            stream = torch.cuda.current_stream().cuda_stream
        for _ in range(n_iters):
            if self.stopped or self.iteration >= self.n_iter:
                return False
            it = self.iteration
            for row in self.jobs(it):
                if self.emulate:
                    for job in row:
                        self.plan.run_job(*job)
                else:
                    self.plan.run_job(*row[self.rank], stream=stream)
                    self._exchange([(j[1] // self.Tm, j[3] // self.Tm) for j in row])
            if self.emulate:
                for r in range(self.R):
                    for mb in (2 * r, 2 * r + 1):
                        self.plan.run_job(0, mb * self.Tm, self.Tm)
            else:
                for mb in (2 * self.rank, 2 * self.rank + 1):
                    self.plan.run_job(0, mb * self.Tm, self.Tm, stream=stream)
                self._exchange([(2 * r, 2 * r + 1) for r in range(self.R)])
            self.plan.end_iteration(stream=stream)
            self.iteration += 1
            # the controller may only stop the fit on a check iteration (src/optimization.cpp:294): read the
            # plan's mapped flag there, so that self.iteration never runs ahead of the device's iteration
            if self.iteration % self.check_freq == 0 or self.iteration % 10 == 0 or self.iteration >= self.n_iter:
                self._sync(stream)
                if self.plan.info()["stopped"]:
                    self.stopped = True
                    return False
        return not self.stopped and self.iteration < self.n_iter

    def _sync(self, stream):
        if self.emulate:
            self.plan.run(0)      # no iterations: records and waits for an event on the plan's own stream
        else:
            import torch
            torch.cuda.current_stream().synchronize()

    def result(self, trace=False):
        if not self.emulate:
            import torch
            torch.cuda.synchronize()
        return self.plan.result(trace=trace)

    def close(self):
        self.plan.close()
