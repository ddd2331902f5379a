"""Sharding of independent fits over the GPUs of one box (one process per GPU, torch.distributed).

This is the reference's only parallelism - `parallel::mclapply` over parameter samples / folds
(/root/reference/R/adaptive_sampling.R:645-672, :1301-1320, :2670-2693) - without fork(): every rank
owns a static share of the job list, runs it with ONE topolow_fit_batch call on its GPU, and the
per-job result rows are gathered on every rank.  No data-path collective: fits are independent.
"""
from __future__ import annotations

from typing import Callable, Sequence


def job_cost(job) -> float:
    """Pair updates a job may execute: n(n-1)/2 x n_iter (longest-first scheduling key)."""
    n = len(job["degrees"])
    return 0.5 * n * (n - 1) * float(job["n_iter"])


def partition(costs: Sequence[float], world_size: int):
    """Longest-processing-time-first assignment: list of job-index lists, one per rank."""
    order = sorted(range(len(costs)), key=lambda j: (-costs[j], j))
    load = [0.0] * world_size
    parts = [[] for _ in range(world_size)]
    for j in order:
        r = min(range(world_size), key=lambda x: (load[x], x))
        parts[r].append(j)
        load[r] += costs[j]
    return parts


def run_sharded(jobs, fit_batch: Callable | None = None, device=None):
    """Run `jobs` (dicts for _lib.fit_batch) across the ranks of the default process group and
    return the full result list on every rank.  Without an initialised group: one local batch."""
    import torch.distributed as dist
    if fit_batch is None:
        from . import _lib
        fit_batch = _lib.fit_batch
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return fit_batch(jobs, device=device or 0)
    rank, world = dist.get_rank(), dist.get_world_size()
    parts = partition([job_cost(j) for j in jobs], world)
    mine = parts[rank]
    local = fit_batch([jobs[j] for j in mine], device=device if device is not None else rank)
    gathered = [None] * world
    dist.all_gather_object(gathered, list(zip(mine, local)))
    out = [None] * len(jobs)
    for chunk in gathered:
        for j, r in chunk:
            out[j] = r
    return out
