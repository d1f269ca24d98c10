"""Multi-GPU sharding of the only axis the path shards along: independent registrations
(nodule locations x keypoint-indexing initialisations, the outer loops of LocationRegistration::Run,
rrl/itkLocationRegistration.cxx:2138,2230,2332 — SURVEY.md §8(e)).

One process per GPU.  Every rank holds a full replica of both volumes, runs a block-cyclic shard of the job
list on its own device (no collective on the data path) and the results are gathered once at the end.
Works over NCCL (GPU ranks) and gloo (CPU tests).
"""
from __future__ import annotations

import numpy as np

from .capi import RESULT_DTYPE


def job_block(njobs: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block [start, stop) of rank `rank`; sizes differ by at most one (used for the gather layout)."""
    base, extra = divmod(njobs, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def job_shard(njobs: int, world: int, rank: int, costs=None) -> np.ndarray:
    """Block-cyclic shard: rank r takes jobs r, r + world, r + 2 world, ...  The 32 candidate initialisations of a location are
    adjacent in the list (rrl/itkLocationRegistration.cxx:2230) and differ a lot in cost (a rejected ROI ... 70 iterations), and
    so do locations; spreading every location over all ranks gives each rank the same mix.
    With `costs` (lrr_batch_expected_cost: the features inside a job's two ROI cubes, the same integers on every rank) the cycle runs
    over the jobs sorted longest-expected-first, so every rank gets the same share of large and of small ROIs."""
    if costs is None:
        return np.arange(rank, njobs, world, dtype=np.int64)
    order = np.argsort(-np.asarray(costs, dtype=np.int64), kind="stable")
    return np.sort(order[rank::world]).astype(np.int64)


def expected_costs(ctx, moving, fixed, jobs: np.ndarray, world: int):
    """Per-job expected cost for the sharding, or None (one rank; a library without the entry point, e.g. the CPU oracle in tests)."""
    if world == 1 or not ctx.lib.has("batch_expected_cost"):
        return None
    return ctx.batch_expected_cost(moving, fixed, jobs)


def run_batch_sharded(ctx, moving, fixed, jobs: np.ndarray, max_iter: int, seed: int, rank: int = 0, world: int = 1, costs=None) -> np.ndarray:
    """This rank's share of the jobs.  The per-job shuffle seed is seed + GLOBAL job index (lrr_batch_run_ids), so the result of a
    job does not depend on how the list was sharded."""
    ids = job_shard(len(jobs), world, rank, costs)
    if len(ids) == 0:
        return np.zeros(0, dtype=RESULT_DTYPE)
    if world == 1:
        return ctx.batch_run(moving, fixed, jobs, max_iter, seed)
    return ctx.batch_run(moving, fixed, jobs[ids], max_iter, seed, ids=ids)


def gather_results(local: np.ndarray, njobs: int, rank: int = 0, world: int = 1, device=None, costs=None) -> np.ndarray | None:
    """Final result gather (~1.3 kB per job).  Returns the full array, in job order, on every rank (all_gather)."""
    if world == 1:
        return local
    import torch
    import torch.distributed as dist
    item = RESULT_DTYPE.itemsize
    cap = max(len(job_shard(njobs, world, r, costs)) for r in range(world))
    buf = np.zeros(cap * item, dtype=np.uint8)
    buf[: len(local) * item] = np.frombuffer(local.tobytes(), dtype=np.uint8)
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    full = np.zeros(njobs, dtype=RESULT_DTYPE)
    for r in range(world):
        ids = job_shard(njobs, world, r, costs)
        full[ids] = np.frombuffer(outs[r].cpu().numpy().tobytes()[: len(ids) * item], dtype=RESULT_DTYPE)
    return full
