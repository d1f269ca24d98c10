"""Multi-GPU sharding of the only axis the path shards along: independent registrations
(nodule locations x keypoint-indexing initialisations, the outer loops of LocationRegistration::Run,
rrl/itkLocationRegistration.cxx:2138,2230,2332 — SURVEY.md §8(e)).

One process per GPU.  Every rank holds a full replica of both volumes, runs a contiguous block of the job
list on its own device (no collective on the data path) and the results are gathered once at the end.
Works over NCCL (GPU ranks) and gloo (CPU tests).
"""
from __future__ import annotations

import numpy as np

from .capi import RESULT_DTYPE


def job_block(njobs: int, world: int, rank: int) -> tuple[int, int]:
    """Contiguous block [start, stop) of rank `rank`; sizes differ by at most one.  The 32 initialisations of a
    location are adjacent in the job list, so every block sees the same mix of cheap and expensive jobs."""
    base, extra = divmod(njobs, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def run_batch_sharded(ctx, moving, fixed, jobs: np.ndarray, max_iter: int, seed: int, rank: int = 0, world: int = 1) -> np.ndarray:
    """This rank's share of the jobs.  The per-job shuffle seed is seed + GLOBAL job index, so the result of a job
    does not depend on how the list was sharded."""
    start, stop = job_block(len(jobs), world, rank)
    if stop == start:
        return np.zeros(0, dtype=RESULT_DTYPE)
    return ctx.batch_run(moving, fixed, jobs[start:stop], max_iter, seed + start)


def gather_results(local: np.ndarray, njobs: int, rank: int = 0, world: int = 1, device=None) -> np.ndarray | None:
    """Final result gather (~1.3 kB per job).  Returns the full array on every rank (all_gather)."""
    if world == 1:
        return local
    import torch
    import torch.distributed as dist
    item = RESULT_DTYPE.itemsize
    cap = max(job_block(njobs, world, r)[1] - job_block(njobs, world, r)[0] for r in range(world))
    buf = np.zeros(cap * item, dtype=np.uint8)
    buf[: len(local) * item] = np.frombuffer(local.tobytes(), dtype=np.uint8)
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    full = np.zeros(njobs, dtype=RESULT_DTYPE)
    for r in range(world):
        a, b = job_block(njobs, world, r)
        full[a:b] = np.frombuffer(outs[r].cpu().numpy().tobytes()[: (b - a) * item], dtype=RESULT_DTYPE)
    return full
