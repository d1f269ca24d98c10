"""A/B of the queue order of the batched mode on the fixed C4 list (8192 jobs), one GPU:
(a) whole list, jobs in list order (batch_sort 0) against longest-expected-first (batch_sort 1);
(b) the N = 8 makespan emulated on one device: each of the 8 shares is run alone and timed, the slowest share is what an 8-GPU box
    waits for — block-cyclic shares in list order (round 2's earlier protocol), block-cyclic shares sorted inside the rank, and
    shares dealt over the globally sorted list (lrr_b200/sharding.py with costs).  Results of every variant must be identical."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from lrr_b200 import lib
from lrr_b200.sharding import job_shard

pair, jobs = bench.make_batch_workload(1)
ctx = lib().context(0)
fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
ctx.batch_run(mv, fx, jobs[:512], 70, bench.BATCH_SEED)
def timed(j, ids=None):
    ctx.synchronize(); t0 = time.perf_counter(); out = ctx.batch_run(mv, fx, j, 70, bench.BATCH_SEED, ids=ids); ctx.synchronize(); return time.perf_counter() - t0, out
res = {}
ctx.set_option("batch_quantum", 0)
for sort in (0, 1, 0, 1):
    ctx.set_option("batch_sort", sort)
    dt, out = timed(jobs)
    res.setdefault(sort, out)
    print(f"whole list  batch_sort={sort}: {dt:.4f} s  {len(jobs) / dt:.0f} registrations/s")
for k in ("status", "converged", "iterations", "n_moving_inside", "n_fixed_inside"):
    assert np.array_equal(res[0][k], res[1][k]), k
for k in ("A", "t", "cov"):   # (the work-queue order of pass A enters the sums of a job at rounding level, run to run)
    assert np.allclose(res[0][k], res[1][k], rtol=1e-9, atol=1e-12), k
print("results identical with and without the ordering (integer fields exact, transforms to 1e-9)")
t0 = time.perf_counter(); costs = ctx.batch_expected_cost(mv, fx, jobs); print(f"expected cost of 8192 jobs: {1e3 * (time.perf_counter() - t0):.2f} ms; min/median/max", costs.min(), int(np.median(costs)), costs.max())
it = res[1]["iterations"].astype(np.int64); true_cost = it * (res[1]["n_moving_inside"].astype(np.int64) + res[1]["n_fixed_inside"])
print("correlation of the expected cost with iterations x ROI points:", round(float(np.corrcoef(costs, true_cost)[0, 1]), 3))
W = 8
for name, sort, use_costs in (("block-cyclic, list order", 0, False), ("block-cyclic, sorted inside the rank", 1, False), ("dealt over the sorted list", 1, True)):
    ctx.set_option("batch_sort", sort)
    ts = []
    for r in range(W):
        ids = job_shard(len(jobs), W, r, costs if use_costs else None)
        dt, _ = timed(jobs[ids], ids)
        ts.append(dt)
    print(f"N=8 emulated, {name}: shares {min(ts):.4f} .. {max(ts):.4f} s  -> {len(jobs) / max(ts):.0f} registrations/s  (ideal from the whole-list time would be 8 x)")

# time slicing (suspend / resume): a job is parked after `batch_quantum` iterations and the block takes the next unit
ctx.set_option("batch_sort", 1)
for q in (0, 4, 8, 16, 8):
    ctx.set_option("batch_quantum", q)
    dt, out = timed(jobs)
    for k in ("status", "converged", "iterations", "n_moving_inside", "n_fixed_inside"):
        assert np.array_equal(out[k], res[1][k]), (q, k)
    for k in ("A", "t", "cov"):
        assert np.allclose(out[k], res[1][k], rtol=1e-9, atol=1e-12), (q, k)
    ts = []
    for r in range(W):
        ids = job_shard(len(jobs), W, r, costs)
        dts, o2 = timed(jobs[ids], ids)
        assert np.array_equal(o2["iterations"], res[1]["iterations"][ids])
        ts.append(dts)
    print(f"batch_quantum={q}: whole list {dt:.4f} s {len(jobs) / dt:.0f} registrations/s; N=8 emulated shares {min(ts):.4f} .. {max(ts):.4f} s -> {len(jobs) / max(ts):.0f} registrations/s")
