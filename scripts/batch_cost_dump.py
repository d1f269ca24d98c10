"""Per-job facts of the fixed C4 list for the scheduling study: iterations, ROI sizes, status, the ranking key of the driver's candidate
loop (mean angle at the initial transform) and the expected cost the queue is ordered by -> gpurun_out/batch_jobs.npz"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from lrr_b200 import lib
pair, jobs = bench.make_batch_workload(1)
ctx = lib().context(0)
fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
out = ctx.batch_run(mv, fx, jobs, 70, bench.BATCH_SEED)
cost = ctx.batch_expected_cost(mv, fx, jobs)
res, dec = ctx.batch_run_locations(mv, fx, jobs, bench.BATCH_INITS, 70, bench.BATCH_SEED, until_verified=False)
# time of single jobs: a sample of 64 jobs run alone (one block), to calibrate cost = a + b * iterations * points
rng = np.random.default_rng(0); sample = rng.choice(len(jobs), 64, replace=False); tj = []
for j in sample:
    ctx.synchronize(); t0 = time.perf_counter(); ctx.batch_run(mv, fx, jobs[j:j + 1], 70, bench.BATCH_SEED, ids=np.array([j])); ctx.synchronize(); tj.append(time.perf_counter() - t0)
np.savez_compressed("gpurun_out/batch_jobs.npz", iterations=out["iterations"], nm=out["n_moving_inside"], nf=out["n_fixed_inside"], status=out["status"], converged=out["converged"],
                    cost=cost, rank_angle_average=dec["rank_angle_average"], rank_angles=dec["rank_angles"], sample=sample, sample_seconds=np.array(tj))
print("dumped", len(jobs), "jobs; mean iterations", out["iterations"].mean(), "sample seconds", np.min(tj), np.median(tj), np.max(tj))
