"""Time slicing of the batched mode: the 8 shares of the fixed C4 list run alone on one device (the N = 8 makespan is the slowest share),
for several quanta and both ways of dealing the shares."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from lrr_b200 import lib
from lrr_b200.sharding import job_shard
pair, jobs = bench.make_batch_workload(1)
ctx = lib().context(0)
fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
ctx.batch_run(mv, fx, jobs[:1024], 70, bench.BATCH_SEED)
costs = ctx.batch_expected_cost(mv, fx, jobs)
def timed(j, ids=None):
    ctx.synchronize(); t0 = time.perf_counter(); out = ctx.batch_run(mv, fx, j, 70, bench.BATCH_SEED, ids=ids); ctx.synchronize(); return time.perf_counter() - t0, out
W = 8
qs = [int(x) for x in sys.argv[1:]] or [0, 6, 8, 10, 12]
for use_costs in (True, False):
    for q in qs:
        ctx.set_option("batch_quantum", q)
        ts = []
        for r in range(W):
            ids = job_shard(len(jobs), W, r, costs if use_costs else None)
            dt, _ = timed(jobs[ids], ids); ts.append(dt)
        print(f"{'sorted deal' if use_costs else 'list order '} q={q:2d}: max {max(ts):.4f} -> {len(jobs) / max(ts):6.0f} reg/s | " + " ".join(f"{t:.3f}" for t in ts))
