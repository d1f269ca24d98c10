import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lrr_b200 import synth, lib
nloc = int(sys.argv[1]) if len(sys.argv) > 1 else 16
ninit = int(sys.argv[2]) if len(sys.argv) > 2 else 8
pair = synth.make_pair(4, 100000, 100000, size=160.0)
jobs = synth.make_jobs(pair, nloc, ninit, seed=4)
ctx = lib().context(0)
fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
ctx.batch_run(mv, fx, jobs, 70, 4)   # warm-up at full size (memory pool, workspace)
ctx.set_profiling(True)
t0 = time.perf_counter(); out = ctx.batch_run(mv, fx, jobs, 70, 4); dt = time.perf_counter() - t0
print("mode", os.environ.get("LRR_BATCH_MODE", "device"), "jobs", len(jobs), "seconds", round(dt, 4), "reg/s", round(len(jobs) / dt, 1), "launches", ctx.launch_count())
print("status", np.bincount(out["status"] + 2), "mean iters", out["iterations"].mean(), "n_in", out["n_moving_inside"].min(), out["n_moving_inside"].max(), "conv", out["converged"].sum())
print({k: (round(v[0], 2), v[1]) for k, v in ctx.phase_times().items()})
np.save("gpurun_out/batch_%s.npy" % os.environ.get("LRR_BATCH_MODE", "device"), out)
