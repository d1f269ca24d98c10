"""lrr_batch_run_multi (one process, one host thread per device, shared job cursor, two launches in flight per device) over the fixed C4 list on
every visible device: results against the single-device call, and the wall time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from lrr_b200 import lib
L = lib()
ndev = torch.cuda.device_count()
pair, jobs = bench.make_batch_workload(1)
ctxs = [L.context(d) for d in range(ndev)]
fxs = [c.upload(pair["fixed"]) for c in ctxs]; mvs = [c.upload(pair["moving"]) for c in ctxs]
single = ctxs[0].batch_run(mvs[0], fxs[0], jobs, 70, bench.BATCH_SEED)
t0 = time.perf_counter(); single = ctxs[0].batch_run(mvs[0], fxs[0], jobs, 70, bench.BATCH_SEED); t1 = time.perf_counter() - t0
L.batch_run_multi(ctxs, mvs, fxs, jobs[:4096], 70, bench.BATCH_SEED)   # warm-up (workspaces of every device)
t0 = time.perf_counter(); multi = L.batch_run_multi(ctxs, mvs, fxs, jobs, 70, bench.BATCH_SEED); t2 = time.perf_counter() - t0
for k in ("status", "converged", "oscillated", "iterations", "n_moving_inside", "n_fixed_inside"):
    assert np.array_equal(multi[k], single[k]), k
for k in ("A", "t", "cov"):
    assert np.allclose(multi[k], single[k], rtol=1e-9, atol=1e-12), k
print(f"{ndev} devices: single-device call {t1:.4f} s ({len(jobs) / t1:.0f}/s), lrr_batch_run_multi {t2:.4f} s ({len(jobs) / t2:.0f} registrations/s, {t1 / t2 / ndev:.2f} of {ndev} x); results identical")
