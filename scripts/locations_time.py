"""The driver's candidate loop on the device (lrr_batch_run_locations) over the fixed C4 list: 256 locations x 32 candidates, all candidates
and until-verified, against the plain job list (lrr_batch_run)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from lrr_b200 import lib
pair, jobs = bench.make_batch_workload(1)
ctx = lib().context(0)
fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
ctx.batch_run(mv, fx, jobs[:1024], 70, bench.BATCH_SEED)
ctx.synchronize(); t0 = time.perf_counter(); plain = ctx.batch_run(mv, fx, jobs, 70, bench.BATCH_SEED); t_plain = time.perf_counter() - t0
print(f"plain list: {t_plain:.4f} s, {len(jobs) / t_plain:.0f} registrations/s")
for uv in (False, True):
    ctx.batch_run_locations(mv, fx, jobs[:32 * 16], bench.BATCH_INITS, 70, bench.BATCH_SEED, until_verified=uv)
    ctx.synchronize(); t0 = time.perf_counter()
    res, dec = ctx.batch_run_locations(mv, fx, jobs, bench.BATCH_INITS, 70, bench.BATCH_SEED, until_verified=uv)
    dt = time.perf_counter() - t0
    ran = int(dec["ran"].sum())
    print(f"candidate loop, until_verified={uv}: {dt:.4f} s, {ran} registrations run ({ran / dt:.0f}/s), {bench.BATCH_LOCATIONS / dt:.0f} locations/s, verified {int(dec['verified'].sum())}, measured {int(dec['measured'].sum())}")
    if not uv:
        ok = dec["ran"] == 1
        assert np.array_equal(res["iterations"][ok], plain["iterations"][ok])
