"""Timing of the CDC E-step (lrr_cdc_compute_weights: k_cdc_match + k_cdc_normalise) and of both objectives over the many-to-many set.
Usage: python scripts/cdc_time.py [n_features] [cube_mm]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lrr_b200 import synth, lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
size = float(sys.argv[2]) if len(sys.argv) > 2 else 300.0
pair = synth.make_pair(3, n, n, size=size)
ctx = lib().context(0)
fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
est = ctx.estimator("cdcl_fwd", mv, fx, 12); est.set_flags(True, True); est.set_transform(*pair["init"])
for it in range(3): est.one_iteration(it)
s = est.cdc_compute_weights()
print("n", n, "cube", size, s)
ctx.set_profiling(True)
reps = 10
ctx.synchronize(); t0 = time.perf_counter()
for _ in range(reps): est.cdc_compute_weights()
ctx.synchronize(); dt = (time.perf_counter() - t0) / reps
ph = ctx.phase_times()
km = ph["match"][0] / ph["match"][1]
# algorithmic bytes: per from-feature 24 (x) + 48 (Cp) read, 16 written; per candidate in the ball 32 (pos4); per evaluated candidate 64 (covariance
# record) ; per kept entry 16 written + 8 atomic
print(f"compute_weights: {1e3 * dt:.3f} ms per call (host), k_cdc_match {1e3 * km:.1f} us; entries/s {s['entries'] / (km * 1e-3):.3e}")
T = est.get_transform(0)
U = np.linalg.cholesky(T["cov"]).T; x = np.concatenate([U[i, i:] for i in range(12)])
th = np.concatenate([T["A"].reshape(9), T["t"]])
for name, fn, arg in (("objcov", est.objcov_eval, x), ("objpar", est.objpar_eval, th)):
    fn(arg); ctx.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn(arg)
    ctx.synchronize(); d2 = (time.perf_counter() - t0) / reps
    print(f"{name} over {s['entries']} entries: {1e6 * d2:.1f} us per evaluation (host, incl. read-back)")
