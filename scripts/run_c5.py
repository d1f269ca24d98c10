"""Config 5: whole-volume pair, 1 M features per side, affine cdcl_estimation_ICP_matching_all<3,12> at
the finest level, then cdcl_obj_fun_cov<3,12> + L-BFGS (<= 5 evaluations) on the final matches."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lrr_b200 import synth, lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
t0 = time.time(); pair = synth.make_pair(5, n, n); print("generate", round(time.time() - t0, 1), "s")
ctx = lib().context(0)
t0 = time.perf_counter(); fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"]); print("upload+grid", round(time.perf_counter() - t0, 3), "s")
est = ctx.estimator("cdcl_fwd", mv, fx, 12)
est.set_flags(True, True); est.set_transform(*pair["init"]); est.set_number_points(n)
t0 = time.perf_counter(); conv, it = est.run(30); dt = time.perf_counter() - t0
print("run: converged", conv, "iterations", it, "seconds", round(dt, 4), "ms/iter", round(1e3 * dt / it, 3))
ctx.set_profiling(True)
for k in range(10):
    est.one_iteration(it + k)
ph = ctx.phase_times(); ctx.set_profiling(False)
print("phases us:", {k: round(1e3 * v[0] / v[1], 1) for k, v in ph.items()})
tot = sum(v[0] / v[1] for v in ph.values())
B = 124 * n + 24 * n
print("iteration (sum of kernels) ms", round(tot, 3), "algorithmic MB", B / 1e6, "GB/s", round(B / (tot * 1e-3) / 1e9, 1), "frac of 6547", round(B / (tot * 1e-3) / 1e9 / 6547.2, 4))
U = np.linalg.cholesky(est.get_transform(0)["cov"]).T
x = np.concatenate([U[i, i:] for i in range(12)])
ctx.set_profiling(True)
for k in range(5):
    est.objcov_eval(x)
ph = ctx.phase_times(); ctx.set_profiling(False)
us = 1e3 * ph["objcov"][0] / ph["objcov"][1]
print("objcov eval us", round(us, 1), "algorithmic MB", 156 * n / 1e6, "GB/s", round(156 * n / (us * 1e-6) / 1e9, 1), "frac", round(156 * n / (us * 1e-6) / 1e9 / 6547.2, 4))
t0 = time.perf_counter(); r = est.objcov_refine(5); print("objcov_refine", r, round(time.perf_counter() - t0, 4), "s")
