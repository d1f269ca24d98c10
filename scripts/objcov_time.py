"""Development aid: cdcl_obj_fun_cov evaluation time at config-5 size (CUDA events over repeated evaluations)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lrr_b200 import synth, lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
pair = synth.make_pair(5, n, n)
ctx = lib().context(0)
fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
est = ctx.estimator("cdcl_fwd", mv, fx, 12)
est.set_flags(True, True); est.set_transform(*pair["init"]); est.set_number_points(n)
nit = int(os.environ.get("LRR_OBJ_ITERS", "3"))   # 0: run to convergence
it, conv = 0, False
while (it < nit) if nit else (not conv and it < 30):
    conv = est.one_iteration(it); it += 1
print("iterations", it, "converged", conv)
U = np.linalg.cholesky(est.get_transform(0)["cov"]).T
x = np.concatenate([U[i, i:] for i in range(12)])
f0, g0 = est.objcov_eval(x)
ctx.set_profiling(True)
for k in range(reps):
    est.objcov_eval(x)
ph = ctx.phase_times(); ctx.set_profiling(False)
us = 1e3 * ph["objcov"][0] / ph["objcov"][1]
print(f"objcov eval {us:.1f} us, {156 * n / (us * 1e-6) / 1e9:.0f} GB/s algorithmic, frac {156 * n / (us * 1e-6) / 1e9 / 6547.2:.3f}, f {f0:.6f} |g| {np.linalg.norm(g0):.6e}")
