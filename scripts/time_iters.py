"""Per-iteration wall time of whole ramped registrations (C3 workload), then the fixed-point steady state.
usage: time_iters.py [n] [kind] [registrations]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lrr_b200 import synth, lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
kind = sys.argv[2] if len(sys.argv) > 2 else "rrl_sym"
nreg = int(sys.argv[3]) if len(sys.argv) > 3 else 3
pair = synth.make_pair(3 if n <= 200000 else 5, n, n)
ctx = lib().context(0)
fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
est = ctx.estimator(kind, mv, fx, 12)
fin = []
for r in range(nreg):
    est.set_transform(*pair["init"])
    if "fwd" in kind:
        est.set_flags(True, True)
    it, conv, rows = 0, False, []
    ctx.synchronize(); t_reg = time.perf_counter()
    while not conv and it < 70:
        t0 = time.perf_counter(); conv = est.one_iteration(it); dt = time.perf_counter() - t0
        st = est.state() if "sym" in kind else None
        rows.append((it, int(st.number_points[0]) if st else n, 1e6 * dt)); it += 1
    t_reg = time.perf_counter() - t_reg
    finest = [u for (_, m, u) in rows if m == n]
    fin += finest[1:]   # (the first finest-level iteration has no certificates yet)
    print(f"registration {r}: {it} iterations, converged {conv}, {1e3 * t_reg:.3f} ms incl. state reads; per iteration us:",
          " ".join(f"{m}:{u:.0f}" for (_, m, u) in rows))
if fin:
    print(f"finest-level iterations inside registrations (first one excluded): n={len(fin)} mean {np.mean(fin):.1f} us median {np.median(fin):.1f} us")
for k in range(20):
    est.one_iteration(it); it += 1
ctx.synchronize(); t0 = time.perf_counter()
for k in range(50):
    est.one_iteration(it); it += 1
ctx.synchronize()
print(f"fixed point, L2 warm: {1e6 * (time.perf_counter() - t0) / 50:.1f} us per iteration")
ctx.set_profiling(True)
for k in range(10):
    est.one_iteration(it); it += 1
print({k: (round(1e3 * v[0] / v[1], 2), v[1]) for k, v in ctx.phase_times().items()})
print("launches", ctx.launch_count())
