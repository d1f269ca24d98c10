"""Per-kernel CUDA-event times of every iteration of one ramped registration (C3 workload)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lrr_b200 import synth, lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
kind = sys.argv[2] if len(sys.argv) > 2 else "rrl_sym"
pair = synth.make_pair(3 if n <= 200000 else 5, n, n)
ctx = lib().context(0)
fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
est = ctx.estimator(kind, mv, fx, 12)
for r in range(2):
    est.set_transform(*pair["init"])
    if "fwd" in kind:
        est.set_flags(True, True)
    it, conv = 0, False
    while not conv and it < 70:
        ctx.set_profiling(True)
        t0 = time.perf_counter(); conv = est.one_iteration(it); dt = time.perf_counter() - t0
        ph = ctx.phase_times(); ctx.set_profiling(False)
        st = est.state() if "sym" in kind else None
        if r == 1:
            print(it, int(st.number_points[0]) if st else n, f"wall {1e6 * dt:.0f} us |", " ".join(f"{k}:{1e3 * v[0]:.0f}({v[1]})" for k, v in ph.items()), est.stats())
        it += 1
