import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lrr_b200 import synth, lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
pair = synth.make_pair(5, n, n)
ctx = lib().context(0)
fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
est = ctx.estimator("cdcl_fwd", mv, fx, 12)
est.set_flags(True, True)
for rep in range(2):
    est.set_transform(*pair["init"]); est.set_number_points(n)
    prev = est.get_transform(0)
    for it in range(20):
        ctx.set_profiling(True)
        t0 = time.perf_counter(); conv = est.one_iteration(it); dt = time.perf_counter() - t0
        ph = ctx.phase_times(); ctx.set_profiling(False)
        T = est.get_transform(0); dA = np.abs(T["A"] - prev["A"]).max(); dtt = np.abs(T["t"] - prev["t"]).max(); prev = T
        if rep == 1:
            print(it, int(conv), f"wall {1e6 * dt:.0f} us dA {dA:.1e} dt {dtt:.1e} |", " ".join(f"{k}:{1e3 * v[0]:.0f}" for k, v in ph.items()), est.stats())
