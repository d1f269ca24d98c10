import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lrr_b200 import synth, lib
pair = synth.make_pair(3, 100000, 100000)
L = lib(); ctx = L.context(0)
for rep in range(3):
    t0 = time.perf_counter(); fx = ctx.upload(pair["fixed"]); t1 = time.perf_counter(); mv = ctx.upload(pair["moving"]); t2 = time.perf_counter()
    est = ctx.estimator("rrl_sym", mv, fx, 12); t3 = time.perf_counter()
    A, t, c = pair["init"]; est.set_transform(A, t, c); t4 = time.perf_counter()
    its = []
    it, conv = 0, False
    while not conv and it < 70:
        ta = time.perf_counter(); conv = est.one_iteration(it); its.append(time.perf_counter() - ta); it += 1
    t5 = time.perf_counter()
    est.close(); fx.close(); mv.close(); t6 = time.perf_counter()
    print(f"upload {t1-t0:.4f} {t2-t1:.4f} est {t3-t2:.4f} settrans {t4-t3:.4f} run {t5-t4:.4f} ({it} its) close {t6-t5:.4f}")
    print("  per-iteration ms:", [round(1e3 * x, 2) for x in its])
