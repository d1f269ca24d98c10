import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lrr_b200 import synth, lib
pair = synth.make_pair(3, 100000, 100000)
ctx = lib().context(0)
fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
est = ctx.estimator("rrl_sym", mv, fx, 12)
A, t, c = pair["init"]
est.set_transform(A, t, c)
est.one_iteration(0)
ctx.set_profiling(True)
for k in range(1, 5):
    est.one_iteration(k)
print("M=1000:", {k: (round(1e3 * v[0] / v[1], 2), v[1]) for k, v in ctx.phase_times().items()})
