"""Short driver for ncu: ramp the C3 workload to the finest level, then run a few finest-level iterations."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lrr_b200 import synth, lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
extra = int(sys.argv[2]) if len(sys.argv) > 2 else 3
kind = sys.argv[3] if len(sys.argv) > 3 else "rrl_sym"
pair = synth.make_pair(3, n, n)
ctx = lib().context(0)
fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
est = ctx.estimator(kind, mv, fx, 12)
A, t, c = pair["init"]
est.set_transform(A, t, c)
if "fwd" in kind:
    est.set_flags(True, True); est.set_number_points(n)
it, conv = 0, False
inreg = os.environ.get("LRR_PROF_INREG")   # ncu --profile-from-start off: capture the finest-level iterations INSIDE the registration (2nd onwards)
if inreg:
    import torch
    started = False
    while not conv and it < 70:
        fin = bool(est.state().finest_level)
        if fin and not started and int(inreg) <= 1:       # LRR_PROF_INREG=1: from the first finest-level iteration (no certificates yet)
            ctx.synchronize(); torch.cuda.cudart().cudaProfilerStart(); started = True
        conv = est.one_iteration(it); it += 1
        if fin and not started:                           # LRR_PROF_INREG=2: from the second one (certificates in place, transform moving)
            ctx.synchronize(); torch.cuda.cudart().cudaProfilerStart(); started = True
    ctx.synchronize(); torch.cuda.cudart().cudaProfilerStop()
    print("in-registration capture: iterations", it, "converged", conv, "launches", ctx.launch_count())
    sys.exit(0)
while not conv and it < 70:
    conv = est.one_iteration(it); it += 1
print("ramp iterations", it, "converged", conv, "launches", ctx.launch_count())
for k in range(int(os.environ.get("LRR_PROF_SETTLE", "3"))):   # settle at the finest level (neighbour certificates in place)
    est.one_iteration(it); it += 1
import time
ctx.synchronize(); t0 = time.perf_counter()
for k in range(0 if os.environ.get("LRR_PROF_NOWALL") else 40):   # un-profiled back-to-back iterations (host wall clock; each call ends in a stream synchronisation)
    est.one_iteration(it); it += 1
ctx.synchronize()
if not os.environ.get("LRR_PROF_NOWALL"): print(f"wall per iteration (L2 warm, 40 calls): {1e6 * (time.perf_counter() - t0) / 40:.1f} us")
rng = os.environ.get("LRR_PROF_RANGE")   # ncu --profile-from-start off: capture only the iterations below
if rng:
    import torch
    torch.cuda.cudart().cudaProfilerStart()
else:
    ctx.set_profiling(True)
for k in range(extra):
    est.one_iteration(it + k)
if rng:
    ctx.synchronize(); torch.cuda.cudart().cudaProfilerStop()
if not rng:
    print({k: (round(1e3 * v[0] / v[1], 2), v[1]) for k, v in ctx.phase_times().items()})
print("launches", ctx.launch_count())
