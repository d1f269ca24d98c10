"""Development aid: per finest-level iteration of a ramped registration, the certificate statistics (failed / left on the work lists) and wall time."""
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
        est.set_flags(True, True); est.set_number_points(n)
    it, conv = 0, False
    while not conv and it < 70:
        prof = r == 1 and os.environ.get("LRR_STATS_PHASES")
        if prof: ctx.set_profiling(True)
        l0 = ctx.launch_count(); t0 = time.perf_counter(); conv = est.one_iteration(it); dt = time.perf_counter() - t0
        if prof:
            ph = {k: round(1e3 * v[0] / max(v[1], 1), 1) for k, v in ctx.phase_times().items() if v[1]}; ctx.set_profiling(False)
        st = est.state() if "sym" in kind else None
        m = int(st.number_points[0]) if st else n
        if r == 1 and m == n:
            s = est.stats()
            print(f"it {it}: {1e6 * dt:.0f} us, {ctx.launch_count() - l0} launches, cert_failed {s['cert_failed']} left {s['work_list_entries']} passb {s['passb_queries']}" + (f" | {ph}" if prof else ""))
        it += 1
