"""Stress: repeat the first two finest-level iterations at 100k many times and compare every run with the first one (bitwise
neighbour sets, weights / transform to 1e-9).  Finds rare races."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lrr_b200 import synth, lib
n = 100000; reps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
pair = synth.make_pair(3, n, n)
G = lib(); ctx = G.context(0)
fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
rel = lambda a, b: float(np.linalg.norm(np.asarray(a) - b) / max(np.linalg.norm(b), 1e-300))
ref = None; bad = 0
for r in range(reps):
    e = ctx.estimator("rrl_sym", mv, fx, 12)
    e.set_transform(*pair["init"])
    e.set_number_points(n, 0); e.set_number_points(n, 1)
    out = []
    for it in range(2):
        e.one_iteration(it)
        for d in (0, 1):
            f, t, w = e.get_matches(d); T = e.get_transform(d)
            out.append((t.copy(), w.copy(), T["A"].copy(), T["t"].copy(), T["cov"].copy()))
    st = e.state(); sc = list(st.scale)
    e.close()
    if ref is None:
        ref = out; continue
    for j, (a, b) in enumerate(zip(out, ref)):
        nd = int(np.count_nonzero(a[0] != b[0])); rw = rel(a[1], b[1]); rA = rel(a[2], b[2])
        if nd or rw > 1e-9 or rA > 1e-9:
            bad += 1
            print(f"rep {r} slot {j} (it {j // 2} dir {j % 2}): nn differ {nd} w rel {rw:.3e} A rel {rA:.3e} scale {sc}")
            break
print("reps", reps, "bad", bad, "window misses?", )
