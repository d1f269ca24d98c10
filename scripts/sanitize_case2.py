"""Second compute-sanitizer case: the CDC E-step (caps, overfull candidate lists, the many-to-many objectives and refinements) and the
time-sliced batched mode (more jobs than blocks, slots changing tenants)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lrr_b200 import synth, lib
L = lib(); ctx = L.context(0)
for size, nf, nm in ((60.0, 3000, 2700), (20.0, 6000, 300)):
    pair = synth.make_pair(71, nf, nm, size=size)
    fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
    e = ctx.estimator("cdcl_fwd", mv, fx, 12); e.set_flags(True, True); e.set_transform(*pair["init"])
    for it in range(2): e.one_iteration(it)
    for cap, mult in ((100, 5.0), (5, 5.0), (1000, 5.0)):
        s = e.cdc_compute_weights(cap, mult); e.cdc_get_matches()
    T = e.get_transform(0); U = np.linalg.cholesky(T["cov"]).T
    e.objcov_eval(np.concatenate([U[i, i:] for i in range(12)])); e.objpar_eval(np.concatenate([T["A"].reshape(9), T["t"]]))
    e.objpar_refine(2); e.objcov_refine(2); e.cdc_one_iteration(7)
    print("cdc", size, s)
    e.close()
bp = synth.make_pair(4, 20000, 20000, size=100.0)
jobs = synth.make_jobs(bp, 40, 5, seed=4)
bm, bf = ctx.upload(bp["moving"]), ctx.upload(bp["fixed"])
for q, slots in ((2, 160), (8, 1024)):
    ctx.set_option("batch_quantum", q); ctx.set_option("batch_slots", slots)
    out = ctx.batch_run(bm, bf, jobs, 12, 3)
    print("batch", q, slots, np.bincount(out["status"]), int(out["iterations"].sum()))
