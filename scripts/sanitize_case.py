"""Small end-to-end case for compute-sanitizer: every kernel family once (forward, symmetric, rigid, metrics, objcov, batch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lrr_b200 import synth, lib
L = lib(); ctx = L.context(0)
pair = synth.make_pair(61, 2500, 2200, size=70.0)
fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
for kind, dof in (("cdcl_fwd", 12), ("rrl_sym", 12), ("cdcl_sym", 12)):
    e = ctx.estimator(kind, mv, fx, dof)
    e.set_flags(True, True); e.set_transform(*pair["init"])
    if "fwd" in kind: e.set_number_points(2200)
    for it in range(10 if kind == "rrl_sym" else 4): e.one_iteration(it)   # rrl_sym: far enough for the certificate pass, the normalisation skip and the mailbox
    for m in ("rms", "weighted", "median", "sheet_angles", "tube_angles"): e.metric(m)
    if "sym" in kind: e.estimate_ls_backward(False); e.transfer_cov(0)
    if kind == "cdcl_fwd":
        U = np.linalg.cholesky(e.get_transform(0)["cov"]).T
        e.objcov_eval(np.concatenate([U[i, i:] for i in range(12)])); e.objcov_refine(3)
    e.get_matches(0); e.close()
pr = synth.make_pair(62, 2500, 2200, size=70.0, rigid=True)
fr, mr = ctx.upload(pr["fixed"]), ctx.upload(pr["moving"])
e = ctx.estimator("cdcl_sym", mr, fr, 6); e.set_transform(*pr["init"])
for it in range(3): e.one_iteration(it)
bp = synth.make_pair(4, 20000, 20000, size=100.0)
jobs = synth.make_jobs(bp, 2, 2, seed=4)
out = ctx.batch_run(ctx.upload(bp["moving"]), ctx.upload(bp["fixed"]), jobs, 12, 3)
print("sanitize case done", out["status"].tolist(), out["iterations"].tolist())
