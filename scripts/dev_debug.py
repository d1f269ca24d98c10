import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lrr_b200 import synth, lib
from lrr_b200.capi import Lib
kind = sys.argv[1] if len(sys.argv) > 1 else "cdcl_fwd"
dof = int(sys.argv[2]) if len(sys.argv) > 2 else 12
norm = int(sys.argv[3]) if len(sys.argv) > 3 else 1
stren = int(sys.argv[4]) if len(sys.argv) > 4 else 1
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 2
pair = synth.make_pair(11, 6000, 5000, rigid=(dof == 6))
ol = Lib("oracle/liblrr_oracle.so", "lrro_")
res = {}
for name, L in (("gpu", lib()), ("cpu", ol)):
    ctx = L.context(); fx = ctx.upload(pair["fixed"]); mv = ctx.upload(pair["moving"])
    e = ctx.estimator(kind, mv, fx, dof); e.set_flags(bool(norm), bool(stren))
    A, t, c = pair["init"]; e.set_transform(A, t, c)
    if "fwd" in kind: e.set_number_points(3000)
    for it in range(iters):
        cv = e.one_iteration(it)
        s = e.state(); T = e.get_transform(0)
        f, to, w = e.get_matches(0)
        print(name, it, "conv", cv, "scale", list(s.scale), "cm", list(s.norm_center_moving), s.norm_radius_moving, "cf", list(s.norm_center_fixed), s.norm_radius_fixed)
        print("   A", T["A"].ravel()[:4], "t", T["t"], "c", T["c"], "cov00", T["cov"][0, 0], "w", w[:4], "rms", e.metric("rms"), "werr", e.metric("weighted"))
    res[name] = (T, w)
print("dA", np.abs(res["gpu"][0]["A"] - res["cpu"][0]["A"]).max(), "dcov rel", np.linalg.norm(res["gpu"][0]["cov"] - res["cpu"][0]["cov"]) / np.linalg.norm(res["cpu"][0]["cov"]))
