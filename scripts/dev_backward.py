import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lrr_b200 import synth, lib
from lrr_b200.capi import Lib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pair = synth.make_pair(54, 9000, 8500, size=100.0)
es = []
for L in (lib(), Lib(os.path.join(ROOT, "oracle", "liblrr_oracle.so"), "lrro_")):
    ctx = L.context(); e = ctx.estimator("rrl_sym", ctx.upload(pair["moving"]), ctx.upload(pair["fixed"]), 12)
    e.set_flags(True, True); e.set_transform(*pair["init"])
    for d in (0, 1): e.set_number_points(10 ** 9, d)
    es.append(e)
g, o = es
def rel(a, b): return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))
nit = int(os.environ.get("NIT", "6"))
for it in range(nit):
    g.one_iteration(it); o.one_iteration(it)
    if os.environ.get("METRICS"):
        for d in (0, 1): g.metric("weighted", d); o.metric("weighted", d)
print("before backward: w rel", [rel(g.get_matches(d)[2], o.get_matches(d)[2]) for d in (0, 1)], "scale", list(g.state().scale), list(o.state().scale))
g.estimate_ls_backward(False); o.estimate_ls_backward(False)
print("after backward: w rel", [rel(g.get_matches(d)[2], o.get_matches(d)[2]) for d in (0, 1)], "scale", list(g.state().scale), list(o.state().scale))
sg, so = g.state(), o.state()
print("norm g", list(sg.norm_center_moving), sg.norm_radius_moving, list(sg.norm_center_fixed), sg.norm_radius_fixed)
print("norm o", list(so.norm_center_moving), so.norm_radius_moving, list(so.norm_center_fixed), so.norm_radius_fixed)
for d in (0, 1):
    G, O = g.get_transform(d), o.get_transform(d)
    print(d, "A", rel(G["A"], O["A"]), "cov", rel(G["cov"], O["cov"]))
