import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lrr_b200 import synth, lib
from lrr_b200.capi import Lib
pair = synth.make_pair(4, 30000, 30000, size=110.0)
jobs = synth.make_jobs(pair, 3, 3, seed=4)
j = int(sys.argv[1]) if len(sys.argv) > 1 else 8
job = jobs[j]
def roi(f, c, h):
    x = f["xyz"]; m = np.all((x >= (c - h)[:, None]) & (x <= (c + h)[:, None]), axis=0); return np.nonzero(m)[0].astype(np.int32)
mi = roi(pair["moving"], job["roi_center_moving"], job["roi_half_size"]); fi = roi(pair["fixed"], job["roi_center_fixed"], job["roi_half_size"])
ests = []
for L in (lib(), Lib("oracle/liblrr_oracle.so", "lrro_")):
    ctx = L.context(); fx = ctx.upload(pair["fixed"]); mv = ctx.upload(pair["moving"])
    e = ctx.estimator("rrl_sym", mv, fx, 12); e.set_subsets(mi, fi); e.set_shuffle(1, 99 + j); e.set_transform(job["A"].reshape(3, 3), job["t"], job["c"])
    ests.append(e)
g, o = ests
for it in range(40):
    cg, co = g.one_iteration(it), o.one_iteration(it)
    sg, so = g.state(), o.state()
    Tg, To = g.get_transform(0), o.get_transform(0)
    gm, om = g.get_matches(0), o.get_matches(0)
    gb, ob = g.get_matches(1), o.get_matches(1)
    print(it, cg, co, sg.number_matches, so.number_matches, sg.oscillation_count, so.oscillation_count, "werr", sg.weighted_error, so.weighted_error,
          "dA", np.abs(Tg["A"] - To["A"]).max(), "idx diff", int((gm[1] != om[1]).sum()), int((gb[1] != ob[1]).sum()), "scale", sg.scale[0], so.scale[0])
    if cg or co: break
