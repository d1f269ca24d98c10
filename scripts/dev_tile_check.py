"""Development check of the finest-level tile match: neighbour indices of find_closest vs the oracle, first call (no candidates) and second call."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lrr_b200 import synth, lib
from lrr_b200.capi import Lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30000
pair = synth.make_pair(int(sys.argv[2]) if len(sys.argv) > 2 else 5, n, n)
G = lib(); O = Lib(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "liblrr_oracle.so"), "lrro_")
es = []
for L in (G, O):
    ctx = L.context(0)
    fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
    e = ctx.estimator("rrl_sym", mv, fx, 12)
    e.set_transform(*pair["init"])
    e.set_number_points(n, 0); e.set_number_points(n, 1)
    es.append(e)
g, o = es
for call in range(2):
    g.find_closest(); o.find_closest()
    for d in (0, 1):
        gf, gt, gw = g.get_matches(d); of, ot, ow = o.get_matches(d)
        bad = np.nonzero(gt != ot)[0]
        print(f"call {call} dir {d}: {len(bad)} of {len(gt)} neighbours differ", bad[:10], gt[bad[:5]], ot[bad[:5]], "from eq", np.array_equal(gf, of))
    print("rms", g.metric("rms", 0), o.metric("rms", 0))
rel = lambda a, b: float(np.linalg.norm(np.asarray(a) - b) / max(np.linalg.norm(b), 1e-300))
for it in range(2):
    g.one_iteration(it); o.one_iteration(it)
    for d in (0, 1):
        gf, gt, gw = g.get_matches(d); of, ot, ow = o.get_matches(d)
        G_, O_ = g.get_transform(d), o.get_transform(d)
        print(f"it {it} dir {d}: nn differ {np.count_nonzero(gt != ot)} w rel {rel(gw, ow):.2e} A {rel(G_['A'], O_['A']):.2e} t {rel(G_['t'], O_['t']):.2e} cov {rel(G_['cov'], O_['cov']):.2e}")
    sg, so = g.state(), o.state()
    print("   scale", list(sg.scale), list(so.scale))
