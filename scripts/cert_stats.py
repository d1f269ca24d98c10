"""Per-iteration certificate statistics at the finest level (development aid): how many queries the certificate pass keeps, how many
go to pass B / the staged search, and how far the transform still moves."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lrr_b200 import synth, lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 60
pair = synth.make_pair(3, n, n)
L = lib(); ctx = L.context(0)
fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
est = ctx.estimator("rrl_sym", mv, fx, 12)
est.set_transform(*pair["init"])
it, conv = 0, False
while not conv and it < 70:
    conv = est.one_iteration(it); it += 1
print("converged after", it)
f = L.dll.lrr_debug_timeline; f.restype = C.c_int; f.argtypes = [C.c_int, C.c_void_p]
NK = 12
out = np.zeros(NK * 2048 * 4, dtype=np.uint64)
prev = est.get_transform(0)
for k in range(iters):
    f(1, None)
    est.one_iteration(it); it += 1
    f(0, out.ctypes.data_as(C.c_void_p))
    c = out.reshape(NK, 2048 * 4)[3].astype(np.int64)
    T = est.get_transform(0)
    dA = np.abs(T["A"] - prev["A"]).max(); dt = np.abs(T["t"] - prev["t"]).max(); prev = T
    tl = out.reshape(NK, 2048, 4).astype(np.float64)
    def span(kid):
        t = tl[kid]; live = t[:, 0] > 0
        if not live.any(): return 0.0
        t = t[live]; end = np.where(t[:, 3] > 0, t[:, 3], np.where(t[:, 2] > 0, t[:, 2], t[:, 1]))
        return (end.max() - t[:, 0].min()) / 1e3
    if k < 10 or k % 5 == 0:
        print(f"it {k:3d}: kept {c[7]:6d} weak {c[18]:5d} nocert {c[17]:5d} noprev {c[16]:5d} | tile chunks {c[0]:5d} queued-B {c[3]:5d} | dA {dA:.2e} dt {dt:.2e} | us: cert {span(5):.1f} tile {span(0):.1f} hard {span(1):.1f} acc {span(2):.1f}")
