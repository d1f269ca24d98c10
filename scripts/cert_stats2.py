"""Development aid: certificate / far-query statistics of every finest-level iteration inside a ramped registration."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lrr_b200 import synth, lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
pair = synth.make_pair(3, n, n)
L = lib(); ctx = L.context(0)
fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
est = ctx.estimator("rrl_sym", mv, fx, 12)
est.set_transform(*pair["init"])
f = L.dll.lrr_debug_timeline; f.restype = C.c_int; f.argtypes = [C.c_int, C.c_void_p]
out = np.zeros(12 * 2048 * 4, dtype=np.uint64)
it, conv = 0, False
prev = est.get_transform(0)
while it < 40:
    f(1, None)
    conv = est.one_iteration(it); it += 1
    f(0, out.ctypes.data_as(C.c_void_p))
    c = out.reshape(12, 2048 * 4)[3].astype(np.int64)
    T = est.get_transform(0); dA = np.abs(T["A"] - prev["A"]).max(); dt = np.abs(T["t"] - prev["t"]).max(); prev = T
    if est.state().number_points[0] == n:
        nf = max(c[21] + c[22], 1)
        print(f"it {it - 1:2d} conv {int(conv)} kept {c[7]} failed {c[18]} (far {c[20]}, margin-limited {c[21]}, second-limited {c[22]}, mean motion {c[23] / nf / 1e3:.3f} um, mean B-dq0 {c[24] / nf / 1e3:.1f} um) nocert {c[17]} | flat: queries {c[25]} rows/q {c[26] / max(c[25], 1):.1f} pts/q {c[27] / max(c[25], 1):.1f} | staged chunks {c[0]} queued-B {c[3]} | dA {dA:.1e} dt {dt:.1e}")
