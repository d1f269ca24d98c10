"""Development aid: whole-iteration kernel timeline (%globaltimer marks) of ONE chosen iteration inside a ramped registration."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lrr_b200 import synth, lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
targets = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "17").split(",")]
pair = synth.make_pair(3, n, n)
L = lib(); ctx = L.context(0)
fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
est = ctx.estimator("rrl_sym", mv, fx, 12)
f = L.dll.lrr_debug_timeline; f.restype = C.c_int; f.argtypes = [C.c_int, C.c_void_p]
NAMES = {5: "k_match_cert", 0: "k_match(_tile)", 1: "k_match_hard", 6: "k_norm1", 7: "k_norm2", 4: "k_weight", 2: "k_accumulate", 8: "k_resid", 11: "k_estimate"}
for rep in range(2):
    est.set_transform(*pair["init"])
    for it in range(max(targets) + 1):
        arm = rep == 1 and it in targets
        if arm:
            f(1, None)
        est.one_iteration(it)
        if arm:
            out = np.zeros(12 * 2048 * 4, dtype=np.uint64)
            f(0, out.ctypes.data_as(C.c_void_p))
            tl = out.reshape(12, 2048, 4).astype(np.float64)
            rows = []
            for kid, name in NAMES.items():
                t = tl[kid]; live = t[:, 0] > 0
                if not live.any():
                    continue
                t = t[live]
                end = np.where(t[:, 3] > 0, t[:, 3], np.where(t[:, 2] > 0, t[:, 2], t[:, 1]))
                rows.append((t[:, 0].min(), name, live.sum(), t[:, 0].min(), np.median(t[:, 1][t[:, 1] > 0]) if (t[:, 1] > 0).any() else 0, t[:, 1].max(), end.max()))
            rows.sort()
            T0 = rows[0][0]
            print(f"iteration {it} ({est.state().number_points[0]} points): us from the first block start")
            for _, name, nb, s0, lm, ll, e in rows:
                print(f"  {name:16s} blocks {nb:5d} start {(s0 - T0) / 1e3:7.1f} | loop end median {(lm - T0) / 1e3:7.1f} last {(ll - T0) / 1e3:7.1f} | end {(e - T0) / 1e3:7.1f}")
