"""Per-block timeline of k_match / k_match_hard / k_accumulate at the finest level (development aid; %globaltimer marks).
marks: 0 block start, 1 end of the element loop, 2 finalizer entered (last block only; accumulate: reduce returned), 3 block end."""
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
it, conv = 0, False
while not conv and it < 70:
    conv = est.one_iteration(it); it += 1
for k in range(int(os.environ.get("LRR_PROF_SETTLE", "3"))):
    est.one_iteration(it); it += 1
f = L.dll.lrr_debug_timeline; f.restype = C.c_int; f.argtypes = [C.c_int, C.c_void_p]
f(1, None)
est.one_iteration(it)
NK = 12
out = np.zeros(NK * 2048 * 4, dtype=np.uint64)
f(0, out.ctypes.data_as(C.c_void_p))
tl = out.reshape(NK, 2048, 4).astype(np.float64)
KERNELS = [(5, "k_match_cert"), (0, "k_match(_tile)"), (1, "k_match_hard"), (6, "k_norm1"), (7, "k_norm2"), (4, "k_weight"), (2, "k_accumulate"), (8, "k_resid")]
# whole-iteration timeline on the global timer: first block start / median and last loop end / last block end of every kernel
marks = np.concatenate([tl[k][tl[k][:, 0] > 0][:, 0] for k, _ in KERNELS if (tl[k][:, 0] > 0).any()])
T0 = marks.min()
print("iteration timeline (us from the first block start of the first kernel):")
for kid, name in KERNELS:
    t = tl[kid]; live = t[:, 0] > 0
    if not live.any():
        continue
    t = t[live]
    end = np.where(t[:, 3] > 0, t[:, 3], np.where(t[:, 2] > 0, t[:, 2], t[:, 1]))
    r = lambda x: (x - T0) / 1e3
    print(f"  {name:16s} blocks {live.sum():5d} first start {r(t[:,0].min()):7.1f} last start {r(t[:,0].max()):7.1f} | loop end median {r(np.median(t[:,1])):7.1f} last {r(t[:,1].max()):7.1f} | last block end {r(end.max()):7.1f}")
for kid, name in KERNELS:
    t = tl[kid]; live = t[:, 0] > 0
    if not live.any():
        continue
    t = t[live]; t0 = t[:, 0].min()
    rel = lambda x: (x - t0) / 1e3
    end = np.where(t[:, 3] > 0, t[:, 3], np.where(t[:, 2] > 0, t[:, 2], t[:, 1]))
    print(f"{name}: blocks {live.sum()}  start spread {rel(t[:,0].max()):.1f} us | loop end: median {np.median(rel(t[:,1])):.1f} p90 {np.percentile(rel(t[:,1]),90):.1f} max {rel(t[:,1].max()):.1f} us"
          f" | finalizer entered {rel(t[:,2][t[:,2]>0].min()) if (t[:,2]>0).any() else -1:.1f} .. {rel(t[:,2].max()):.1f} us | last block end {rel(end.max()):.1f} us")
    body = (t[:, 1] - t[:, 0]) / 1e3
    print(f"   per-block loop time: min {body.min():.1f} median {np.median(body):.1f} p90 {np.percentile(body,90):.1f} max {body.max():.1f} us")
c = out.reshape(NK, 2048 * 4)[3][:8].astype(np.int64)
if c[0]:
    print(f"tile chunks {c[0]} (both directions): fallback rows {c[1]} cap {c[2]} | queued for pass B {c[3]} queries | per chunk: staged points {c[4]/c[0]:.0f} rows {c[5]/c[0]:.0f} queries {c[6]/c[0]:.1f} | kept by certificate {c[7]} queries")
ph = out.reshape(NK, 2048 * 4)[3][8:16].astype(np.float64)
if ph.sum() > 0:
    names = ["load+map+cand", "row bounds", "stage", "filter", "exact", "check", "cooperative", "finish"]
    print("tile kernel, warp-time per chunk (us): " + ", ".join(f"{n} {v / 1.965e3 / max(c[0], 1):.2f}" for n, v in zip(names, ph)))
cc = out.reshape(NK, 2048 * 4)[3][16:24].astype(np.int64)
print("certificate pass: no previous neighbour", cc[0], "| no certificate", cc[1], "| certificate too weak for the motion", cc[2])

# serial tail of k_accumulate (marks of the finalizer blocks only): reduce returned / pair complete / scalars staged / normal
# equations assembled / inverse / covariance / transform chain / end
t2, t9, t10 = tl[2], tl[9], tl[10]
if (t9 > 0).any():
    base = tl[2][:, 1][tl[2][:, 1] > 0].max()   # last block's loop end
    g = lambda arr, slot: (arr[:, slot][arr[:, slot] > 0].max() - base) / 1e3 if (arr[:, slot] > 0).any() else float("nan")
    print(f"accumulate tail (us after the last loop end): reduce done {g(t9,0):.1f} | pair done {g(t2,2):.1f} | staged {g(t9,1):.1f} | AtA {g(t9,2):.1f} | inverse {g(t9,3):.1f} | cov {g(t10,0):.1f} | transforms {g(t10,1):.1f} | end {g(t2,3):.1f}")
