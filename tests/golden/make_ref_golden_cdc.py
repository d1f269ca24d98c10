"""Golden vectors of the full CDC estimator's E-step, produced by the REFERENCE ITSELF.

cdcl_estimation<3,12>::compute_weights (cdcl/cdcl_estimation.txx:342-545) — covariance-driven radius matching, Mahalanobis
Beaton-Tukey weights, normalisation by row and to-feature sums — compiled UNMODIFIED from /root/reference in
oracle/_ref/liblrr_ref.so (oracle/lrr_ref_cdc.cpp), run on the committed inputs tests/golden/inputs_affine.npz.  Two scenes: the
inputs as they are (a few to-features per search ball) and the same clouds shrunk to 12 % of their extent with the feature
covariances kept (hundreds of to-features per ball: the cap of 100, :397).  The shuffle before the cap is the identity and the
radius search returns ascending indices (Q4, third-party order: unpinned), so the cap keeps the lowest indices.
Frozen as tests/golden/ref_cdc_12.npz; tests/test_cdc.py checks the CPU oracle and (-m gpu) the CUDA path against it.
Regenerate with:  python tests/golden/make_ref_golden_cdc.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from lrr_b200.capi import Lib  # noqa: E402
import make_ref_golden as mrg  # noqa: E402

SCENES = (("wide", 1.0), ("dense", 0.12))


def shrunk(pair, s):
    out = {"fixed": dict(pair["fixed"]), "moving": dict(pair["moving"])}
    for side in ("fixed", "moving"):
        out[side]["xyz"] = np.ascontiguousarray(pair[side]["xyz"] * s)
    A, t, c = pair["init"]
    out["init"] = (A.copy(), t * s, c * s)
    return out


def run_cdc(lib, pair, tag, icp_iterations=3):
    """The one protocol shared by the generator and the checks: a few ICP iterations for a transform with a covariance, the E-step,
    the many-to-many set in CSR form, both CDC objectives over it at the current point and at a perturbed one."""
    ctx = lib.context()
    fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
    est = ctx.estimator("cdcl_fwd", mv, fx, 12)
    est.set_flags(True, True)
    est.set_transform(*pair["init"])
    for it in range(icp_iterations):
        est.one_iteration(it)
    out = {}
    s = est.cdc_compute_weights(100, 5.0)
    out[tag + "_counts"] = np.array([s["rows"], s["entries"], s["min_matches"], s["max_matches"]], dtype=np.int64)
    out[tag + "_avg_largest_eval"] = np.array([s["avg_largest_eval"]])
    rf, rs, ti, w = est.cdc_get_matches()
    out[tag + "_row_from"], out[tag + "_row_start"], out[tag + "_to"], out[tag + "_w"] = rf, rs, ti, w
    T = est.get_transform(0)
    U = np.linalg.cholesky(T["cov"]).T
    x = np.concatenate([U[i, i:] for i in range(12)])
    th = np.concatenate([T["A"].reshape(9), T["t"]])
    rng = np.random.default_rng(5)
    xs = (x, x * (1 + 0.1 * rng.standard_normal(x.size)))
    ths = (th, th + np.concatenate([0.002 * rng.standard_normal(9), 0.05 * rng.standard_normal(3)]))
    for k in range(2):
        f, g = est.objcov_eval(xs[k]); out[f"{tag}_objcov{k}_f"] = np.array([f]); out[f"{tag}_objcov{k}_g"] = g
        f, g = est.objpar_eval(ths[k]); out[f"{tag}_objpar{k}_f"] = np.array([f]); out[f"{tag}_objpar{k}_g"] = g
    if lib.has("test_cdc_normalisation"):   # cdcl_normalize_matches over the many-to-many set (reference only; the backends use it inside the refinements)
        n8 = np.zeros(8)
        import ctypes as C
        assert lib.fn("test_cdc_normalisation")(est.h, n8.ctypes.data_as(C.POINTER(C.c_double))) == 0
        out[tag + "_norm8"] = n8
    return out


def main():
    ref = Lib(os.path.join(ROOT, "oracle", "_ref", "liblrr_ref.so"), "lrrr_")
    pair = mrg.load_inputs(False)
    out = {}
    for tag, s in SCENES:
        out.update(run_cdc(ref, shrunk(pair, s), tag))
        print(tag, out[tag + "_counts"], out[tag + "_avg_largest_eval"])
    np.savez_compressed(os.path.join(HERE, "ref_cdc_12.npz"), **out)


if __name__ == "__main__":
    main()
