"""Golden vectors of BASELINE.json configs[0] produced by the REFERENCE ITSELF: cdcl_estimation_ICP<2,4> with
cdcl_trans_similarity2d (cdcl/examples/estimate.cxx, `-method 4`) on a synthetic planar point set with normals.

`oracle/_ref/liblrr_ref.so` holds the reference's own cdcl_estimation_ICP.txx / cdcl_trans_similarity2d.cxx compiled UNMODIFIED
against oracle/vxl_shim (oracle/lrr_ref_icp2d.cpp, recipe oracle/Makefile).  It exists only in the build container, so its
outputs on the committed inputs (tests/golden/inputs_c1.npz, made by lrr_b200.synth.make_pair_2d(seed=1, init_ang_deg=9, init_trans=0.12, outliers=0.02)) are frozen here as
tests/golden/ref_c1_icp2d.npz: tests/test_config1_icp2d.py checks the CPU oracle (-m "not gpu") and the CUDA path (-m gpu)
against them.  Regenerate with:  python tests/golden/make_ref_golden_c1.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lrr_b200.capi import Lib  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
FLAGS = [(0, 0), (1, 0), (0, 1), (1, 1)]   # (normalize_matches, weight_by_strength); the demo runs (0, 0)
MAX_ITER = 30                               # cdcl_estimation_ICP::run (cdcl_estimation_ICP.txx:336-346)
METRICS = ("rms", "weighted", "median")


def load_inputs():
    z = np.load(os.path.join(HERE, "inputs_c1.npz"))
    pair = {"fixed": {}, "moving": {}}
    for k in z.files:
        for side in ("fixed", "moving"):
            if k.startswith(side + "_"):
                pair[side][k[len(side) + 1:]] = z[k]
    pair["init"] = (z["init_A"], z["init_t"])
    return pair


def run_case(lib, pair, normalize, wbs, max_iter=MAX_ITER):
    """The one protocol shared by the generator and the checks."""
    ctx = lib.context()
    fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
    est = ctx.estimator("cdcl_icp", mv, fx, 4, 2)
    est.set_flags(bool(normalize), bool(wbs))
    est.set_transform(*pair["init"])
    out = {}
    it = 0
    conv = False
    while not conv and it < max_iter:
        conv = est.one_iteration(it)
        T = est.get_transform(0)
        f, to, w = est.get_matches(0)
        p = f"it{it}_"
        out[p + "A"], out[p + "t"], out[p + "cov"] = T["A"], T["t"], T["cov"]
        out[p + "from"], out[p + "to"], out[p + "w"] = f, to, w
        out[p + "metrics"] = np.array([est.metric(m, 0) for m in METRICS])
        it += 1
    out["iterations"] = np.array([it]); out["converged"] = np.array([int(conv)])
    out["end_covJ"] = est.transfer_cov(0)
    return out


def main():
    from lrr_b200 import synth
    pair = synth.make_pair_2d(1, 2000, 2000, init_ang_deg=9.0, init_trans=0.12, outliers=0.02)
    flat = {"init_A": pair["init"][0], "init_t": pair["init"][1]}
    for side in ("fixed", "moving"):
        for k, v in pair[side].items():
            flat[f"{side}_{k}"] = v
    np.savez_compressed(os.path.join(HERE, "inputs_c1.npz"), **flat)
    lib = Lib(os.path.join(ROOT, "oracle", "_ref", "liblrr_ref.so"), "lrrr_")
    gold = {}
    for nz, wbs in FLAGS:
        out = run_case(lib, load_inputs(), nz, wbs)
        for k, v in out.items():
            gold[f"n{nz}s{wbs}_{k}"] = v
        n = int(out["iterations"][0])
        print("normalize", nz, "strength", wbs, "iterations", n, "converged", int(out["converged"][0]), "A", out[f"it{n - 1}_A"][:2, :2].ravel(),
              "rms", out[f"it{n - 1}_metrics"][0])
    np.savez_compressed(os.path.join(HERE, "ref_c1_icp2d.npz"), **gold)


if __name__ == "__main__":
    main()
