"""Golden vectors produced by the REFERENCE ITSELF.

`oracle/_ref/liblrr_ref.so` is the reference's own estimator code (cdcl_estimation_ICP_matching_all,
cdcl_estimation_symmetric_ICP_matching_all, rrl_estimation_ICP_matching_all, rrl_estimation_symmetric_ICP_matching_all,
cdcl_trans_affine / _rigid3d, cdcl_obj_fun_cov) compiled UNMODIFIED from /root/reference against oracle/vxl_shim
(recipe: oracle/Makefile).  It only exists in the build container, so its outputs on the committed inputs
(tests/golden/inputs_*.npz) are frozen here as tests/golden/ref_<kind>_<dof>.npz and travel to the GPU box:
tests/test_reference_golden.py checks the CPU oracle (-m "not gpu") and the CUDA path (-m gpu) against them.

Third-party arithmetic that is absent from /root/reference stays restated (oracle/vxl_shim): MUSE (rrel), the k-d tree
(rsdl; exact NN, lowest index wins ties), SVD / eigen-decomposition (vnl_algo), the ideal Voronoi map (ITK), and the
moving-set shuffle is the identity (Q4: order injectable).   Regenerate with:  python tests/golden/make_ref_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lrr_b200.capi import Lib  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [("cdcl_fwd", 12, False), ("cdcl_fwd", 6, True), ("cdcl_sym", 12, False), ("cdcl_sym", 6, True),
         ("rrl_fwd", 12, False), ("rrl_sym", 12, False)]
ITERS = 8
METRICS = ("rms", "weighted", "median", "sheet_angles", "tube_angles")


def load_inputs(rigid):
    z = np.load(os.path.join(HERE, f"inputs_{'rigid' if rigid else 'affine'}.npz"))
    pair = {"fixed": {}, "moving": {}}
    for k in z.files:
        for side in ("fixed", "moving"):
            if k.startswith(side + "_"):
                pair[side][k[len(side) + 1:]] = z[k]
    pair["init"] = (z["init_A"], z["init_t"], z["init_c"])
    return pair


def run_case(lib, pair, kind, dof, iters=ITERS, with_objcov=True):
    """The one protocol shared by the generator and the checks (tests/test_reference_golden.py imports it)."""
    ctx = lib.context()
    fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
    est = ctx.estimator(kind, mv, fx, dof)
    est.set_flags(True, True)
    est.set_transform(*pair["init"])
    dirs = (0, 1) if "sym" in kind else (0,)
    out = {}
    for it in range(iters):
        out[f"it{it}_conv"] = np.array([int(est.one_iteration(it))])
        for d in dirs:
            T = est.get_transform(d)
            f, to, w = est.get_matches(d)
            p = f"it{it}_d{d}_"
            out[p + "A"], out[p + "t"], out[p + "c"], out[p + "cov"] = T["A"], T["t"], T["c"], T["cov"]
            out[p + "from"], out[p + "to"], out[p + "w"] = f, to, w
            out[p + "metrics"] = np.array([est.metric(m, d) for m in METRICS])
        st = est.state()
        out[f"it{it}_state"] = np.array([st.number_points[0], st.number_points[1], st.number_matches, st.finest_level,
                                         st.oscillation_count], dtype=np.int64)
        out[f"it{it}_errstate"] = np.array([st.error_difference, st.weighted_error])
    for d in dirs:
        out[f"end_d{d}_covJ"] = est.transfer_cov(d)
    out["end_oscillated"] = np.array([int(est.oscillated())])
    if with_objcov and kind == "cdcl_fwd" and dof == 12:
        # cdcl_obj_fun_cov<3,12>::compute at x = vec(upper Cholesky factor of the estimated covariance), and at a perturbed x
        cov = est.get_transform(0)["cov"]
        U = np.linalg.cholesky(cov).T
        x = np.concatenate([U[i, i:] for i in range(12)])
        rng = np.random.default_rng(7)
        x2 = x * (1.0 + 0.05 * rng.standard_normal(x.size))
        for tag, xx in (("x0", x), ("x1", x2)):
            f, g = est.objcov_eval(xx, 0)
            out[f"objcov_{tag}"], out[f"objcov_{tag}_f"], out[f"objcov_{tag}_g"] = xx, np.array([f]), g
        # cdcl_obj_fun_par<3,12>::compute at the estimated parameters and at perturbed ones (some matches turn into outliers)
        T = est.get_transform(0)
        th = np.concatenate([T["A"].reshape(9), T["t"]])
        th2 = th + np.concatenate([0.002 * rng.standard_normal(9), 0.5 * rng.standard_normal(3)])
        for tag, xx in (("x0", th), ("x1", th2)):
            f, g = est.objpar_eval(xx, 0)
            out[f"objpar_{tag}"], out[f"objpar_{tag}_f"], out[f"objpar_{tag}_g"] = xx, np.array([f]), g
    if "sym" in kind:
        # estimate_LS_backward(false): decision-measurement path of the driver (rrl/itkLocationRegistration.cxx:1840), followed by
        # the metrics the driver reads next (:1842-1845).  cdcl symmetric <3,6>: the rigid branch (cdcl sym :708-731, :856-862) maps
        # the already mapped clones once more, so these metrics see them through a third map.
        est.estimate_ls_backward(False)
        for d in dirs:
            T = est.get_transform(d)
            out[f"lsb_d{d}_cov"], out[f"lsb_d{d}_A"] = T["cov"], T["A"]
            out[f"lsb_d{d}_w"] = est.get_matches(d)[2]
            out[f"lsb_d{d}_metrics"] = np.array([est.metric(m, d) for m in ("rms", "weighted", "sheet_angles", "tube_angles")])
    return out


def main():
    lib = Lib(os.path.join(ROOT, "oracle", "_ref", "liblrr_ref.so"), "lrrr_")
    for kind, dof, rigid in CASES:
        out = run_case(lib, load_inputs(rigid), kind, dof)
        np.savez_compressed(os.path.join(HERE, f"ref_{kind}_{dof}.npz"), **out)
        print(kind, dof, "A00", out[f"it{ITERS - 1}_d0_A"][0, 0], "conv", [int(out[f"it{i}_conv"][0]) for i in range(ITERS)])


if __name__ == "__main__":
    main()
