"""E-step of the full CDC estimator: cdcl_estimation<3,12>::compute_weights (cdcl/cdcl_estimation.txx:342-545) — SURVEY §8(f)3.

Covariance-driven radius matching (radius 5 sqrt(lambda_max(Jth Sigma Jth^T + Jp Cp Jp^T)), rsdl points_in_radius, at most 100
to-features per from-feature), Mahalanobis Beaton-Tukey weights, normalisation by row and to-feature sums; then both CDC objectives
(cdcl_obj_fun_cov / cdcl_obj_fun_par) and the refinements (estimate_covariance / estimate_parameters) over the many-to-many set, and
cdcl_estimation::one_iteration (:88-186).

Pins: tests/golden/ref_cdc_12.npz = outputs of the reference's own cdcl_estimation.txx compiled unmodified (oracle/lrr_ref_cdc.cpp,
generator tests/golden/make_ref_golden_cdc.py); live comparison oracle <-> compiled reference where oracle/_ref exists.
CPU: the oracle reproduces the reference.  GPU (-m gpu): the CUDA path reproduces the reference golden and the oracle.
Gates: rows / entries / CSR index arrays identical; weights, f, g 1e-9 relative (observed 1e-13); refined transforms 1e-5.
"""
import ctypes as C
import os
import sys

import numpy as np
import pytest

from lrr_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.join(ROOT, "tests", "golden")
sys.path.insert(0, HERE)
import make_ref_golden as mrg  # noqa: E402
import make_ref_golden_cdc as mrc  # noqa: E402

REF_SO = os.path.join(ROOT, "oracle", "_ref", "liblrr_ref.so")
TOL = 1e-9


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def check_golden(lib, tol=TOL):
    gold = np.load(os.path.join(HERE, "ref_cdc_12.npz"))
    pair = mrg.load_inputs(False)
    for tag, s in mrc.SCENES:
        out = mrc.run_cdc(lib, mrc.shrunk(pair, s), tag)
        for k in gold.files:
            if not k.startswith(tag + "_") or k.endswith("_norm8"):
                continue
            g, o = gold[k], out[k]
            if k.endswith(("_counts", "_row_from", "_row_start", "_to")):
                assert np.array_equal(o, g), (k, o if o.size < 8 else np.count_nonzero(o != g))
            else:
                assert rel(o, g) < tol, (k, rel(o, g))
    # the cap of 100 is exercised by the dense scene (97 of the 100 lowest indices keep a positive weight)
    assert gold["dense_counts"][3] > 90 and gold["wide_counts"][3] < 50


def test_oracle_reproduces_reference_cdc_golden(oracle_lib):
    check_golden(oracle_lib, 1e-11)


@pytest.mark.gpu
def test_gpu_reproduces_reference_cdc_golden(gpu_lib):
    check_golden(gpu_lib)


def make(lib, pair, subsets=None, icp=3, flags=(True, True)):
    ctx = lib.context()
    fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
    est = ctx.estimator("cdcl_fwd", mv, fx, 12)
    est.set_flags(*flags)
    if subsets is not None:
        est.set_subsets(*subsets)
    est.set_transform(*pair["init"])
    for it in range(icp):
        est.one_iteration(it)
    return est


def compare_sets(a, b, tol=TOL):
    sa, sb = a.cdc_compute_weights(), b.cdc_compute_weights()
    for k in ("rows", "entries", "min_matches", "max_matches"):
        assert sa[k] == sb[k], (k, sa, sb)
    assert sa["avg_largest_eval"] == pytest.approx(sb["avg_largest_eval"], rel=1e-10)
    ma, mb = a.cdc_get_matches(), b.cdc_get_matches()
    for x, y in zip(ma[:3], mb[:3]):
        assert np.array_equal(x, y)
    assert rel(ma[3], mb[3]) < tol
    return sa


def compare_objectives(a, b, tol=TOL):
    T = b.get_transform(0)
    U = np.linalg.cholesky(T["cov"]).T
    x = np.concatenate([U[i, i:] for i in range(12)])
    th = np.concatenate([T["A"].reshape(9), T["t"]])
    rng = np.random.default_rng(9)
    for xx in (x, x * (1 + 0.1 * rng.standard_normal(x.size))):
        (fa, ga), (fb, gb) = a.objcov_eval(xx), b.objcov_eval(xx)
        assert fa == pytest.approx(fb, rel=tol) and rel(ga, gb) < tol
    for tt in (th, th + np.concatenate([0.002 * rng.standard_normal(9), 0.05 * rng.standard_normal(3)])):
        (fa, ga), (fb, gb) = a.objpar_eval(tt), b.objpar_eval(tt)
        assert fa == pytest.approx(fb, rel=tol) and rel(ga, gb) < tol


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/liblrr_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("size,n,sub", [(300.0, 2500, False), (45.0, 3000, False), (60.0, 3000, True)])
def test_oracle_matches_compiled_reference(oracle_lib, size, n, sub):
    """live: the reference's own compute_weights / cdcl_obj_fun_cov / cdcl_obj_fun_par over the many-to-many set"""
    from lrr_b200.capi import Lib
    ref = Lib(REF_SO, "lrrr_")
    pair = synth.make_pair(71, n, n - 300, size=size)
    subsets = None
    if sub:
        rng = np.random.default_rng(4)
        subsets = (rng.permutation(n - 300)[:1500].astype(np.int32), np.zeros(0, np.int32))
    o, r = make(oracle_lib, pair, subsets), make(ref, pair, subsets)
    s = compare_sets(o, r, 1e-11)
    assert s["entries"] > s["rows"] > 0
    compare_objectives(o, r, 1e-10)
    # cdcl_normalize_matches over the many-to-many set: the reference's numbers against the oracle's (read back through objpar_refine's state)
    n8 = np.zeros(8)
    assert ref.fn("test_cdc_normalisation")(r.h, n8.ctypes.data_as(C.POINTER(C.c_double))) == 0
    o.objpar_refine(1)
    st = o.state()
    got = np.array(list(st.norm_center_moving) + [st.norm_radius_moving] + list(st.norm_center_fixed) + [st.norm_radius_fixed])
    assert rel(got, n8) < 1e-11


def test_resolution_switch_decision(oracle_lib):
    """cdcl_estimation.txx:494-508: finer below 2 spacing[res], coarser above 2 spacing[res + 1], never both, only when allowed"""
    sp = np.array([1.0, 2.0, 4.0])
    fn = oracle_lib.fn("cdc_resolution_switch")
    def sw(avg, res, allowed=1):
        out = C.c_int(-1)
        assert fn(avg, sp.ctypes.data_as(C.POINTER(C.c_double)), res, 2, allowed, C.byref(out)) == 0
        return out.value
    assert sw(3.0, 1) == 0          # 3 < 2 * 2: one level finer
    assert sw(5.0, 1) == 1          # 4 <= 5 <= 8: stays
    assert sw(9.0, 1) == 2          # 9 > 2 * 4: one level coarser
    assert sw(1.0, 0) == 0          # already finest
    assert sw(100.0, 2) == 2        # already coarsest
    assert sw(3.0, 1, allowed=0) == 1


# ---------------------------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("size,n,sub", [(300.0, 6000, False), (45.0, 3000, False), (60.0, 5000, True), (30.0, 4000, False)])
def test_gpu_compute_weights_matches_oracle(gpu_lib, oracle_lib, size, n, sub):
    """identical CSR index arrays, weights / objectives 1e-9; the 30 mm scene holds ~1000 to-features per ball (cap, bisection on the index)"""
    pair = synth.make_pair(73, n, n - 300, size=size)
    subsets = None
    if sub:
        rng = np.random.default_rng(4)
        subsets = (rng.permutation(n - 300)[:2500].astype(np.int32), np.zeros(0, np.int32))
    g, o = make(gpu_lib, pair, subsets), make(oracle_lib, pair, subsets)
    s = compare_sets(g, o)
    assert s["entries"] > s["rows"] > 0
    compare_objectives(g, o)
    # the resolution switch is host arithmetic on avg_largest_eval: same decision from both libraries
    sp = np.array([0.5 * s["avg_largest_eval"], 4.0 * s["avg_largest_eval"]])
    for lib in (gpu_lib, oracle_lib):
        out = C.c_int(-1)
        assert lib.fn("cdc_resolution_switch")(s["avg_largest_eval"], sp.ctypes.data_as(C.POINTER(C.c_double)), 1, 1, 1, C.byref(out)) == 0
        assert out.value == 0


@pytest.mark.gpu
def test_gpu_cap_and_multiplier_arguments(gpu_lib, oracle_lib):
    """max_matches and the radius multiplier are arguments of the backend (constants 100 / 5.0 in the reference): small caps, small radii,
    the capacity retry of the flat set (first guess 48 entries per from-feature) and a second call on the same estimator"""
    pair = synth.make_pair(75, 3000, 2800, size=40.0)
    g, o = make(gpu_lib, pair), make(oracle_lib, pair)
    for cap, mult in ((100, 5.0), (7, 5.0), (100, 1.5), (1, 2.0), (64, 8.0)):
        sg, so = g.cdc_compute_weights(cap, mult), o.cdc_compute_weights(cap, mult)
        assert (sg["rows"], sg["entries"], sg["min_matches"], sg["max_matches"]) == (so["rows"], so["entries"], so["min_matches"], so["max_matches"]), (cap, mult, sg, so)
        mg, mo = g.cdc_get_matches(), o.cdc_get_matches()
        assert all(np.array_equal(x, y) for x, y in zip(mg[:3], mo[:3])), (cap, mult)
        assert rel(mg[3], mo[3]) < TOL
        assert sg["max_matches"] <= cap
    # a new 1-NN match drops the many-to-many set: the objectives run over the ICP matches again
    g.find_closest(); o.find_closest()
    compare_objectives(g, o)


@pytest.mark.gpu
def test_gpu_search_balls_larger_than_the_candidate_list(gpu_lib, oracle_lib):
    """6000 fixed features inside a 20 mm cube: every search ball (radius >= 5 sigma) holds most of them, more than the 4096 entries of a
    warp's candidate list — the list is pruned to the lowest indices while the scan goes on, exactly"""
    pair = synth.make_pair(81, 6000, 600, size=20.0)
    g, o = make(gpu_lib, pair, icp=2), make(oracle_lib, pair, icp=2)
    s = compare_sets(g, o)
    assert s["max_matches"] > 50
    sg, so = g.cdc_compute_weights(1000, 5.0), o.cdc_compute_weights(1000, 5.0)
    assert (sg["rows"], sg["entries"], sg["max_matches"]) == (so["rows"], so["entries"], so["max_matches"])
    mg, mo = g.cdc_get_matches(), o.cdc_get_matches()
    assert all(np.array_equal(x, y) for x, y in zip(mg[:3], mo[:3])) and rel(mg[3], mo[3]) < TOL


@pytest.mark.gpu
def test_gpu_refinements_and_em_iterations(gpu_lib, oracle_lib):
    """estimate_parameters / estimate_covariance over the many-to-many set (normalised as cdcl_estimation_ICP.txx:751-832) and
    cdcl_estimation::one_iteration for iterations 4..9 (the parameter step starts after startEM = 5)"""
    pair = synth.make_pair(77, 4000, 3600, size=80.0)
    g, o = make(gpu_lib, pair), make(oracle_lib, pair)
    compare_sets(g, o)
    rg, ro = g.objpar_refine(5), o.objpar_refine(5)
    assert rg[2] == ro[2] and rg[3] == ro[3]
    assert rg[0] == pytest.approx(ro[0], rel=1e-8) and rg[1] == pytest.approx(ro[1], rel=1e-8)
    sg, so = g.state(), o.state()
    assert rel(list(sg.norm_center_moving) + [sg.norm_radius_moving] + list(sg.norm_center_fixed) + [sg.norm_radius_fixed],
               list(so.norm_center_moving) + [so.norm_radius_moving] + list(so.norm_center_fixed) + [so.norm_radius_fixed]) < 1e-10
    for it in range(4, 10):
        (cg, s1), (co, s2) = g.cdc_one_iteration(it), o.cdc_one_iteration(it)
        assert cg == co
        assert (s1["rows"], s1["entries"]) == (s2["rows"], s2["entries"]), (it, s1, s2)
        G, O = g.get_transform(0), o.get_transform(0)
        for k in ("A", "t", "cov"):
            assert rel(G[k], O[k]) < 1e-5, (it, k, rel(G[k], O[k]))


@pytest.mark.gpu
def test_gpu_compute_weights_at_size(gpu_lib, oracle_lib):
    """C2-sized clouds (50k / 50k): every from-feature's row identical to the oracle's"""
    pair = synth.make_pair(79, 50000, 50000)
    g, o = make(gpu_lib, pair, icp=2), make(oracle_lib, pair, icp=2)
    s = compare_sets(g, o)
    assert s["rows"] > 20000
    compare_objectives(g, o)
