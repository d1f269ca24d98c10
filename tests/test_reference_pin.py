"""Oracle pinned against the REFERENCE ITSELF, live (build container only).

oracle/_ref/liblrr_ref.so = the reference's estimator sources compiled unmodified from /root/reference against
oracle/vxl_shim (oracle/Makefile).  These tests run the oracle and that library on the same seeded inputs and compare
every iteration.  They are skipped where the library does not exist (it needs /root/reference to build; the built .so
travels to the GPU box with the snapshot, so they also run there); the frozen fixtures of
tests/test_reference_golden.py cover the same ground everywhere.
"""
import ctypes as C
import os

import numpy as np
import pytest

from lrr_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "liblrr_ref.so")
pytestmark = pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/liblrr_ref.so not built (needs /root/reference)")

METRICS = ("rms", "weighted", "median", "sheet_angles", "tube_angles")
TOL = 1e-9


@pytest.fixture(scope="module")
def ref_lib():
    from lrr_b200.capi import Lib
    return Lib(REF_SO, "lrrr_")


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def pd(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def make(lib, pair, kind, dof, flags=(True, True), subsets=None):
    ctx = lib.context()
    fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
    est = ctx.estimator(kind, mv, fx, dof)
    est.set_flags(*flags)
    if subsets is not None:
        est.set_subsets(*subsets)
    est.set_transform(*pair["init"])
    return est


def compare(o, r, kind, tol=TOL):
    for d in ((0, 1) if "sym" in kind else (0,)):
        (of, ot, ow), (rf, rt, rw) = o.get_matches(d), r.get_matches(d)
        assert np.array_equal(of, rf) and np.array_equal(ot, rt), (d, np.count_nonzero(ot != rt))
        assert rel(ow, rw) < tol, (d, "w", rel(ow, rw))
        O, R = o.get_transform(d), r.get_transform(d)
        for k in ("A", "t", "cov"):
            assert rel(O[k], R[k]) < tol, (d, k, rel(O[k], R[k]))
        assert np.allclose(O["c"], R["c"], rtol=tol, atol=1e-12)
        for m in METRICS:
            a, b = o.metric(m, d), r.metric(m, d)
            assert (np.isnan(a) and np.isnan(b)) or abs(a - b) <= tol * max(1.0, abs(b)), (d, m, a, b)
    so, sr = o.state(), r.state()
    assert list(so.number_points) == list(sr.number_points)
    assert (so.number_matches, so.finest_level, so.oscillation_count) == (sr.number_matches, sr.finest_level, sr.oscillation_count)
    assert so.weighted_error == pytest.approx(sr.weighted_error, rel=1e-9, abs=1e-12)
    assert so.error_difference == pytest.approx(sr.error_difference, rel=1e-6, abs=1e-9)


CASES = [("cdcl_fwd", 12), ("cdcl_fwd", 6), ("cdcl_sym", 12), ("cdcl_sym", 6), ("rrl_fwd", 12), ("rrl_sym", 12)]


@pytest.mark.parametrize("kind,dof", CASES)
def test_every_iteration_matches_reference(oracle_lib, ref_lib, kind, dof):
    pair = synth.make_pair(41 + dof, 3500, 3000, rigid=(dof == 6))
    o, r = make(oracle_lib, pair, kind, dof), make(ref_lib, pair, kind, dof)
    for it in range(6):
        assert o.one_iteration(it) == r.one_iteration(it)
        compare(o, r, kind)
    assert rel(o.transfer_cov(0), r.transfer_cov(0)) < TOL
    assert o.oscillated() == r.oscillated()


@pytest.mark.parametrize("kind", ["cdcl_sym", "rrl_sym"])
def test_run_to_convergence_matches_reference(oracle_lib, ref_lib, kind):
    """Whole registration: ramp 1000 -> x2 -> all, oscillation logic, convergence test; same iteration count."""
    pair = synth.make_pair(52, 5000, 4500, size=120.0)
    o, r = make(oracle_lib, pair, kind, 12), make(ref_lib, pair, kind, 12)
    co, no = o.run(70)
    cr, nr = r.run(70)
    assert (co, no) == (cr, nr)
    compare(o, r, kind)
    if kind == "rrl_sym":
        o.estimate_ls_backward(False); r.estimate_ls_backward(False)     # decision measurements, driver :1840
        compare(o, r, kind)


@pytest.mark.parametrize("kind,flags", [("cdcl_fwd", (False, False)), ("cdcl_fwd", (True, False)), ("rrl_fwd", (False, True)),
                                        ("rrl_sym", (True, False)), ("cdcl_sym", (True, False))])
def test_flags_match_reference(oracle_lib, ref_lib, kind, flags):
    pair = synth.make_pair(61, 2500, 2200)
    o, r = make(oracle_lib, pair, kind, 12, flags), make(ref_lib, pair, kind, 12, flags)
    for it in range(3):
        assert o.one_iteration(it) == r.one_iteration(it)
        compare(o, r, kind)


def test_subsets_match_reference(oracle_lib, ref_lib):
    """set_features(moving_inside, fixed_inside): ROI subsets in a scrambled order (the order is the Q4 injection point)."""
    pair = synth.make_pair(62, 4000, 4000, size=100.0)
    rng = np.random.default_rng(3)
    mi = rng.permutation(4000)[:2600].astype(np.int32)
    fi = rng.permutation(4000)[:2400].astype(np.int32)
    o, r = make(oracle_lib, pair, "rrl_sym", 12, subsets=(mi, fi)), make(ref_lib, pair, "rrl_sym", 12, subsets=(mi, fi))
    for it in range(5):
        assert o.one_iteration(it) == r.one_iteration(it)
        compare(o, r, "rrl_sym")


def test_objcov_matches_reference(oracle_lib, ref_lib):
    """cdcl_obj_fun_cov<3,12>::compute (cdcl/cdcl_obj_fun_cov.txx:44-134): f and all 78 gradient entries."""
    pair = synth.make_pair(63, 2500, 2200)
    o, r = make(oracle_lib, pair, "cdcl_fwd", 12), make(ref_lib, pair, "cdcl_fwd", 12)
    for it in range(3):
        o.one_iteration(it); r.one_iteration(it)
    U = np.linalg.cholesky(r.get_transform(0)["cov"]).T
    x = np.concatenate([U[i, i:] for i in range(12)])
    rng = np.random.default_rng(1)
    for xx in (x, x * (1 + 0.1 * rng.standard_normal(x.size))):
        fo, go = o.objcov_eval(xx, 0)
        fr, gr = r.objcov_eval(xx, 0)
        assert fo == pytest.approx(fr, rel=1e-10)
        assert rel(go, gr) < 1e-9


def test_objpar_matches_reference(oracle_lib, ref_lib):
    """cdcl_obj_fun_par<3,12>::compute (cdcl/cdcl_obj_fun_par.txx:46-144): f and all 12 gradient entries, at the current parameters
    and at perturbed ones (some matches become outliers: the `inlier` branch, :97)."""
    pair = synth.make_pair(63, 2500, 2200)
    o, r = make(oracle_lib, pair, "cdcl_fwd", 12), make(ref_lib, pair, "cdcl_fwd", 12)
    for it in range(3):
        o.one_iteration(it); r.one_iteration(it)
    T = r.get_transform(0)
    x = np.concatenate([T["A"].reshape(9), T["t"]])
    rng = np.random.default_rng(2)
    for xx in (x, x + np.concatenate([0.002 * rng.standard_normal(9), 0.5 * rng.standard_normal(3)])):
        fo, go = o.objpar_eval(xx, 0)
        fr, gr = r.objpar_eval(xx, 0)
        assert fo == pytest.approx(fr, rel=1e-10)
        assert rel(go, gr) < 1e-9
    # finite differences of the oracle's f (the gradient omits the inlier-set boundary, so a small step on a smooth stretch)
    f0, g0 = o.objpar_eval(x, 0)
    for k in (0, 4, 9, 11):
        h = 1e-6 * max(1.0, abs(x[k]))
        xp, xm = x.copy(), x.copy(); xp[k] += h; xm[k] -= h
        fd = (o.objpar_eval(xp, 0, want_g=False)[0] - o.objpar_eval(xm, 0, want_g=False)[0]) / (2 * h)
        assert abs(fd - g0[k]) <= 1e-4 * max(1.0, abs(g0[k])), (k, fd, g0[k])


@pytest.mark.parametrize("u", [0.0, 0.5, 1.0, 4.0, 15.99, 16.0, 16.01, 100.0])
def test_robust_kernels_match_reference(oracle_lib, ref_lib, u):
    """weight_BT / rho / rho_p (cdcl/cdcl_utils.h:142-204)."""
    for name in ("test_weight_BT", "test_rho", "test_rho_p"):
        fo, fr = getattr(oracle_lib.dll, "lrro_" + name), getattr(ref_lib.dll, "lrrr_" + name)
        fo.restype = fr.restype = C.c_double
        assert fo(C.c_double(u)) == fr(C.c_double(u)), name


def test_mad_scale_matches_reference(oracle_lib, ref_lib):
    """rrel_util_median_abs_dev_scale copy (cdcl :520-536, Q12)."""
    rng = np.random.default_rng(2)
    fo, fr = oracle_lib.dll.lrro_test_mad_scale, ref_lib.dll.lrrr_test_mad_scale
    fo.restype = fr.restype = C.c_double
    for n in (2, 3, 4, 5, 10, 11, 1000, 1001):
        v = rng.standard_normal(n)
        assert fo(pd(v.copy()), C.c_int64(n)) == fr(pd(v.copy()), C.c_int64(n))


def test_transform_models_match_reference(oracle_lib, ref_lib):
    """normalize / unnormalize incl. covariance reparametrisation, inverse, recompose_increment
    (cdcl/cdcl_trans_affine.txx:54-286, cdcl/cdcl_trans_rigid3d.cxx:113-326)."""
    rng = np.random.default_rng(4)
    for kind, dof in ((0, 12), (1, 6)):
        A = np.eye(3) + 0.1 * rng.standard_normal((3, 3))
        if kind == 1:
            Uq, _, Vt = np.linalg.svd(A); A = Uq @ Vt
        t, c, cm, cf = rng.standard_normal(3) * 10, rng.standard_normal(3) * 5, rng.standard_normal(3) * 50, rng.standard_normal(3) * 50
        M = rng.standard_normal((dof, dof)); cov = M @ M.T
        for un in (0, 1):
            res = []
            for lib, pre in ((oracle_lib, "lrro_"), (ref_lib, "lrrr_")):
                a, tt, cc, cv = A.ravel().copy(), t.copy(), c.copy(), cov.ravel().copy()
                getattr(lib.dll, pre + "test_trans_roundtrip")(C.c_int(kind), pd(a), pd(tt), pd(cc), pd(cv), C.c_double(37.5), C.c_double(41.25), pd(cm), pd(cf), C.c_int(un))
                res.append((a, tt, cc, cv))
            for x, y in zip(*res):
                assert rel(x, y) < 1e-12
        res = []
        for lib, pre in ((oracle_lib, "lrro_"), (ref_lib, "lrrr_")):
            a, tt, cc = A.ravel().copy(), t.copy(), c.copy()
            getattr(lib.dll, pre + "test_trans_inverse")(C.c_int(kind), pd(a), pd(tt), pd(cc))
            res.append((a, tt, cc))
        for x, y in zip(*res):
            assert rel(x, y) < 1e-12
    R = np.linalg.qr(rng.standard_normal((3, 3)))[0]; t = rng.standard_normal(3)
    dR = np.eye(3) + 0.01 * rng.standard_normal((3, 3)); dt = rng.standard_normal(3)
    res = []
    for lib, pre in ((oracle_lib, "lrro_"), (ref_lib, "lrrr_")):
        a, tt = R.ravel().copy(), t.copy()
        getattr(lib.dll, pre + "test_recompose")(pd(a), pd(tt), pd(dR.ravel().copy()), pd(dt.copy()))
        res.append((a, tt))
    for x, y in zip(*res):
        assert rel(x, y) < 1e-12
