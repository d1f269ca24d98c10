"""Parity on BASELINE.json's configurations AT SIZE: C2 (50k/50k cdcl forward), C3 as the survey defines it
(cdcl symmetric <3,6> from a rigid T0 -> rrl symmetric <3,12> seeded from it at 100k/100k, then set_features with a larger cube
and another run) and C5 (1 M per side: neighbours, two iterations, cdcl_obj_fun_cov f / g).  Everything is compared with the CPU
oracle on the same inputs: neighbour index sets with np.array_equal in every iteration, iteration counts and control state
identical, parameters and covariances within 1e-5 relative (north_star tolerance; observed ~1e-13)."""
import numpy as np
import pytest

from lrr_b200 import synth

pytestmark = pytest.mark.gpu
TOL = 1e-5


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def same_state(g, o, dirs, what):
    for d in dirs:
        gf, gt, gw = g.get_matches(d); of, ot, ow = o.get_matches(d)
        assert np.array_equal(gf, of), (what, d, "from indices")
        assert np.array_equal(gt, ot), (what, d, "neighbour indices differ", int(np.count_nonzero(gt != ot)))
        assert rel(gw, ow) < 1e-8, (what, d, "weights", rel(gw, ow))
        G, O = g.get_transform(d), o.get_transform(d)
        for k in ("A", "t", "cov"):
            assert rel(G[k], O[k]) < TOL, (what, d, k, rel(G[k], O[k]))
    if len(dirs) == 2:
        sg, so = g.state(), o.state()
        assert (sg.number_matches, sg.finest_level, sg.oscillation_count) == (so.number_matches, so.finest_level, so.oscillation_count), what
        assert list(sg.number_points) == list(so.number_points), what


def test_c2_cdcl_forward_50k(gpu_lib, oracle_lib):
    """C2: cdcl_estimation_ICP_matching_all<3,12>, 50 000 fixed / 50 000 moving shape features, normalize_matches and
    weight_by_strength on, every point in use, run() (<= 30 iterations, cdcl :413-428) — compared after every iteration."""
    n = 50_000
    pair = synth.make_pair(2, n, n)
    es = []
    for lib in (gpu_lib, oracle_lib):
        ctx = lib.context()
        e = ctx.estimator("cdcl_fwd", ctx.upload(pair["moving"]), ctx.upload(pair["fixed"]), 12)
        e.set_flags(True, True); e.set_transform(*pair["init"]); e.set_number_points(n)
        es.append(e)
    g, o = es
    it, cg = 0, False
    while not cg and it < 30:
        cg, co = g.one_iteration(it), o.one_iteration(it)
        assert cg == co, it
        same_state(g, o, (0,), ("c2", it))
        it += 1
    assert cg, "C2 did not converge in 30 iterations"
    for m in ("rms", "weighted"):
        a, b = g.metric(m, 0), o.metric(m, 0)
        assert abs(a - b) <= 1e-9 * max(1.0, abs(b)), (m, a, b)
    # the default of the class (1000 points, no ramp in the forward estimators) through run() as well
    for e in es:
        e.set_transform(*pair["init"])
    rg, ro = g.run(30), o.run(30)
    assert rg == ro, (rg, ro)
    same_state(g, o, (0,), "c2 run(30) at 1000 points")


def test_c3_rigid_then_affine_100k(gpu_lib, oracle_lib):
    """C3: stage 1 cdcl_estimation_symmetric_ICP_matching_all<3,6> from a rigid T0, stage 2
    rrl_estimation_symmetric_ICP_matching_all<3,12> seeded from its result (A = R), both over 100 000 features per volume and to
    convergence with the reference's ramp; then the driver's set_features + set_transform re-invocation with a larger cube
    (rrl/itkLocationRegistration.cxx:1635-1637) and another ramped run.  Iteration counts, ramp / oscillation state and the
    neighbour index sets of both directions are compared after EVERY iteration."""
    n = 100_000
    pair = synth.make_pair(3, n, n)
    A0, t0, c0 = pair["init"]
    U, _, Vt = np.linalg.svd(A0)
    R0 = U @ Vt
    libs = (gpu_lib, oracle_lib)
    ctxs = [lib.context() for lib in libs]
    sets = [(c.upload(pair["moving"]), c.upload(pair["fixed"])) for c in ctxs]

    def run_both(es, tag, max_iter=70):
        it, cg = 0, False
        while not cg and it < max_iter:
            cg, co = es[0].one_iteration(it), es[1].one_iteration(it)
            assert cg == co, (tag, it)
            same_state(es[0], es[1], (0, 1), (tag, it))
            it += 1
        return it, cg

    # stage 1: rigid
    s1 = [c.estimator("cdcl_sym", mv, fx, 6) for c, (mv, fx) in zip(ctxs, sets)]
    for e in s1:
        e.set_transform(R0, t0, c0, kind=1)
    it1, conv1 = run_both(s1, "stage1 cdcl_sym<3,6>", 40)
    T1 = [e.get_transform(0) for e in s1]
    assert rel(T1[0]["A"], T1[1]["A"]) < TOL and rel(T1[0]["t"], T1[1]["t"]) < TOL
    assert np.allclose(T1[0]["A"] @ T1[0]["A"].T, np.eye(3), atol=1e-9), "stage 1 must stay a rotation"
    # stage 2: affine, seeded from stage 1 (each side from its own stage-1 result, like two runs of the same program)
    s2 = [c.estimator("rrl_sym", mv, fx, 12) for c, (mv, fx) in zip(ctxs, sets)]
    for e, T in zip(s2, T1):
        e.set_transform(T["A"], T["t"], T["c"])
    it2, conv2 = run_both(s2, "stage2 rrl_sym<3,12>")
    assert conv2 and s2[0].state().number_points[0] == n, (it2, conv2)
    assert s2[0].oscillated() == s2[1].oscillated()
    # region growing = set_features with a larger cube, set_transform, run again
    centre_f = np.full(3, 150.0)
    Ag, tg, cg_ = pair["gt"]
    centre_m = (centre_f - tg) @ np.linalg.inv(Ag).T + cg_
    for half in (50.0, 80.0):
        mi = np.nonzero(np.all(np.abs(pair["moving"]["xyz"].T - centre_m) <= half, axis=1))[0].astype(np.int32)
        fi = np.nonzero(np.all(np.abs(pair["fixed"]["xyz"].T - centre_f) <= half, axis=1))[0].astype(np.int32)
        assert len(mi) >= 1000 and len(fi) >= 1000
        Ts = [e.get_transform(0) for e in s2]
        for e, T in zip(s2, Ts):
            e.set_subsets(mi, fi)
            e.set_transform(T["A"], T["t"], T["c"])
        itr, convr = run_both(s2, f"cube half {half}")
        assert convr, (half, itr)
        for m in ("rms", "weighted", "sheet_angles", "tube_angles"):
            for d in (0, 1):
                a, b = s2[0].metric(m, d), s2[1].metric(m, d)
                assert abs(a - b) <= 1e-9 * max(1.0, abs(b)), (half, m, d, a, b)
    for e in s2:
        e.estimate_ls_backward(False)
    same_state(s2[0], s2[1], (0, 1), "estimate_LS_backward after the last cube")


def test_c5_one_million(gpu_lib, oracle_lib):
    """C5: 1 000 000 features per side.  Exact neighbours of the mapped moving set (bit-exact), two iterations of the forward
    affine estimator over every point, and cdcl_obj_fun_cov<3,12>::compute f and all 78 gradient entries on the resulting matches."""
    n = 1_000_000
    pair = synth.make_pair(5, n, n)
    A, t, c = pair["init"]
    q = np.ascontiguousarray(((pair["moving"]["xyz"].T - c) @ A.T + t).T)
    gctx, octx = gpu_lib.context(), oracle_lib.context()
    gfx, ofx = gctx.upload(pair["fixed"]), octx.upload(pair["fixed"])
    assert np.array_equal(gfx.nearest(q), ofx.nearest(q))
    es = []
    for ctx, fx in ((gctx, gfx), (octx, ofx)):
        e = ctx.estimator("cdcl_fwd", ctx.upload(pair["moving"]), fx, 12)
        e.set_flags(True, True); e.set_transform(A, t, c); e.set_number_points(n)
        es.append(e)
    g, o = es
    for it in range(2):
        assert g.one_iteration(it) == o.one_iteration(it)
        same_state(g, o, (0,), ("c5", it))
    cov = o.get_transform(0)["cov"]
    U = np.linalg.cholesky(cov).T
    x = np.concatenate([U[i, i:] for i in range(12)])
    fg_, gg = g.objcov_eval(x); fo, go = o.objcov_eval(x)
    assert abs(fg_ - fo) <= 1e-9 * abs(fo), (fg_, fo)
    assert rel(gg, go) < 1e-8, rel(gg, go)


def test_async_upload_of_directions(gpu_lib):
    """lrr_ctx_set_async_upload: the direction arrays arrive while the registration runs; the angle metrics (their only readers) wait for
    them on the device and give bit for bit what the synchronous upload gives"""
    import torch
    pair = synth.make_pair(41, 30000, 30000)
    res = []
    for mode in (False, True):
        ctx = gpu_lib.context()
        ctx.set_async_upload(mode)
        sets = []
        for side in ("fixed", "moving"):
            f = dict(pair[side]); f["C6"] = None
            if mode:   # pinned host memory: the copy is truly asynchronous
                keep = torch.from_numpy(np.ascontiguousarray(f["dirs"])).pin_memory(); f["dirs"] = keep.numpy(); f["_keep"] = keep
            sets.append((ctx.upload({k: v for k, v in f.items() if k != "_keep"}), f))
        (fx, _), (mv, _) = sets
        est = ctx.estimator("rrl_sym", mv, fx, 12)
        est.set_transform(*pair["init"])
        est.find_closest()
        a0 = [est.metric(m, d) for m in ("sheet_angles", "tube_angles") for d in (0, 1)]     # straight after the upload
        for it in range(4):
            est.one_iteration(it)
        a1 = [est.metric(m, d) for m in ("sheet_angles", "tube_angles") for d in (0, 1)]
        ctx.synchronize()
        res.append((a0, a1, est.get_transform(0)["A"].copy()))
        est.close(); fx.close(); mv.close()
    assert res[0][0] == res[1][0]      # same matches, same directions: bit for bit
    # (after iterations the sums of the work-queue passes are added in queue order, which differs from run to run by rounding)
    assert np.allclose(res[0][1], res[1][1], rtol=1e-11, atol=0) and np.allclose(res[0][2], res[1][2], rtol=1e-11, atol=1e-13)


def test_context_options_switch_paths_without_changing_results(gpu_lib):
    """lrr_ctx_set_option: certificates / tile / speculative / pipeline / tile_sparse / first_sparse select launch plans, never results (index sets
    bit-exact, transforms to rounding)"""
    pair = synth.make_pair(43, 20000, 20000)
    ref = None
    for opts in ({}, {"certificates": 0}, {"tile": 0}, {"speculative": 2}, {"pipeline": 0}, {"tile_sparse": 0}, {"tile_sparse": 32}, {"norm_skip": 0, "flat_scan": 0}, {"first_sparse": 0}):
        ctx = gpu_lib.context()
        for k, v in opts.items():
            ctx.set_option(k, v)
        fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
        est = ctx.estimator("rrl_sym", mv, fx, 12)
        est.set_transform(*pair["init"])
        it, conv = 0, False
        while not conv and it < 40:
            conv = est.one_iteration(it); it += 1
        for k in range(3):
            est.one_iteration(it + k)
        out = (it, est.get_matches(0)[1].copy(), est.get_matches(1)[1].copy(), est.get_transform(0)["A"].copy(), est.get_transform(0)["cov"].copy())
        if ref is None:
            ref = out
        else:
            assert out[0] == ref[0], opts
            assert np.array_equal(out[1], ref[1]) and np.array_equal(out[2], ref[2]), opts
            assert np.allclose(out[3], ref[3], rtol=1e-10, atol=1e-13) and np.allclose(out[4], ref[4], rtol=1e-8, atol=1e-18), opts
        est.close(); fx.close(); mv.close()
    import pytest as _pt
    with _pt.raises(Exception):
        gpu_lib.context().set_option("no_such_switch", 1)
