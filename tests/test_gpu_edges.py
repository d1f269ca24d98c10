"""Edge cases of the hot path on the GPU against the CPU oracle: ties and duplicates in the neighbour search, tiny and
ragged feature sets, degenerate order statistics (all keys equal / capped), subsets with repeated entries, empty sets."""
import numpy as np
import pytest

from lrr_b200 import synth
from lrr_b200.capi import LrrError

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def both(gpu_lib, oracle_lib, pair, kind, n_iter, setup=None, dof=12):
    out = []
    for lib in (gpu_lib, oracle_lib):
        ctx = lib.context()
        fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
        est = ctx.estimator(kind, mv, fx, dof)
        est.set_flags(True, True)
        est.set_transform(*pair["init"])
        if setup:
            setup(est)
        conv = [est.one_iteration(i) for i in range(n_iter)]
        out.append((est, conv))
    (g, cg), (o, co) = out
    assert cg == co
    for d in ((0, 1) if "sym" in kind else (0,)):
        gf, gt, gw = g.get_matches(d); of, ot, ow = o.get_matches(d)
        assert np.array_equal(gf, of) and np.array_equal(gt, ot)
        assert rel(gw, ow) < 1e-8
        G, O = g.get_transform(d), o.get_transform(d)
        assert rel(G["A"], O["A"]) < 1e-5 and rel(G["t"], O["t"]) < 1e-5 and rel(G["cov"], O["cov"]) < 1e-5
    return g, o


def test_duplicate_fixed_points_tie_to_lowest_index(gpu_lib, oracle_lib):
    pair = synth.make_pair(41, 3000, 2000)
    f = pair["fixed"]
    # every feature of the second half is an exact copy (location) of one in the first half: ties everywhere
    n = f["xyz"].shape[1]; h = n // 2
    f["xyz"][:, h:2 * h] = f["xyz"][:, :h]
    q = np.ascontiguousarray(f["xyz"][:, :500] + 1e-9)
    gi = gpu_lib.context().upload(f).nearest(q)
    oi = oracle_lib.context().upload(f).nearest(q)
    assert np.array_equal(gi, oi)
    exact = gpu_lib.context().upload(f).nearest(np.ascontiguousarray(f["xyz"][:, h:h + 300]))
    assert np.array_equal(exact, np.arange(300))          # the copy with the lower index wins
    both(gpu_lib, oracle_lib, pair, "rrl_sym", 4)


def test_tiny_sets_below_the_1000_point_start(gpu_lib, oracle_lib):
    """number_points_ = min(1000, |moving|): 60 / 45 features, first iteration already at the finest level."""
    pair = synth.make_pair(42, 60, 45, size=40.0)
    g, o = both(gpu_lib, oracle_lib, pair, "cdcl_sym", 5)
    assert g.state().number_points[0] == 45 and g.state().number_points[1] == 60
    both(gpu_lib, oracle_lib, pair, "rrl_fwd", 3)


def test_ragged_sizes(gpu_lib, oracle_lib):
    pair = synth.make_pair(43, 7001, 1033, size=90.0)     # very different set sizes, not multiples of anything
    both(gpu_lib, oracle_lib, pair, "rrl_sym", 6)
    both(gpu_lib, oracle_lib, pair, "cdcl_fwd", 3, setup=lambda e: e.set_number_points(1033))


def test_equal_and_capped_strengths(gpu_lib, oracle_lib):
    """All strength products equal (order statistic among identical keys) and all above the 90000 cap (cdcl :843)."""
    pair = synth.make_pair(44, 3000, 2500, size=80.0)
    for s in (7.0, 1000.0):
        pair["fixed"]["strength"][:] = s; pair["moving"]["strength"][:] = s
        both(gpu_lib, oracle_lib, pair, "rrl_sym", 4)
        both(gpu_lib, oracle_lib, pair, "cdcl_fwd", 3, setup=lambda e: e.set_number_points(2500))


def test_subset_with_repeated_entries_and_order(gpu_lib, oracle_lib):
    pair = synth.make_pair(45, 4000, 3000, size=80.0)
    rng = np.random.default_rng(1)
    mi = rng.integers(0, 3000, size=1800).astype(np.int32)      # with repetitions, unsorted
    fi = rng.permutation(4000)[:2200].astype(np.int32)

    def setup(e):
        e.set_subsets(mi, fi)
        e.set_transform(*pair["init"])
    both(gpu_lib, oracle_lib, pair, "rrl_sym", 5, setup=setup)


def test_empty_inputs(gpu_lib):
    ctx = gpu_lib.context()
    empty = {"xyz": np.zeros((3, 0)), "P6": np.zeros((6, 0)), "C6": None, "strength": np.zeros(0), "shape": np.zeros(0, np.uint8), "dirs": None}
    fs = ctx.upload(empty)
    assert fs.nearest(np.zeros((3, 5))).tolist() == [-1] * 5       # empty set: no neighbour
    pair = synth.make_pair(46, 500, 400, size=40.0)
    mv = ctx.upload(pair["moving"])
    est = ctx.estimator("cdcl_fwd", mv, fs, 12)
    est.set_transform(*pair["init"])
    with pytest.raises(LrrError):
        est.one_iteration(0)                                       # nothing to match against: reported, not crashed
    assert ctx.batch_run(mv, ctx.upload(pair["fixed"]), np.zeros(0, dtype=synth.make_jobs(pair, 1, 1, 0).dtype), 10, 0).shape == (0,)


def test_bad_arguments(gpu_lib):
    ctx = gpu_lib.context()
    pair = synth.make_pair(47, 300, 300, size=30.0)
    fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
    with pytest.raises(LrrError):
        ctx.estimator("rrl_sym", mv, fx, 6)                        # Q6: rrl is not instantiated for <3,6>
    est = ctx.estimator("cdcl_fwd", mv, fx, 12)
    with pytest.raises(LrrError):
        est.set_subsets(np.array([0, 5, 300], dtype=np.int32))     # index out of range
    with pytest.raises(LrrError):
        est.set_order(np.arange(10, dtype=np.int32))               # wrong length


@pytest.mark.parametrize("scale,noise", [(1.0, 1e-7), (1000.0, 1e-3), (1e-3, 1e-9), (1.0, 0.3)])
def test_neighbours_exact_under_scale_offset_and_tiny_distances(gpu_lib, oracle_lib, scale, noise):
    """Neighbour distances far below the coordinate magnitude, huge and tiny coordinate scales and an offset origin must
    all give the oracle's indices (this also guards any future reduced-precision pre-filter of the candidate scan: an
    fp32 pre-filter was tried in round 1, passed this test, and was dropped because it was slower)."""
    rng = np.random.default_rng(7)
    pair = synth.make_pair(48, 6000, 5000, size=120.0)
    f, m = pair["fixed"], pair["moving"]
    A, t, c = pair["init"]
    # moving := preimages of (a subset of) the fixed locations, plus noise: every query has a neighbour at ~noise
    src = rng.integers(0, 6000, size=5000)
    y = f["xyz"][:, src].T + rng.normal(scale=noise, size=(5000, 3))
    m["xyz"] = np.ascontiguousarray(((y - t) @ np.linalg.inv(A).T + c).T)
    off = np.array([7.0e3, -4.0e3, 1.0e3])
    for k in (f, m):
        k["xyz"] = np.ascontiguousarray((k["xyz"] + off[:, None]) * scale)
    init = (A, (t + off - A @ off) * scale + 0.0, c * scale)     # T'(x') = A (x' - c') + t' in the shifted, scaled frame
    init = (A, scale * (t + off) , scale * (c + off))
    out = []
    for lib in (gpu_lib, oracle_lib):
        ctx = lib.context()
        e = ctx.estimator("cdcl_fwd", ctx.upload(m), ctx.upload(f), 12)
        e.set_flags(True, False); e.set_transform(*init); e.set_number_points(5000)
        res = []
        for it in range(3):
            e.one_iteration(it)
            res.append(e.get_matches(0)[1].copy())
        out.append(res)
    for a, b in zip(*out):
        assert np.array_equal(a, b)
    if noise <= 1e-3 * scale:
        assert np.mean(out[0][0] == src) > 0.9          # the planted neighbours are found


@pytest.mark.parametrize("kind,dof", [("rrl_sym", 12), ("cdcl_fwd", 6)])
def test_neighbour_certificates_under_transform_jitter(gpu_lib, oracle_lib, kind, dof):
    """Finest-level matches are kept without a search while the neighbour certificate of a from-feature (lrr_tile.cu: proven
    position, lower bound on every other to-feature) covers the motion of its mapped position.  Jitter the transform by amounts
    from far below to far above the certificates' reach, in both orders, with duplicate fixed points in the set (ties must never
    be certified): the index sets must stay bit-exact against the oracle's fresh k-d tree search every time."""
    pair = synth.make_pair(61, 20000, 18000, size=120.0, rigid=(dof == 6))
    f = pair["fixed"]
    for k in f:   # exact duplicates of 50 fixed features at the end: lowest index wins ties
        if f[k] is not None:
            f[k] = np.ascontiguousarray(np.concatenate([f[k], f[k][..., :50]], axis=-1))
    A0, t0, c0 = pair["init"]
    rng = np.random.default_rng(3)
    ests = []
    for lib in (gpu_lib, oracle_lib):
        ctx = lib.context()
        est = ctx.estimator(kind, ctx.upload(pair["moving"]), ctx.upload(pair["fixed"]), dof)
        est.set_flags(True, True)
        ests.append(est)
    dirs = (0, 1) if "sym" in kind else (0,)
    n_kept_changes = 0
    prev = None
    for step, mag in enumerate([0.0, 1e-7, 1e-5, 1e-3, 1e-2, 0.1, 1.0, 5.0, 1e-4, 0.0, 0.3, 1e-6, 2.0, 1e-3]):
        w = rng.normal(size=3) * mag / 60.0   # rotation that moves points at 60 mm from the centre by ~mag
        K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
        R, _ = np.linalg.qr(np.eye(3) + K)
        R = R * np.sign(np.diag(R))
        A = R @ A0 if dof == 12 else R @ A0
        t = t0 + rng.normal(size=3) * mag
        res = []
        for est in ests:
            est.set_transform(A, t, c0)
            for d in dirs:
                est.set_number_points(10 ** 9, d)
            est.find_closest()
            res.append([est.get_matches(d) for d in dirs])
        for d in dirs:
            assert np.array_equal(res[0][d][0], res[1][d][0]), (step, mag, d)
            bad = np.count_nonzero(res[0][d][1] != res[1][d][1])
            assert bad == 0, (step, mag, d, bad)
        if prev is not None:
            n_kept_changes += int(np.count_nonzero(prev != res[1][0][1]))
        prev = res[1][0][1]
    assert n_kept_changes > 0   # the jitter did change neighbours (the certificates had something to get wrong)


def test_certificates_with_far_queries_and_near_ties(gpu_lib, oracle_lib):
    """The hard cases of the neighbour certificates: far queries (many fixed features at almost the same distance) and near ties
    (two fixed features mirrored about the query), where the margin between the nearest and the second nearest is tiny.  Tiny
    translations along x flip the winner, larger ones break every bound: the index sets must stay bit-exact against the
    oracle's fresh search at every step (kept by certificate, staged search and queued pass B alike)."""
    rng = np.random.default_rng(9)
    nf, nm = 6000, 4000
    # fixed: two thin slabs mirrored about x = 0 (every moving feature near x = 0 has near-tied neighbours on both sides) + a dense blob
    yz = rng.uniform(-40, 40, size=(2, nf // 3))
    slab_a = np.vstack([np.full(nf // 3, -15.0) + rng.normal(0, 1e-4, nf // 3), yz])
    slab_b = np.vstack([np.full(nf // 3, 15.0) + rng.normal(0, 1e-4, nf // 3), yz + rng.normal(0, 1e-4, yz.shape)])
    blob = rng.normal(0, 3.0, size=(3, nf - 2 * (nf // 3))) + np.array([[60.0], [0.0], [0.0]])
    floc = np.ascontiguousarray(np.hstack([slab_a, slab_b, blob]))
    mloc = np.vstack([rng.normal(0, 0.02, nm), rng.uniform(-35, 35, size=(2, nm))])
    mloc[:, : nm // 4] = rng.normal(0, 3.0, size=(3, nm // 4)) + np.array([[60.0], [0.0], [0.0]])   # ordinary near queries too
    mloc = np.ascontiguousarray(mloc)
    shape = lambda n: rng.integers(0, 3, n).astype(np.uint8)
    frame = lambda n: np.linalg.qr(rng.normal(size=(n, 3, 3)))[0]
    pair = {"fixed": synth.make_features(np.ascontiguousarray(floc.T), shape(nf), frame(nf), 1),
            "moving": synth.make_features(np.ascontiguousarray(mloc.T), shape(nm), frame(nm), 2)}
    ests = []
    for lib in (gpu_lib, oracle_lib):
        ctx = lib.context()
        est = ctx.estimator("rrl_sym", ctx.upload(pair["moving"]), ctx.upload(pair["fixed"]), 12)
        est.set_flags(True, True)
        ests.append(est)
    flips = 0
    prev = None
    for step, dx in enumerate([0.0, 1e-6, -1e-6, 3e-5, -3e-5, 1e-3, -1e-3, 2e-3, 0.05, -0.05, 1e-5, -1e-5, 0.5, 1e-4, -1e-4, 0.0]):
        res = []
        for est in ests:
            est.set_transform(np.eye(3), np.array([dx, 1e-7 * step, 0.0]), np.zeros(3))
            for d in (0, 1):
                est.set_number_points(10 ** 9, d)
            est.find_closest()
            res.append([est.get_matches(d) for d in (0, 1)])
        for d in (0, 1):
            assert np.array_equal(res[0][d][0], res[1][d][0]), (step, dx, d)
            bad = np.count_nonzero(res[0][d][1] != res[1][d][1])
            assert bad == 0, (step, dx, d, bad)
        if prev is not None:
            flips += int(np.count_nonzero(prev != res[1][0][1]))
        prev = res[1][0][1]
    assert flips > 1000   # the mirrored slabs did flip the neighbours of the near-tied queries


def test_normalisation_is_recomputed_whenever_a_match_changes(gpu_lib, oracle_lib):
    """k_norm1 / k_norm2 leave at entry while every query of the source direction is kept by its certificate (the match set, which
    is all the normalisation of an affine estimate depends on, is unchanged).  Run far past convergence (skips active), then jump
    the transform (matches change, the normalisation must be recomputed), then converge again: transform, covariance and index
    sets must follow the oracle — which recomputes everything every time — at every step, far inside the 1e-5 gate."""
    pair = synth.make_pair(77, 24000, 22000, size=150.0)
    A0, t0, c0 = pair["init"]
    ests = []
    for lib in (gpu_lib, oracle_lib):
        ctx = lib.context()
        est = ctx.estimator("rrl_sym", ctx.upload(pair["moving"]), ctx.upload(pair["fixed"]), 12)
        est.set_transform(A0, t0, c0)
        ests.append(est)
    g, o = ests

    def step(it, tol):
        cg, co = g.one_iteration(it), o.one_iteration(it)
        assert cg == co, it
        for d in (0, 1):
            assert np.array_equal(g.get_matches(d)[1], o.get_matches(d)[1]), (it, d)
        G, O = g.get_transform(0), o.get_transform(0)
        assert rel(G["A"], O["A"]) < tol and rel(G["t"], O["t"]) < tol and rel(G["cov"], O["cov"]) < tol, (it, rel(G["A"], O["A"]), rel(G["cov"], O["cov"]))

    for it in range(45):
        step(it, 1e-8)
    assert g.state().finest_level == 1
    # jump: a rotation of ~0.5 degree and 1 mm, both estimators restart from the same transform (set_transform resets the ramp)
    T = o.get_transform(0)
    w = np.array([0.004, -0.006, 0.008])
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    A1 = (np.eye(3) + K) @ T["A"]; t1 = T["t"] + np.array([1.0, -0.7, 0.4])
    for est in ests:
        est.set_transform(A1, t1, T["c"])
    for it in range(40):
        step(it, 1e-8)
