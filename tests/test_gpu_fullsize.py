"""Parity at BASELINE.json's full sizes (100 000 features per volume): direct oracle comparison where the oracle still
finishes in seconds, plus size-independent properties (known-transform recovery, invariance to the feature order,
independence of the batch results from the sharding)."""
import numpy as np
import pytest

from lrr_b200 import synth
from lrr_b200.sharding import job_block

pytestmark = pytest.mark.gpu
N = 100_000


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.fixture(scope="module")
def pair100k():
    return synth.make_pair(3, N, N)


def test_neighbours_bit_exact_100k(gpu_lib, oracle_lib, pair100k):
    A, t, c = pair100k["init"]
    y = (pair100k["moving"]["xyz"].T - c) @ A.T + t
    q = np.ascontiguousarray(y.T)
    g = gpu_lib.context().upload(pair100k["fixed"]).nearest(q)
    o = oracle_lib.context().upload(pair100k["fixed"]).nearest(q)
    assert np.array_equal(g, o)


def test_symmetric_iterations_at_finest_level_100k(gpu_lib, oracle_lib, pair100k):
    ests = []
    for lib in (gpu_lib, oracle_lib):
        ctx = lib.context()
        fx, mv = ctx.upload(pair100k["fixed"]), ctx.upload(pair100k["moving"])
        e = ctx.estimator("rrl_sym", mv, fx, 12)
        e.set_transform(*pair100k["init"])
        e.set_number_points(N, 0); e.set_number_points(N, 1)
        ests.append(e)
    g, o = ests
    for it in range(2):
        g.one_iteration(it); o.one_iteration(it)
        for d in (0, 1):
            gf, gt, gw = g.get_matches(d); of, ot, ow = o.get_matches(d)
            assert np.array_equal(gt, ot), (it, d, "neighbours differ", int(np.count_nonzero(gt != ot)))
            assert rel(gw, ow) < 1e-8, (it, d, "weights", rel(gw, ow), list(g.state().scale), list(o.state().scale))
            G, O = g.get_transform(d), o.get_transform(d)
            assert rel(G["A"], O["A"]) < 1e-5 and rel(G["t"], O["t"]) < 1e-5 and rel(G["cov"], O["cov"]) < 1e-5, \
                (it, d, rel(G["A"], O["A"]), rel(G["t"], O["t"]), rel(G["cov"], O["cov"]))


def test_recovers_known_affine_100k(gpu_lib):
    """Noise-free pair with exact correspondences (moving = T_gt^-1(fixed) + 1e-4 mm): the ramped run ends at T_gt."""
    rng = np.random.default_rng(9)
    scene = synth.Scene(9)
    loc, shape, frame = scene.sample(N, 91)
    fixed = synth.make_features(loc, shape, frame, 92)
    A, t = synth.random_affine(rng, 8.0, 0.04, 8.0)
    c0 = np.full(3, 150.0)
    ml = (loc - c0 - t) @ np.linalg.inv(A).T + c0 + rng.normal(scale=1e-4, size=loc.shape)
    moving = synth.make_features(ml, shape, frame, 93)
    dA, dt = synth.random_affine(rng, 1.5, 0.0, 1.5)
    ctx = gpu_lib.context()
    est = ctx.estimator("rrl_sym", ctx.upload(moving), ctx.upload(fixed), 12)
    est.set_transform(dA @ A, c0 + t + dt, c0)
    conv, it = est.run(70)
    T = est.get_transform(0)
    err = np.abs(((ml - T["c"]) @ T["A"].T + T["t"]) - loc).max()
    assert conv and err < 5e-3, (it, err)
    assert est.state().number_points[0] == N


def test_feature_order_invariance_100k(gpu_lib, pair100k):
    """Sums are order independent up to rounding: permuting both feature sets leaves the estimate unchanged."""
    rng = np.random.default_rng(2)
    res = []
    for perm in (False, True):
        f, m = pair100k["fixed"], pair100k["moving"]
        if perm:
            f = synth.subset(f, rng.permutation(N)); m = synth.subset(m, rng.permutation(N))
        ctx = gpu_lib.context()
        e = ctx.estimator("cdcl_fwd", ctx.upload(m), ctx.upload(f), 12)
        e.set_flags(True, True); e.set_transform(*pair100k["init"]); e.set_number_points(N)
        for it in range(3):
            e.one_iteration(it)
        res.append(e.get_transform(0))
    assert rel(res[1]["A"], res[0]["A"]) < 1e-9 and rel(res[1]["t"], res[0]["t"]) < 1e-9 and rel(res[1]["cov"], res[0]["cov"]) < 1e-8


def test_batch_independent_of_sharding(gpu_lib):
    """A job's result depends on its GLOBAL index only (shuffle seed), not on the shard it is run in: block-cyclic shards
    (lrr_batch_run_ids, what bench.py --gpus N runs per rank) and contiguous runs reproduce the single launch."""
    from lrr_b200.sharding import job_shard
    pair = synth.make_pair(4, N, N, size=160.0)
    jobs = synth.make_jobs(pair, 6, 4, seed=4)
    ctx = gpu_lib.context()
    fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
    full = ctx.batch_run(mv, fx, jobs, 70, 11)
    cyc = np.zeros_like(full)
    for r in range(3):
        ids = job_shard(len(jobs), 3, r)
        cyc[ids] = ctx.batch_run(mv, fx, jobs[ids], 70, 11, ids=ids)
    parts = []
    for r in range(3):
        a, b = job_block(len(jobs), 3, r)
        parts.append(ctx.batch_run(mv, fx, jobs[a:b], 70, 11 + a))
    for cat in (cyc, np.concatenate(parts)):
        for k in ("status", "converged", "oscillated", "iterations", "n_moving_inside", "n_fixed_inside"):
            assert np.array_equal(cat[k], full[k]), k
        assert rel(cat["A"], full["A"]) < 1e-9 and rel(cat["cov"], full["cov"]) < 1e-9
    assert (full["status"] == 0).sum() >= 20 and full["converged"].sum() >= 12


def test_two_contexts_in_one_process_and_multi_device_entry(gpu_lib):
    """The ABI keeps no mutable global state: two contexts of one process (on two devices when the box has them, else both on
    device 0) run the full-size path side by side, and lrr_batch_run_multi — one host thread per context, dynamic job cursor —
    returns what the single-device call returns."""
    import torch
    ndev = torch.cuda.device_count()
    devs = [0, 1 if ndev > 1 else 0]
    pair = synth.make_pair(4, N, N, size=160.0)
    jobs = synth.make_jobs(pair, 12, 32, seed=4)
    ctxs = [gpu_lib.context(d) for d in devs]
    fxs = [c.upload(pair["fixed"]) for c in ctxs]; mvs = [c.upload(pair["moving"]) for c in ctxs]
    # the finest-level kernels (dynamic shared memory above 48 KB: per-device function attributes) on both contexts
    res = []
    for c, mv, fx in zip(ctxs, mvs, fxs):
        e = c.estimator("rrl_sym", mv, fx, 12)
        e.set_transform(*pair["init"]); e.set_number_points(N, 0); e.set_number_points(N, 1)
        for it in range(3):
            e.one_iteration(it)
        res.append(e.get_transform(0))
    assert rel(res[1]["A"], res[0]["A"]) < 1e-12 and rel(res[1]["cov"], res[0]["cov"]) < 1e-10
    single = ctxs[0].batch_run(mvs[0], fxs[0], jobs, 70, 7)
    multi = gpu_lib.batch_run_multi(ctxs, mvs, fxs, jobs, 70, 7)
    for k in ("status", "converged", "oscillated", "iterations", "n_moving_inside", "n_fixed_inside"):
        assert np.array_equal(multi[k], single[k]), k
    assert rel(multi["A"], single["A"]) < 1e-9 and rel(multi["cov"], single["cov"]) < 1e-9
