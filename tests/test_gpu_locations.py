"""The driver's candidate loop on the device (lrr_batch_run_locations; rrl/itkLocationRegistration.cxx:2230-2332, 1816-1871) against
the CPU oracle's restatement of it: ranking angles and rank of every candidate, the registrations, the eight decision measurements
(incl. the sampled transfer-error covariance, forward and backward, and estimate_LS_backward(false)), and the early exit at a
location's first verified candidate."""
import numpy as np
import pytest

from lrr_b200 import synth

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


@pytest.mark.parametrize("until_verified", [False, True])
def test_candidate_loop_matches_the_oracle(gpu_lib, oracle_lib, until_verified):
    pair = synth.make_pair(4, 40000, 40000, size=120.0)
    ncand = 4
    jobs = synth.make_jobs(pair, 5, ncand, seed=6)
    res = []
    for lib in (gpu_lib, oracle_lib):
        ctx = lib.context()
        fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
        res.append(ctx.batch_run_locations(mv, fx, jobs, ncand, 70, 9, until_verified=until_verified, voxel_spacing=1.0))
    (go, gd), (oo, od) = res
    for k in ("rank", "ran", "verified", "measured"):
        assert np.array_equal(gd[k], od[k]), (k, gd[k], od[k])
    for k in ("status", "converged", "oscillated", "iterations", "n_moving_inside", "n_fixed_inside"):
        assert np.array_equal(go[k], oo[k]), (k, go[k], oo[k])
    ok = oo["status"] == 0
    assert np.allclose(gd["rank_angles"][ok], od["rank_angles"][ok], rtol=1e-9, atol=1e-9)
    assert np.allclose(gd["rank_angle_average"][ok], od["rank_angle_average"][ok], rtol=1e-9, atol=1e-9)
    ran = od["ran"] == 1
    assert ran.sum() >= 5 and (not until_verified or ran.sum() < ok.sum()), (int(ran.sum()), int(ok.sum()))
    for k in ("A", "t", "cov"):
        for a, b in zip(go[k][ran], oo[k][ran]):
            assert rel(a, b) < 1e-5, (k, rel(a, b))
    m = od["measured"] == 1
    assert m.sum() >= 3
    for k in ("measurements", "weighted_error", "max_trace", "sheet_angles", "tube_angles"):
        assert np.allclose(gd[k][m], od[k][m], rtol=1e-7, atol=1e-9), (k, gd[k][m], od[k][m])
    # the best-ranked candidate of a location is usually the near-correct initialisation (candidate 0 of make_jobs)
    first = [int(np.argmin(np.where(od["rank"][l * ncand:(l + 1) * ncand] >= 0, od["rank"][l * ncand:(l + 1) * ncand], 99))) for l in range(5)]
    assert first.count(0) >= 3, first
