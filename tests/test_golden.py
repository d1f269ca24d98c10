"""Golden-vector checks (tests/golden/*.npz, produced by tests/golden/make_golden.py from the CPU oracle —
see its docstring: they are NOT reference outputs).  CPU: the oracle still reproduces them.  GPU: the CUDA path
reproduces them from the stored inputs alone (no oracle involved)."""
import os

import numpy as np
import pytest

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = [("cdcl_fwd", 12, False), ("rrl_fwd", 12, False), ("cdcl_sym", 12, False), ("rrl_sym", 12, False), ("cdcl_fwd", 6, True)]


def load_inputs(rigid):
    z = np.load(os.path.join(HERE, f"inputs_{'rigid' if rigid else 'affine'}.npz"))
    pair = {"fixed": {}, "moving": {}}
    for k in z.files:
        for side in ("fixed", "moving"):
            if k.startswith(side + "_"):
                pair[side][k[len(side) + 1:]] = z[k]
    pair["init"] = (z["init_A"], z["init_t"], z["init_c"])
    return pair


def check(lib, kind, dof, rigid, tol):
    gold = np.load(os.path.join(HERE, f"golden_{kind}_{dof}.npz"))
    pair = load_inputs(rigid)
    ctx = lib.context()
    fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
    est = ctx.estimator(kind, mv, fx, dof)
    est.set_flags(True, True)
    est.set_transform(*pair["init"])
    if "fwd" in kind:
        est.set_number_points(1200)
    rel = lambda a, b: np.linalg.norm(np.asarray(a) - b) / max(np.linalg.norm(b), 1e-300)
    for it in range(4):
        conv = est.one_iteration(it)
        T = est.get_transform(0)
        f, to, w = est.get_matches(0)
        assert int(conv) == int(gold[f"it{it}_conv"][0])
        assert np.array_equal(f, gold[f"it{it}_from"]) and np.array_equal(to, gold[f"it{it}_to"])       # index sets bit-exact
        assert rel(w, gold[f"it{it}_w"]) < tol
        for k in ("A", "t", "cov"):
            assert rel(T[k], gold[f"it{it}_{k}"]) < tol, (it, k, rel(T[k], gold[f"it{it}_{k}"]))
        assert abs(est.metric("rms") - gold[f"it{it}_rms"][0]) <= tol * max(1.0, gold[f"it{it}_rms"][0])
        assert abs(est.metric("weighted") - gold[f"it{it}_werr"][0]) <= tol * max(1.0, gold[f"it{it}_werr"][0])


@pytest.mark.parametrize("kind,dof,rigid", CASES)
def test_oracle_reproduces_golden(oracle_lib, kind, dof, rigid):
    check(oracle_lib, kind, dof, rigid, 1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,dof,rigid", CASES)
def test_gpu_reproduces_golden(gpu_lib, kind, dof, rigid):
    check(gpu_lib, kind, dof, rigid, 1e-8)
