"""Verification of a candidate by the trained classifier (rrl/itkLocationRegistration.cxx:2087-2109, 2508-2528 through
libsvm/svm_interface.cpp:18-102): lrr_svm_* of liblrr_io.so against the LIBSVM sources the reference tree carries, compiled unmodified
in oracle/_ref (oracle/lrr_ref_svm.cpp).  Live where that library exists; tests/golden/ref_svm.npz (its outputs on committed model
files, generator: this file run as a script) everywhere.  Host only.  Gate: labels identical, probabilities 1e-12."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from lrr_b200 import io as lio  # noqa: E402

REF_SO = os.path.join(ROOT, "oracle", "_ref", "liblrr_ref.so")
GOLD = os.path.join(ROOT, "tests", "golden", "ref_svm.npz")
FMIN = np.array([0.823246, 0.17355, 13.5656, 14.8583, 0.776814, 0.124, 10.3695, 13.049])
FMAX = np.array([16.1025, 139.276, 62.7776, 63.5239, 1.64664, 60.897, 59.8739, 60.8141])
KERNELS = ("linear", "polynomial", "rbf", "sigmoid")


def model_text(seed, nr_class, kernel, prob, base=1, nfeat=8):
    """A syntactically complete LIBSVM model with random support vectors (sparse: some components left out)."""
    rng = np.random.default_rng(seed)
    nsv = rng.integers(3, 9, size=nr_class)
    lines = ["svm_type c_svc", f"kernel_type {kernel}"]
    if kernel == "polynomial":
        lines.append("degree 3")
    if kernel != "linear":
        lines.append(f"gamma {rng.uniform(0.05, 0.4):.17g}")
    if kernel in ("polynomial", "sigmoid"):
        lines.append(f"coef0 {rng.uniform(-0.5, 0.5):.17g}")
    npair = nr_class * (nr_class - 1) // 2
    lines += [f"nr_class {nr_class}", f"total_sv {int(nsv.sum())}", "rho " + " ".join(f"{v:.17g}" for v in rng.normal(0, 0.3, npair)),
              "label " + " ".join(str(v) for v in ([1, -1] if nr_class == 2 else list(range(1, nr_class + 1))))]
    if prob:
        lines += ["probA " + " ".join(f"{v:.17g}" for v in rng.uniform(-3.0, -0.5, npair)), "probB " + " ".join(f"{v:.17g}" for v in rng.normal(0, 0.3, npair))]
    lines += ["nr_sv " + " ".join(str(int(v)) for v in nsv), "SV"]
    for _ in range(int(nsv.sum())):
        coef = " ".join(f"{v:.17g}" for v in rng.normal(0, 1.0, nr_class - 1))
        keep = rng.uniform(size=nfeat) < 0.85
        vec = " ".join(f"{base + i}:{rng.uniform(-1, 1):.17g}" for i in range(nfeat) if keep[i])
        lines.append(f"{coef} {vec} ")
    return "\n".join(lines) + "\n"


def cases():
    out = []
    for k, kernel in enumerate(KERNELS):
        out.append((10 + k, 2, kernel, True))
        out.append((20 + k, 2, kernel, False))
    out += [(31, 3, "rbf", True), (32, 4, "linear", True), (33, 3, "polynomial", False)]
    return out


def measurements(seed, n=12):
    rng = np.random.default_rng(1000 + seed)
    return FMIN + (FMAX - FMIN) * rng.uniform(-0.1, 1.1, size=(n, 8))


def ref_predict(ref, path, x, base, prob_flag):
    lab = C.c_double(0); prob = np.zeros(8); nc = C.c_int(0)
    dp = C.POINTER(C.c_double)
    rc = ref.lrrr_test_svm_predict(path.encode(), x.ctypes.data_as(dp), 8, FMIN.ctypes.data_as(dp), FMAX.ctypes.data_as(dp), base, int(prob_flag),
                                   C.byref(lab), prob.ctypes.data_as(dp), C.byref(nc))
    assert rc == 0
    return lab.value, prob[: nc.value].copy()


def run_case(tmp, case, predict):
    seed, nc, kernel, prob = case
    out = {}
    for base in (0, 1):   # 0: node indices as the driver sets them (:2099), 1: the LIBSVM convention
        path = os.path.join(tmp, f"m_{seed}_{base}.model")
        open(path, "w").write(model_text(seed, nc, kernel, prob))
        labs, probs = [], []
        for x in measurements(seed):
            lab, p = predict(path, np.ascontiguousarray(x), base, prob)
            labs.append(lab); probs.append(p if p is not None and len(p) else np.zeros(nc))
        out[f"s{seed}_b{base}_label"] = np.array(labs); out[f"s{seed}_b{base}_prob"] = np.array(probs)
    return out


def ours(path, x, base, prob):
    m = lio.SvmModel(path)
    try:
        return m.predict(x, FMIN, FMAX, index_base=base, probability=prob)
    finally:
        m.close()


def compare(got, want):
    for k in want:
        if k.endswith("_label"):
            assert np.array_equal(got[k], want[k]), k
        else:
            assert np.allclose(got[k], want[k], rtol=1e-12, atol=1e-14), (k, np.abs(got[k] - want[k]).max())


@pytest.mark.parametrize("case", cases())
def test_matches_reference_golden(tmp_path, case):
    gold = np.load(GOLD)
    got = run_case(str(tmp_path), case, ours)
    compare(got, {k: gold[k] for k in got})


@pytest.mark.skipif(not os.path.exists(REF_SO), reason="oracle/_ref/liblrr_ref.so not built (needs /root/reference)")
@pytest.mark.parametrize("case", cases())
def test_matches_compiled_libsvm(tmp_path, case):
    ref = C.CDLL(REF_SO)
    want = run_case(str(tmp_path), case, lambda p, x, b, pr: ref_predict(ref, p, x, b, pr))
    got = run_case(str(tmp_path), case, ours)
    compare(got, want)


def test_driver_verification(tmp_path):
    """lrr_svm_verify = the driver's acceptance test: its scaling constants, node indices from 0, probability of the first class > 0.5"""
    path = os.path.join(str(tmp_path), "two.model")
    open(path, "w").write(model_text(10 + 2, 2, "rbf", True))
    m = lio.SvmModel(path)
    assert m.nr_class == 2 and m.has_probability and list(m.labels) == [1, -1]
    seen = set()
    for x in measurements(77, 40):
        v, p, lab = m.verify(x)
        lab2, pr = m.predict(x, FMIN, FMAX, index_base=0, probability=True)
        assert p == pr[0] and lab == lab2 and v == (pr[0] > 0.5)
        assert abs(pr.sum() - 1.0) < 1e-12
        seen.add(v)
    assert seen == {True, False}
    m.close()
    # a model without probability information, or with more than two classes, is refused (the driver exits: svm_interface.cpp:82-86)
    for case in ((20, 2, "linear", False), (31, 3, "rbf", True)):
        open(path, "w").write(model_text(*case))
        m = lio.SvmModel(path)
        with pytest.raises(lio.IoError):
            m.verify(measurements(1, 1)[0])
        m.close()
    with pytest.raises(lio.IoError):
        lio.SvmModel(os.path.join(str(tmp_path), "missing.model"))
    open(path, "w").write("svm_type c_svc\nkernel_type rbf\nbogus 1\n")
    with pytest.raises(lio.IoError):
        lio.SvmModel(path)


if __name__ == "__main__":   # regenerate tests/golden/ref_svm.npz from the compiled reference
    import tempfile
    ref = C.CDLL(REF_SO)
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for case in cases():
            out.update(run_case(tmp, case, lambda p, x, b, pr: ref_predict(ref, p, x, b, pr)))
    np.savez_compressed(GOLD, **out)
    print("wrote", GOLD, len(out), "arrays")
