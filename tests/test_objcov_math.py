"""The branch-free logarithm and reciprocal of k_objcov's pipelined loop (lrr_b200/csrc/lrr_objcov.cu: obj_log, obj_rcp), restated in
numpy with the constants READ FROM THE KERNEL SOURCE, against numpy's correctly rounded log and division.  The kernel replaces
vcl_log(det) and 1 / det of cdcl/cdcl_obj_fun_cov.txx:96-118 by these forms to keep its loop free of slow-path branches; the GPU parity
tests bound the sums (f 1e-9, g 1e-8), this bounds the functions themselves (1 ulp / 2 ulp) on the CPU."""
import os
import re

import numpy as np

SRC = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lrr_b200", "csrc", "lrr_objcov.cu")


def _constants():
    text = open(SRC).read()
    body = text[text.index("double obj_log(double x)"):]
    body = body[:body.index("\n}\n")]
    vals = {}
    for name in ("ln2_hi", "ln2_lo", "L1", "L2", "L3", "L4", "L5", "L6", "L7"):
        m = re.search(r"\b%s\s*=\s*([0-9.eE+-]+)" % name, body)
        assert m, name
        vals[name] = float(m.group(1))
    m = re.search(r"\(hx \+ (0x[0-9a-fA-F]+)\) & 0x100000", body)
    assert m
    vals["split"] = int(m.group(1), 16)
    return vals


def _obj_log(x, c):
    x = np.asarray(x, dtype=np.float64)
    ix = x.view(np.int64)
    hx = ix >> 32
    lx = ix & 0xFFFFFFFF
    k = (hx >> 20) - 1023
    hx = hx & 0x000FFFFF
    i = (hx + c["split"]) & 0x100000
    hx = hx | (i ^ 0x3FF00000)
    k = k + (i >> 20)
    m = ((hx << 32) | lx).astype(np.int64).view(np.float64)
    f = m - 1.0
    s = f / (2.0 + f)
    dk = k.astype(np.float64)
    z = s * s
    w = z * z
    t1 = w * (c["L2"] + w * (c["L4"] + w * c["L6"]))
    t2 = z * (c["L1"] + w * (c["L3"] + w * (c["L5"] + w * c["L7"])))
    R = t2 + t1
    hfsq = 0.5 * f * f
    return dk * c["ln2_hi"] - ((hfsq - (s * (hfsq + R) + dk * c["ln2_lo"])) - f)


def test_obj_log_within_one_ulp_of_log():
    c = _constants()
    rng = np.random.default_rng(0)
    x = np.concatenate([np.exp(rng.uniform(-700, 700, 400_000)), rng.uniform(0.5, 2.0, 400_000), 1.0 + rng.uniform(-1e-6, 1e-6, 50_000),
                        [2.0, 0.5, np.sqrt(2.0), 1.0 / np.sqrt(2.0), 1e-300, 1e300]])
    a, b = _obj_log(x, c), np.log(x)
    assert np.all(np.abs(a - b) <= np.spacing(np.abs(b))), float(np.max(np.abs(a - b) / np.spacing(np.abs(b))))
    assert _obj_log(np.array([1.0]), c)[0] == 0.0


def test_obj_rcp_newton_from_a_20_bit_seed():
    """rcp.approx.ftz.f64 delivers the reciprocal's upper ~20 mantissa bits; three Newton steps r += r (1 - d r) reach 2 ulp without a
    fused multiply-add (the kernel's fma form is at least as accurate)."""
    text = open(SRC).read()
    body = text[text.index("double obj_rcp(double d)"):]
    body = body[:body.index("\n}\n")]
    assert body.count("r = fma(r, e, r);") == 3
    rng = np.random.default_rng(1)
    d = np.exp(rng.uniform(-200, 200, 400_000))
    seed = (1.0 / d).view(np.int64) & ~np.int64((1 << 32) - 1)   # lower 32 bits zero, as the instruction returns
    r = seed.view(np.float64)
    for _ in range(3):
        e = 1.0 - d * r
        r = r + r * e
    exact = 1.0 / d
    assert np.all(np.abs(r - exact) <= 2 * np.spacing(exact))
