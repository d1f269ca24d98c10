"""Older, ORACLE-derived regression vectors (the reference-generated ones are ref_*.npz, see make_ref_golden.py).
The reference ships no fixtures, so these are outputs of the CPU ORACLE (oracle/, pinned by tests/test_oracle_kat.py), NOT of the reference itself:
they freeze the oracle's behaviour (regression pin) and give the GPU tests an input/output set that needs no
oracle at run time.  Regenerate with:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from lrr_b200 import synth  # noqa: E402
from lrr_b200.capi import Lib  # noqa: E402

CASES = [("cdcl_fwd", 12, False), ("rrl_fwd", 12, False), ("cdcl_sym", 12, False), ("rrl_sym", 12, False), ("cdcl_fwd", 6, True)]


def run_case(lib, pair, kind, dof, iters=4):
    ctx = lib.context()
    fx, mv = ctx.upload(pair["fixed"]), ctx.upload(pair["moving"])
    est = ctx.estimator(kind, mv, fx, dof)
    est.set_flags(True, True)
    est.set_transform(*pair["init"])
    if "fwd" in kind:
        est.set_number_points(1200)
    out = {}
    for it in range(iters):
        conv = est.one_iteration(it)
        T = est.get_transform(0)
        f, to, w = est.get_matches(0)
        out[f"it{it}_A"] = T["A"]; out[f"it{it}_t"] = T["t"]; out[f"it{it}_c"] = T["c"]; out[f"it{it}_cov"] = T["cov"]
        out[f"it{it}_from"] = f; out[f"it{it}_to"] = to; out[f"it{it}_w"] = w; out[f"it{it}_conv"] = np.array([int(conv)])
        out[f"it{it}_rms"] = np.array([est.metric("rms")]); out[f"it{it}_werr"] = np.array([est.metric("weighted")])
        out[f"it{it}_scale"] = np.array([est.state().scale[0]])
    return out


def make_inputs(rigid):
    return synth.make_pair(31 + int(rigid), 1500, 1200, size=80.0, rigid=rigid)


def main():
    lib = Lib(os.path.join(ROOT, "oracle", "liblrr_oracle.so"), "lrro_")
    here = os.path.dirname(os.path.abspath(__file__))
    for rigid in (False, True):
        pair = make_inputs(rigid)
        flat = {}
        for side in ("fixed", "moving"):
            for k, v in pair[side].items():
                flat[f"{side}_{k}"] = v
        for i, n in enumerate(("A", "t", "c")):
            flat[f"init_{n}"] = np.asarray(pair["init"][i])
        np.savez_compressed(os.path.join(here, f"inputs_{'rigid' if rigid else 'affine'}.npz"), **flat)
    for kind, dof, rigid in CASES:
        out = run_case(lib, make_inputs(rigid), kind, dof)
        np.savez_compressed(os.path.join(here, f"golden_{kind}_{dof}.npz"), **out)
        print(kind, dof, "A00", out["it3_A"][0, 0])


if __name__ == "__main__":
    main()
