import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests are skipped (not failed) on a machine without a CUDA device."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device: the product path has no CPU fallback")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle_lib():
    """CPU oracle (test infrastructure): oracle/liblrr_oracle.so, built on demand with gcc."""
    import subprocess
    from lrr_b200.capi import Lib
    so = os.path.join(ROOT, "oracle", "liblrr_oracle.so")
    src = [os.path.join(ROOT, "oracle", f) for f in ("lrr_oracle_capi.cpp", "lrr_oracle_core.hpp", "lrr_oracle_linalg.hpp", "lrr_oracle_muse.hpp")]
    if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liblrr_oracle.so"])
    return Lib(so, "lrro_")


@pytest.fixture(scope="session")
def gpu_lib():
    """The product: lrr_b200/liblrr_b200.so (hand-written CUDA behind the C ABI).  No fallback."""
    import lrr_b200
    return lrr_b200.lib()
