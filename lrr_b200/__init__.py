"""lrr_b200 — B200-native (sm_100a) backend for the CDCL robust-ICP hot path of msofka/LRR.

The product is the C-ABI shared library ``lrr_b200/liblrr_b200.so`` (include/lrr_gpu.h) built from
``lrr_b200/csrc``; this Python package is only the ctypes binding used by tests and bench.  There is no CPU
fallback: loading fails loudly when the library has not been built, and ``lrr_ctx_create`` fails with
LRR_ERR_NO_DEVICE when no CUDA device is present.
"""
import os

from .capi import Lib, Context, FeatureSet, Estimator, LrrError  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liblrr_b200.so")
_lib = None


def lib() -> Lib:
    """The product library (hand-written CUDA kernels behind the C ABI).  Raises if it is not built."""
    global _lib
    if _lib is None:
        _lib = Lib(os.environ.get("LRR_B200_LIB", LIB_PATH), "lrr_")   # override only for A/B builds of the same product
    return _lib
