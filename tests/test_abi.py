"""CPU checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/lrr_gpu.h declares,
the oracle exports the same ABI under its own prefix, and the product refuses to run without a CUDA device."""
import ctypes as C
import os
import re

import pytest

from lrr_b200.capi import ABI_SYMBOLS, Lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    text = open(os.path.join(ROOT, "include", "lrr_gpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(lrr_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_agree():
    hs = header_symbols()
    assert sorted("lrr_" + s for s in ABI_SYMBOLS) == hs


def test_product_library_exports_every_symbol(gpu_lib):
    for s in header_symbols():
        assert hasattr(gpu_lib.dll, s), s


def test_oracle_exports_same_abi(oracle_lib):
    for s in header_symbols():
        assert hasattr(oracle_lib.dll, "lrro_" + s[len("lrr_"):]), s


def test_every_entry_point_cites_the_reference():
    text = open(os.path.join(ROOT, "include", "lrr_gpu.h")).read()
    cites = re.findall(r"\.(?:txx|cxx|h|cpp)\s*:\s*\d+|(?:cdcl|rrl)(?: sym)? \.?(?:txx|h)?:\d+|:\d+-\d+", text)
    assert len(cites) >= 25, len(cites)


def test_no_cpu_fallback(gpu_lib):
    """Without a CUDA device the product fails loudly (LRR_ERR_NO_DEVICE = 4); it never routes through the oracle."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert gpu_lib.fn("ctx_create")(0, C.byref(h)) == 4
    src = ""
    for root, _, files in os.walk(os.path.join(ROOT, "lrr_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                src += open(os.path.join(root, f)).read()
    assert "liblrr_oracle" not in src and "lrro_" not in src and "oracle/" not in src
