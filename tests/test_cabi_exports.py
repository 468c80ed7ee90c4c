"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU, exports every symbol
include/b2kmeans.h declares, and fails loudly (no fallback) when there is no CUDA device."""
import ctypes
import os
import re

import pytest

from spark_rapids_ml_b200 import _native

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "b2kmeans.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b2k_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert _header_symbols() == sorted(_native.EXPORTED_SYMBOLS)


def test_library_loads_and_exports_every_symbol():
    assert os.path.exists(_native.LIB_PATH), "run __graft_entry__.build() first"
    L = _native.load_library()
    for name in _header_symbols():
        assert hasattr(L, name), name
    assert L.b2k_version() == 100


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    L = _native.load_library()
    h = ctypes.c_void_p()
    rc = L.b2k_ctx_create(0, ctypes.byref(h))
    assert rc != 0
    assert b"no CPU fallback" in L.b2k_last_error(None)
    with pytest.raises(_native.B2KError):
        _native.Context(0)


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under spark_rapids_ml_b200/ may reference it."""
    pkg = os.path.join(ROOT, "spark_rapids_ml_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f
                assert "kmeans_oracle" not in txt, f
