"""CPU: the C-ABI library loads and exports every symbol include/nolzss_b200.h declares."""
import ctypes
import os
import re

import pytest

from nolzss_b200 import _lib as L
from nolzss_b200 import build as B

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "nolzss_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(nlz_[a-z0-9_]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    B.build()
    lib = ctypes.CDLL(B.LIB)
    names = _declared_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert set(names) == set(L.SIGNATURES), "python signature table out of sync with the header"


def test_version_and_error_string():
    lib = L.load()
    assert lib.nlz_version().decode().startswith("1.2.0")
    assert isinstance(lib.nlz_last_error(), bytes)


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CUDA device"):
        L.factorize_array(L.MODE_GENERAL, b"abracadabra")
