"""The C-ABI shared library: loads without a GPU, exports every symbol include/bhr.h declares,
and fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import numpy as np
import pytest

from util import ROOT


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "bhr.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(bhr_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_exported_and_bound():
    from black_hole_renderer_b200 import _lib
    lib = _lib.load()
    names = _declared_symbols()
    assert len(names) >= 28
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/bhr.h but not exported by libbhr.so"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes prototype"
    for n in _lib.SIGNATURES:
        assert n in names, f"{n} bound but not declared in include/bhr.h"
    assert lib.bhr_version() >= 100


def test_struct_layouts_match_header():
    from black_hole_renderer_b200 import _lib
    assert ctypes.sizeof(_lib.BhrConfig) == 12 * 4
    assert ctypes.sizeof(_lib.BhrCamera) == 16 * 4
    assert ctypes.sizeof(_lib.BhrEntity) == 16 + 8 + 8 + 64


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from black_hole_renderer_b200 import Renderer, _lib
    with pytest.raises(_lib.BhrError):
        Renderer(16, 8, np.zeros((8, 16, 3), np.float32), np.zeros((16, 64, 4), np.float32))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "black_hole_renderer_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "liboracle" not in text, f
    assert "oracle" not in open(os.path.join(ROOT, "render.py")).read()
