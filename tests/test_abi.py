"""CPU: the C-ABI library loads without a GPU and exports every symbol include/b200ocr.h declares."""
import ctypes
import os
import re

import pytest

from onnxocr_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "b200ocr.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(b200ocr_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(lib_built):
    names = declared_symbols()
    assert len(names) >= 20
    L = ctypes.CDLL(lib_built)
    for n in names:
        assert hasattr(L, n), "missing export: " + n


def test_binding_covers_every_declared_symbol(lib_built):
    L = _lib.lib()
    for n in declared_symbols():
        getattr(L, n)


def test_no_cpu_fallback_without_device(lib_built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(RuntimeError, match="no CUDA device|no CPU fallback"):
        _lib.Engine(0)


def test_version_string(lib_built):
    assert b"sm_100a" in _lib.lib().b200ocr_version()


def test_det_params_struct_layout(lib_built):
    p = _lib.DetParams()
    _lib.lib().b200ocr_det_params_default(ctypes.byref(p))
    assert (p.limit_side_len, p.limit_type, p.max_candidates, p.min_size, p.use_dilation) == (960.0, 0, 1000, 3, 0)
    assert abs(p.thresh - 0.3) < 1e-7 and p.box_thresh == 0.6 and p.unclip_ratio == 1.5
    assert p.score_mode == 0 and ctypes.sizeof(p) == 48
    assert _lib.det_params(score_mode="slow").score_mode == 1
    with pytest.raises(ValueError):
        _lib.det_params(score_mode="medium")
