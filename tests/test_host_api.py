"""The C-ABI library loads without a GPU, exports exactly what include/ppf_b200.h declares, and
reports errors through status codes (never exit(), unlike the reference's HANDLE_ERROR)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def header_symbols():
    src = open(os.path.join(ROOT, "include", "ppf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ppf_[a-z_0-9]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from objective_slam_b200 import _capi
    syms = header_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(_capi.lib, s), f"{s} declared in include/ppf_b200.h but not exported"
    assert sorted(_capi.EXPORTS) == syms


def test_no_torch_types_in_the_abi():
    src = open(os.path.join(ROOT, "include", "ppf_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)            # declarations only, comments stripped
    assert "torch" not in src and "at::" not in src and "std::" not in src and "&" not in src


def test_import_fails_loudly_without_the_library(tmp_path, monkeypatch):
    import importlib
    from objective_slam_b200 import _capi
    monkeypatch.setattr(_capi, "LIB_PATH", str(tmp_path / "missing.so"))
    with pytest.raises(ImportError):
        _capi._load()
    importlib.reload(_capi)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "objective_slam_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", "Makefile")):
                text = open(os.path.join(dirpath, f)).read()
                for bad in ("import oracle", "from oracle", "liboracle", "libppf_ref", "ppf_oracle", "oracle/"):
                    assert bad not in text, f"{f} references the oracle ({bad})"


def test_bad_arguments_return_status_codes():
    from objective_slam_b200 import _capi as C
    h = ctypes.c_void_p()
    pts = np.zeros((4, 3), np.float32)
    assert C.lib.ppf_scene_create(None, 3, None, 3, 4, 0, ctypes.byref(h)) == C.PPF_ERR_INVALID
    assert b"null" in C.lib.ppf_last_error().lower()
    assert C.lib.ppf_scene_create(pts.ctypes.data, 2, pts.ctypes.data, 3, 4, 0, ctypes.byref(h)) == C.PPF_ERR_INVALID
    assert C.lib.ppf_lookup_get_stats(None, None) == C.PPF_ERR_INVALID
    assert C.lib.ppf_scene_num_points(None) == 0
    C.lib.ppf_scene_destroy(None)            # destroying NULL is a no-op
    C.lib.ppf_model_destroy(None)
    C.lib.ppf_lookup_destroy(None)


def test_cuda_failure_is_an_error_code_not_a_crash():
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a box without a GPU")
    from objective_slam_b200 import _capi as C
    h = ctypes.c_void_p()
    pts = np.ones((4, 3), np.float32)
    rc = C.lib.ppf_scene_create(pts.ctypes.data, 3, pts.ctypes.data, 3, 4, 0, ctypes.byref(h))
    assert rc == C.PPF_ERR_CUDA and C.lib.ppf_last_error()
    poses = np.zeros(16, np.float32)
    cd = C.CloudDesc(pts.ctypes.data, 3, pts.ctypes.data, 3, 4)
    dd = np.array([1.0], np.float32)
    rc = C.lib.ppf_registration(ctypes.byref(cd), 1, ctypes.byref(cd), 1, dd.ctypes.data, 1, 0.4, 0, 0, 0, 0, None,
                                poses.ctypes.data, None)
    assert rc == C.PPF_ERR_CUDA
