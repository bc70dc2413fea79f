"""CPU-only: the C-ABI library builds/loads and exports every symbol include/r3d_b200.h declares;
argument validation that needs no device; the product fails loudly without a GPU or library."""
import ctypes
import os
import re

import pytest

from conftest import ROOT, pkg, sub


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "r3d_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(r3d_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    C = pkg()._capi
    if not os.path.exists(C.LIB_PATH):
        import importlib
        importlib.import_module("2dto3dreconstruction_b200.build").build()
    names = _declared_symbols()
    assert len(names) >= 25
    handle = ctypes.CDLL(C.LIB_PATH)
    for n in names:
        assert hasattr(handle, n), "library does not export %s" % n
    # and the Python binding covers exactly the header
    assert sorted(C.SIGNATURES) == names


def test_pure_host_entry_points():
    C = pkg()._capi
    L = C.lib()
    assert L.r3d_version() >= 100
    assert L.r3d_status_string(0) == b"ok"
    assert L.r3d_status_string(-2) == b"workspace too small"
    assert L.r3d_backproject_workspace_bytes(2, 480, 640) > 0
    assert L.r3d_backproject_workspace_bytes(0, 480, 640) == 0
    one = L.r3d_icp_workspace_bytes(1, 307200, 307200)
    eight = L.r3d_icp_workspace_bytes(8, 307200, 307200)
    assert 0 < one < eight < 2 * 8 * one
    assert L.r3d_register_workspace_bytes(1024, 8, 480, 640, 2) > 2 * eight
    assert L.r3d_ransac2d_workspace_bytes(100, 1100, 1100, 480, 640) > 480 * 640 * 4
    # argument validation happens before any CUDA call
    assert L.r3d_backproject(None, 0, 1, 480, 640, 1, 525.0, 525.0, 319.5, 239.5, 1.0, 0, None, None, None, 0, None) == -1
    assert L.r3d_kabsch(None, 3, 1, None, 3, 1, 10, None, None, 0, None) == -1
    assert L.r3d_ransac3d_score(None, 1, None, None, 1, 10.0, None, None) == -1


def test_no_silent_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import numpy as np
    C = pkg()._capi
    icp = sub("utils.icp")
    with pytest.raises(C.R3DError):
        icp.icp(np.zeros((10, 3)), np.zeros((10, 3)))
    with pytest.raises(C.R3DError):
        sub("utils.utils").depth_to_voxel_ld(np.zeros((4, 4), np.uint16))


def test_product_never_imports_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "2dto3dreconstruction_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", text, flags=re.M) or "/root/reference" in text:
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
