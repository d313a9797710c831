"""CPU: the C-ABI library loads and exports exactly what include/s2d_b200.h declares; argument
validation works without a GPU (no compute calls)."""
import ctypes
import os
import re

import pytest

from sparse2dense_b200 import _lib

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "s2d_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(s2d_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_bound_and_exported():
    syms = declared_symbols()
    assert len(syms) >= 13
    assert sorted(_lib.SIGNATURES) == syms, "python binding and header disagree"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in the header but not exported"


def test_version_and_error_text():
    lib = _lib.load()
    assert lib.s2d_version() >= 100
    out = (ctypes.c_int * 3)()
    # bad geometry -> S2D_ERR_INVALID with a message, no crash
    rc = lib.s2d_conv_out_shape(_lib.ints([8, 8, 8]), _lib.ints([3, 3, 3]), _lib.ints([0, 1, 1]), _lib.ints([0, 0, 0]),
                                _lib.ints([1, 1, 1]), out)
    assert rc == -1 and b"conv geometry" in lib.s2d_last_error()
    with pytest.raises(_lib.S2DError):
        _lib.check(rc, "s2d_conv_out_shape")


def test_conv_out_shape_matches_reference_shapes():
    from sparse2dense_b200 import ops
    s = (41, 1504, 1504)                                 # scn.py:159 and the shape comments in scn.py:116-149
    s = ops.conv_out_shape(s, 3, 2, 1); assert s == (21, 752, 752)
    s = ops.conv_out_shape(s, 3, 2, 1); assert s == (11, 376, 376)
    s = ops.conv_out_shape(s, 3, 2, (0, 1, 1)); assert s == (5, 188, 188)
    s = ops.conv_out_shape(s, (3, 1, 1), (2, 1, 1), 0); assert s == (2, 188, 188)


def test_workspace_queries_and_argument_checks():
    lib = _lib.load()
    assert lib.s2d_voxelize_workspace_bytes(180000, 4, 5, 150000) > 180000 * 8
    assert lib.s2d_voxelize_workspace_bytes(-1, 4, 5, 150000) == 0
    assert lib.s2d_grid_index_bytes(4, _lib.ints([41, 1504, 1504]), 600000) > 4 * 41 * 1504 * 1504 // 4
    rc = lib.s2d_voxelize(None, _lib.ints([0, 0]), 0, 99, 5, _lib.floats([0] * 6), _lib.floats([1] * 3), 5, 10, None,
                          None, None, None, 0, None, None, 0, None)
    assert rc == -1 and b"batch" in lib.s2d_last_error()
    rc = lib.s2d_spconv_fwd(None, 0, None, None, 0, 0, 16, 16, 64, None, None, None, 0, None, 0, None)
    assert rc == -1 and b"K=" in lib.s2d_last_error()


def test_ops_refuse_cpu_tensors():
    import torch
    from sparse2dense_b200 import ops
    with pytest.raises(_lib.S2DError):
        ops.voxel_mean(torch.zeros(2, 5, 5), torch.ones(2, dtype=torch.int32))
