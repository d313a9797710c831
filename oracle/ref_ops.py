"""ctypes front-end of ``libs2d_oracle.so`` (numpy in / numpy out).  TEST INFRASTRUCTURE ONLY.

Every function restates one reference operation; citations are in ``s2d_oracle.c``.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

_i32p = ctypes.POINTER(ctypes.c_int)
_f32p = ctypes.POINTER(ctypes.c_float)
_i64p = ctypes.POINTER(ctypes.c_int64)


def build(force=False):
    so = os.path.join(_HERE, "libs2d_oracle.so")
    src = os.path.join(_HERE, "s2d_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libs2d_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
        _LIB.orc_rulebook_subm.restype = ctypes.c_int64
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(t) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def num_threads():
    return int(lib().orc_num_threads())


def set_num_threads(n):
    lib().orc_set_num_threads(int(n))


def grid_size(voxel_size, coors_range):
    g = np.zeros(3, np.int32)
    lib().orc_grid_size(_p(_f32(voxel_size), _f32p), _p(_f32(coors_range), _f32p), _p(g, _i32p))
    return g


def points_to_voxel(points, voxel_size, coors_range, max_points=35, reverse_index=True, max_voxels=20000):
    """Same signature and return value as the reference ``points_to_voxel`` (point_cloud_ops.py:112)."""
    assert reverse_index, "the reference hot path always uses reverse_index=True (voxel_generator.py:23-30)"
    points = _f32(points)
    n, f = points.shape
    voxels = np.zeros((max_voxels, max_points, f), np.float32)
    coors = np.zeros((max_voxels, 3), np.int32)
    num = np.zeros((max_voxels,), np.int32)
    m = lib().orc_points_to_voxel(_p(points, _f32p), n, f, _p(_f32(voxel_size), _f32p),
                                  _p(_f32(coors_range), _f32p), int(max_points), int(max_voxels), None,
                                  _p(voxels, _f32p), _p(coors, _i32p), _p(num, _i32p))
    assert m >= 0
    return voxels[:m], coors[:m], num[:m]


def voxel_mean(voxels, num_points, num_features=None):
    voxels = _f32(voxels)
    m, p, f = voxels.shape
    c = f if num_features is None else num_features
    out = np.empty((m, c), np.float32)
    lib().orc_voxel_mean(_p(voxels, _f32p), _p(_i32(num_points), _i32p), m, p, f, c, _p(out, _f32p))
    return out


def _triple(v):
    if isinstance(v, (int, np.integer)):
        return np.array([v, v, v], np.int32)
    v = np.asarray(v, np.int32)
    assert v.shape == (3,)
    return np.ascontiguousarray(v)


def rulebook_subm(coors, shape, ksize=3, dilation=1):
    """-> (tbl int32 [K,N] k-major with -1 holes, n_pairs)."""
    coors = _i32(coors)
    n = coors.shape[0]
    ks, dl = _triple(ksize), _triple(dilation)
    k = int(ks.prod())
    tbl = np.empty((k, max(n, 1)), np.int32)
    pairs = lib().orc_rulebook_subm(_p(coors, _i32p), n, _p(_triple(shape), _i32p), _p(ks, _i32p),
                                    _p(dl, _i32p), _p(tbl, _i32p), tbl.shape[1])
    return tbl[:, :n], int(pairs)


def conv_out_shape(shape, ksize, stride, pad, dilation=1):
    out = np.zeros(3, np.int32)
    lib().orc_conv_out_shape(_p(_triple(shape), _i32p), _p(_triple(ksize), _i32p), _p(_triple(stride), _i32p),
                             _p(_triple(pad), _i32p), _p(_triple(dilation), _i32p), _p(out, _i32p))
    return out


def rulebook_sparse(coors, shape, ksize, stride, pad, dilation=1):
    """-> (out_coors int32 [n_out,4] ascending, tbl int32 [K,n_out], out_shape, n_pairs)."""
    coors = _i32(coors)
    n = coors.shape[0]
    ks, st, pd, dl = _triple(ksize), _triple(stride), _triple(pad), _triple(dilation)
    k = int(ks.prod())
    cap = max(n * k, 1)
    out_coors = np.empty((cap, 4), np.int32)
    tbl = np.empty((k, cap), np.int32)
    pairs = ctypes.c_int64(0)
    n_out = lib().orc_rulebook_sparse(_p(coors, _i32p), n, _p(_triple(shape), _i32p), _p(ks, _i32p),
                                      _p(st, _i32p), _p(pd, _i32p), _p(dl, _i32p), _p(out_coors, _i32p),
                                      cap, _p(tbl, _i32p), cap, ctypes.byref(pairs))
    assert n_out >= 0
    return (out_coors[:n_out].copy(), np.ascontiguousarray(tbl[:, :n_out]),
            conv_out_shape(shape, ks, st, pd, dl), int(pairs.value))


def spconv_fwd(feats, weight, tbl, wide=False):
    """feats [N_in,Cin], weight [kD,kH,kW,Cin,Cout] (spconv layout) or [K,Cin,Cout], tbl [K,N_out]."""
    feats = _f32(feats)
    tbl = _i32(tbl)
    k, n_out = tbl.shape
    w = _f32(weight).reshape(k, feats.shape[1], -1)
    cout = w.shape[2]
    out = np.empty((n_out, cout), np.float32)
    lib().orc_spconv_fwd(_p(feats, _f32p), _p(w, _f32p), _p(tbl, _i32p), max(n_out, 1) if tbl.strides[0] == 0
                         else tbl.strides[0] // 4, n_out, feats.shape[1], cout, k, _p(out, _f32p), int(bool(wide)))
    return out


def bn_act(x, scale, shift, residual=None, relu=True):
    x = _f32(x).copy()
    lib().orc_bn_act(_p(x, _f32p), x.shape[0], x.shape[1], _p(_f32(scale), _f32p), _p(_f32(shift), _f32p),
                     _p(_f32(residual), _f32p) if residual is not None else None, int(bool(relu)))
    return x


def dense_bev(feats, coors, batch_size, spatial_shape):
    """SparseConvTensor.dense() followed by view(N, C*D, H, W) (scn.py:173-176)."""
    feats, coors = _f32(feats), _i32(coors)
    d, h, w = [int(v) for v in spatial_shape]
    c = feats.shape[1]
    bev = np.zeros((batch_size, c * d, h, w), np.float32)
    lib().orc_dense_bev(_p(feats, _f32p), _p(coors, _i32p), feats.shape[0], c, d, h, w, _p(bev, _f32p))
    return bev
