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


# ------------------------------------------------------------------------------------------
# CenterHead.predict: decode + post_processing + rotated NMS  (center_head.py:293-495,
# box_torch_ops.py:449-470, iou3d_nms.cpp:90-136)
# ------------------------------------------------------------------------------------------
def iou_bev_matrix(a, b):
    a, b = _f32(a), _f32(b)
    out = np.empty((a.shape[0], b.shape[0]), np.float32)
    lib().orc_iou_bev_matrix(_p(a, _f32p), a.shape[0], _p(b, _f32p), b.shape[0], _p(out, _f32p))
    return out


def nms_sorted(boxes, thresh):
    """iou3d_nms_cuda.nms_gpu on score-sorted boxes [n,7] -> (keep indices int64, min |IoU - thresh| seen)."""
    boxes = _f32(boxes)
    n = boxes.shape[0]
    keep = np.empty((max(n, 1),), np.int64)
    margin = ctypes.c_float(0)
    lib().orc_nms_sorted.restype = ctypes.c_int
    nk = lib().orc_nms_sorted(_p(boxes, _f32p), n, ctypes.c_float(thresh), _p(keep, _i64p), ctypes.byref(margin))
    return keep[:nk].copy(), float(margin.value)


def centerhead_decode(preds, out_size_factor, voxel_size, pc_range):
    """center_head.py:342-419 for one task, without double flip.  preds: dict of NHWC float32 arrays
    (reg [B,H,W,2], height [B,H,W,1], dim [B,H,W,3], rot [B,H,W,2], hm [B,H,W,C]) ->
    (boxes [B,H*W,7], hm sigmoid [B,H*W,C])."""
    f = np.float32
    hm = (f(1) / (f(1) + np.exp(-preds["hm"].astype(f)))).astype(f)
    dim = np.exp(preds["dim"].astype(f)).astype(f)
    rot = np.arctan2(preds["rot"][..., 0:1].astype(f), preds["rot"][..., 1:2].astype(f)).astype(f)
    B, H, W, C = hm.shape
    ys, xs = np.meshgrid(np.arange(H, dtype=f), np.arange(W, dtype=f), indexing="ij")
    xs = xs.reshape(1, -1, 1) + preds["reg"].reshape(B, H * W, 2)[:, :, 0:1].astype(f)
    ys = ys.reshape(1, -1, 1) + preds["reg"].reshape(B, H * W, 2)[:, :, 1:2].astype(f)
    xs = ((xs * f(out_size_factor)).astype(f) * f(voxel_size[0])).astype(f) + f(pc_range[0])
    ys = ((ys * f(out_size_factor)).astype(f) * f(voxel_size[1])).astype(f) + f(pc_range[1])
    boxes = np.concatenate([xs.astype(f), ys.astype(f), preds["height"].reshape(B, H * W, 1).astype(f),
                            dim.reshape(B, H * W, 3), rot.reshape(B, H * W, 1)], axis=2)
    return boxes.astype(f), hm.reshape(B, H * W, C)


def post_processing(boxes, hm, score_threshold, post_center_range, iou_threshold, pre_max, post_max):
    """center_head.py:450-495 + rotate_nms_pcdet for one sample: boxes [HW,7], hm [HW,C] (after sigmoid).
    Ties in the score sort are broken by the lower cell index (torch.sort gives no guarantee).
    -> dict(box3d_lidar, scores, label_preds, cells) + the NMS decision margin."""
    scores = hm.max(axis=-1)
    labels = hm.argmax(axis=-1)
    r = np.asarray(post_center_range, np.float32)
    mask = (scores > np.float32(score_threshold)) & (boxes[:, :3] >= r[:3]).all(1) & (boxes[:, :3] <= r[3:]).all(1)
    cells = np.nonzero(mask)[0]
    b, s, l = boxes[cells], scores[cells], labels[cells]
    order = np.lexsort((cells, -s.astype(np.float64)))[:pre_max]
    keep, margin = nms_sorted(b[order], iou_threshold)
    sel = order[keep][:post_max]
    return dict(box3d_lidar=b[sel], scores=s[sel], label_preds=l[sel], cells=cells[sel]), margin


# ------------------------------------------------------------------------------------------
# Second stage (two_stage.py:49-199, bird_eye_view.py:24-40, center_utils.py:93-122, roi_head.py:70-106,
# roi_head_template.py:153-183).  PINNED by tests/golden/two_stage.npz, generated with the reference's own
# modules imported through the shim (tests/golden/make_golden.py second).
# ------------------------------------------------------------------------------------------
def box_sample_points(boxes, num_point=5):
    """get_box_center: [n,7] -> [num_point, n, 2] BEV points: centre, front, back, left, right."""
    f = np.float32
    boxes = boxes.astype(f)
    c = boxes[:, :2]
    if num_point == 1:
        return c[None]
    norm = np.array([[-0.5, -0.5], [-0.5, 0.5], [0.5, 0.5], [0.5, -0.5]], f)
    corners = boxes[:, None, 3:5] * norm[None]
    cs, sn = np.cos(boxes[:, 6]).astype(f), np.sin(boxes[:, 6]).astype(f)
    x = corners[..., 0] * cs[:, None] + corners[..., 1] * sn[:, None]
    y = corners[..., 0] * (-sn[:, None]) + corners[..., 1] * cs[:, None]
    corners = np.stack([x, y], -1).astype(f) + c[:, None]
    mid = lambda i, j: ((corners[:, i] + corners[:, j]) / f(2)).astype(f)  # noqa: E731
    return np.stack([c, mid(0, 1), mid(2, 3), mid(0, 3), mid(1, 2)], 0)


def bilinear_sample(im, x, y):
    """bilinear_interpolate_torch: im [H,W,C], x/y [n] in cell units; indices clamped before the weights are formed."""
    f = np.float32
    x, y = x.astype(f), y.astype(f)
    x0 = np.floor(x).astype(np.int64); x1 = x0 + 1
    y0 = np.floor(y).astype(np.int64); y1 = y0 + 1
    x0 = np.clip(x0, 0, im.shape[1] - 1); x1 = np.clip(x1, 0, im.shape[1] - 1)
    y0 = np.clip(y0, 0, im.shape[0] - 1); y1 = np.clip(y1, 0, im.shape[0] - 1)
    wa = (x1.astype(f) - x) * (y1.astype(f) - y)
    wb = (x1.astype(f) - x) * (y - y0.astype(f))
    wc = (x - x0.astype(f)) * (y1.astype(f) - y)
    wd = (x - x0.astype(f)) * (y - y0.astype(f))
    return (im[y0, x0] * wa[:, None] + im[y1, x0] * wb[:, None] + im[y0, x1] * wc[:, None] + im[y1, x1] * wd[:, None]).astype(f)


def roi_features(bev_nhwc, boxes, pc_start, voxel_size, out_stride, num_point=5):
    """One sample: bev [H,W,C], boxes [n,7] -> [n, num_point*C]."""
    f = np.float32
    pts = box_sample_points(boxes, num_point)
    feats = []
    for q in range(pts.shape[0]):
        xs = ((pts[q][:, 0] - f(pc_start[0])) / f(voxel_size[0]) / f(out_stride)).astype(f)
        ys = ((pts[q][:, 1] - f(pc_start[1])) / f(voxel_size[1]) / f(out_stride)).astype(f)
        feats.append(bilinear_sample(bev_nhwc, xs, ys))
    return np.concatenate(feats, 1)


def _fc_chain(state, prefix, x, wide=True):
    """Conv1d(k=1) [+ BatchNorm1d eval] [+ ReLU] chain of an nn.Sequential stored under ``prefix`` in ``state``."""
    idx = sorted({int(k[len(prefix) + 1:].split(".")[0]) for k in state if k.startswith(prefix + ".")})
    acc = np.float64 if wide else np.float32
    i = 0
    while i < len(idx):
        j = idx[i]
        w = state[f"{prefix}.{j}.weight"]
        assert w.ndim == 3, "expected a Conv1d weight"
        y = (x.astype(acc) @ w[:, :, 0].T.astype(acc))
        if f"{prefix}.{j}.bias" in state:
            y = y + state[f"{prefix}.{j}.bias"].astype(acc)
        nxt = idx[i + 1] if i + 1 < len(idx) else None
        if nxt is not None and f"{prefix}.{nxt}.running_mean" in state:
            g, b = state[f"{prefix}.{nxt}.weight"], state[f"{prefix}.{nxt}.bias"]
            m, v = state[f"{prefix}.{nxt}.running_mean"], state[f"{prefix}.{nxt}.running_var"]
            y = (y - m) / np.sqrt(v.astype(acc) + 1e-5) * g + b
            y = np.maximum(y, 0)                       # every BN of these heads is followed by ReLU
            i += 2
        else:
            i += 1
        x = y.astype(np.float32)
    return x


def roi_head_forward(state, feats):
    """RoIHead.forward (eval): [n, 2560] -> (rcnn_cls [n,1], rcnn_reg [n,7]); ``state`` = RoIHead state dict (numpy)."""
    shared = _fc_chain(state, "shared_fc_layer", feats)
    return _fc_chain(state, "cls_layers", shared), _fc_chain(state, "reg_layers", shared)


def roi_refine(rois, roi_scores, rcnn_cls, rcnn_reg):
    """generate_predicted_boxes + post_process: -> (boxes [n,7], scores [n])."""
    f = np.float32
    rois, reg = rois.astype(f), rcnn_reg.astype(f)
    cs, sn = np.cos(rois[:, 6]).astype(f), np.sin(rois[:, 6]).astype(f)
    out = np.empty_like(rois)
    out[:, 0] = reg[:, 0] * cs + reg[:, 1] * sn + rois[:, 0]
    out[:, 1] = reg[:, 0] * (-sn) + reg[:, 1] * cs + rois[:, 1]
    out[:, 2] = reg[:, 2] + rois[:, 2]
    out[:, 3:7] = reg[:, 3:7] + rois[:, 3:7]
    sg = (f(1) / (f(1) + np.exp(-rcnn_cls.reshape(-1).astype(f)))).astype(f)
    return out, np.sqrt(sg * roi_scores.astype(f)).astype(f)


# ------------------------------------------------------------------------------------------
# Loss values (centernet_loss.py:6-54, trainer.py:38-76,783-789); maps are NCHW numpy arrays.
# PINNED by tests/golden/losses.npz (the reference's own FastFocalLoss / RegLoss classes through the shim).
# ------------------------------------------------------------------------------------------
def _gather_map(m, ind):
    B, C, H, W = m.shape
    f = m.transpose(0, 2, 3, 1).reshape(B, H * W, C)
    return np.take_along_axis(f, ind[:, :, None].astype(np.int64), axis=1)          # [B,M,C]


def fast_focal_loss(out, target, ind, mask, cat):
    out, target = out.astype(np.float32), target.astype(np.float32)
    m = mask.astype(np.float32)
    neg = (np.log(1 - out) * out ** 2 * (1 - target) ** 4).astype(np.float64).sum()
    pp = np.take_along_axis(_gather_map(out, ind), cat[:, :, None].astype(np.int64), axis=2)[:, :, 0]
    pos = (np.log(pp) * (1 - pp) ** 2 * m).astype(np.float64).sum()
    n = m.sum()
    return np.float32(-neg if n == 0 else -(pos + neg) / n)


def reg_loss(output, mask, ind, target, squared=False):
    """RegLoss (L1) / distill_reg_loss (squared, ``target`` an NCHW map)."""
    pred = _gather_map(output.astype(np.float32), ind)
    tgt = _gather_map(target.astype(np.float32), ind) if target.ndim == 4 else target.astype(np.float32)
    m = mask.astype(np.float32)[:, :, None]
    e = pred * m - tgt * m
    e = e * e if squared else np.abs(e)
    return (e.astype(np.float64).sum(axis=(0, 1)) / (m.sum() + 1e-4)).astype(np.float32)


def sparse2dense_loss(F_S_a, F_D_a, F_S_b, F_D_b):
    def terms(s, d):
        q = (s.astype(np.float32) - d.astype(np.float32)).astype(np.float64) ** 2
        p = d > 0
        return q[p].mean(), q[~p].mean()
    pa, na = terms(F_S_a, F_D_a)
    pb, nb = terms(F_S_b, F_D_b)
    return np.float32(10 * pa + 20 * na + 5 * pb + 20 * nb)
