"""CPU restatement (plain torch.nn.functional, fp32) of the PointPillars + S2D student's eval forward.
TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

  pfn_forward          det3d/models/readers/pillar_encoder.py:41-56 (PFNLayer), :114-154 (PillarFeatureNet.forward)
  scatter_s2d_forward  det3d/models/readers/pillar_encoder.py:337-394 (PointPillarsScatter_S2D.forward, eval)
  rpn_forward          det3d/models/necks/rpn.py:153-162 (RPN.forward: outer F.relu after every block)

PINNED against the reference's own modules imported through the shim (tests/golden/make_golden.py pp ->
tests/golden/pillars_s2d.npz).  ``state`` maps reference state-dict keys to tensors / numpy arrays.
"""
import numpy as np
import torch
import torch.nn.functional as F

from .neck_head import _bn, _conv, _rpn_block, _t, _tconv


def pfn_forward(state, features, num_voxels, coors, voxel_size, pc_range, eps=1e-3):
    """features [M,P,5] zero padded, num_voxels [M], coors [M,4] (b,z,y,x) -> [M,64]."""
    features = torch.as_tensor(features, dtype=torch.float32)
    num = torch.as_tensor(num_voxels)
    coors = torch.as_tensor(coors)
    vx, vy = voxel_size[0], voxel_size[1]
    x_off, y_off = vx / 2 + pc_range[0], vy / 2 + pc_range[1]
    mean = features[:, :, :3].sum(dim=1, keepdim=True) / num.type_as(features).view(-1, 1, 1)
    f_cluster = features[:, :, :3] - mean
    f_center = torch.zeros_like(features[:, :, :2])
    f_center[:, :, 0] = features[:, :, 0] - (coors[:, 3].to(features.dtype).unsqueeze(1) * vx + x_off)
    f_center[:, :, 1] = features[:, :, 1] - (coors[:, 2].to(features.dtype).unsqueeze(1) * vy + y_off)
    x = torch.cat([features, f_cluster, f_center], dim=-1)
    mask = (torch.arange(x.shape[1]).view(1, -1) < num.view(-1, 1)).unsqueeze(-1).type_as(x)
    x = x * mask
    n_layers = len({k.split(".")[1] for k in state if k.startswith("pfn_layers.")})
    for i in range(n_layers):
        p = f"pfn_layers.{i}"
        y = F.linear(x, _t(state, p + ".linear.weight"))
        y = F.batch_norm(y.permute(0, 2, 1), _t(state, p + ".norm.running_mean"), _t(state, p + ".norm.running_var"),
                         _t(state, p + ".norm.weight"), _t(state, p + ".norm.bias"), False, 0.0, eps).permute(0, 2, 1)
        y = F.relu(y)
        y_max = y.max(dim=1, keepdim=True)[0]
        x = y_max if i == n_layers - 1 else torch.cat([y, y_max.repeat(1, x.shape[1], 1)], dim=2)
    return x.squeeze(1)


def scatter_s2d_forward(state, voxel_features, coords, batch_size, nx, ny):
    """-> (F_S_a, F_S_b) NCHW [B,64,ny,nx] (eval: the PCR generator is skipped)."""
    vf = torch.as_tensor(voxel_features, dtype=torch.float32)
    coords = torch.as_tensor(coords).long()
    C = vf.shape[1]
    canvas = torch.zeros((batch_size, C, ny * nx), dtype=torch.float32)
    for b in range(batch_size):
        m = coords[:, 0] == b
        canvas[b][:, coords[m, 2] * nx + coords[m, 3]] = vf[m].t()
    canvas = canvas.view(batch_size, C, ny, nx)
    e, g = 1e-5, F.gelu
    a = F.max_pool2d(canvas, 2, 2)
    a = g(_bn(state, _conv(state, a, "encoder_1.1"), "encoder_1.2", e))
    a = g(_bn(state, _conv(state, a, "encoder_1.4", 2, 0), "encoder_1.5", e))
    y_1 = g(_bn(state, _conv(state, a, "encoder_1.7"), "encoder_1.8", e))
    a = g(_bn(state, _conv(state, y_1, "encoder_2.0", 2, 1), "encoder_2.1", e))
    y_2 = g(_bn(state, _conv(state, a, "encoder_2.3", 1, 1), "encoder_2.4", e))

    def convnext(att, p):
        t = F.conv2d(att, _t(state, p + ".0.weight"), _t(state, p + ".0.bias"), 1, 3, 1, att.shape[1])
        t = F.layer_norm(t, tuple(t.shape[1:]), _t(state, p + ".1.weight"), _t(state, p + ".1.bias"), 1e-6)
        return _conv(state, g(_conv(state, t, p + ".2")), p + ".4")
    att = convnext(y_2, "convnext_block_1") + y_2
    att = convnext(att, "convnext_block_2") + att
    att = convnext(att, "convnext_block_3") + att
    d1 = F.interpolate(g(_bn(state, _conv(state, att, "decoder_1.0", 1, 1), "decoder_1.1", e)), size=(117, 117))
    y_3 = torch.cat([d1, y_1], 1)
    t = g(_bn(state, _conv(state, y_3, "decoder_2.0", 1, 1), "decoder_2.1", e))
    t = g(_bn(state, _tconv(state, t, "decoder_2.3", 2, 1), "decoder_2.4", e))
    t = g(_bn(state, _conv(state, t, "decoder_2.6"), "decoder_2.7", e))
    F_S_b = F.interpolate(t, scale_factor=2)
    F_S_a = g(_bn(state, _conv(state, F_S_b, "fusion_dense.0"), "fusion_dense.1", e)) + \
        g(_bn(state, _conv(state, canvas, "fusion_sparse.0"), "fusion_sparse.1", e))
    return F_S_a, F_S_b


def rpn_forward(state, x, layer_nums, ds_layer_strides, us_layer_strides, eps=1e-3):
    ups = []
    start = len(layer_nums) - len(us_layer_strides)
    for i in range(len(layer_nums)):
        x = F.relu(_rpn_block(state, x, i, layer_nums[i], ds_layer_strides[i], eps))
        d = i - start
        if d >= 0:
            s = us_layer_strides[d]
            p = f"deblocks.{d}"
            u = _tconv(state, x, p + ".0", s, 0) if s > 1 else _conv(state, x, p + ".0", int(round(1 / s)), 0)
            ups.append(F.relu(_bn(state, u, p + ".1", eps)))
    return torch.cat(ups, 1) if ups else x
