"""PCR (point-cloud reconstruction) branch of the student and its losses, training only:
``S2D_RPN.forward`` lines rpn.py:314-323 (``out_conv`` -> view [N,128,5,H,W] -> ``generator_1`` -> ``gen_out_4`` /
``gen_mask_4`` -> ``generator_2`` -> ``gen_mask_2`` / ``gen_out_2``) and ``KD_VoxelNet.mask_offset_loss`` with its targets
(det3d/models/detectors/voxelnet.py:171-185,194-215,229-249).

3-D feature maps are rows ``[B*H*W*D, C]`` in (b, y, x, z) order (z fastest), which is what ``view(N,128,5,H,W)`` of the
NHWC ``out_conv`` rows gives after one small transpose.  ``Conv3d(1x1x1)`` is the gather-GEMM with an identity table,
``ConvTranspose3d(4, 2, 1)`` its eight sub-voxel classes of 2x2x2 taps (tables below are index arithmetic on the regular
grid, built once per shape); BatchNorm3d / ReLU are the rows kernels of train.cu.  The losses never build the dense
``[N,5,D,H,W]`` targets nor the grid tensor (csrc/losses.cu, ``s2d_pcr_loss``)."""
import ctypes

import numpy as np
import torch

from . import _lib, ops
from . import autograd as AG
from .dense import ACT_GELU, ACT_NONE, ACT_RELU

_TABLES = {}


def _identity_table(device, n):
    key = (str(device), "id", n)
    t = _TABLES.get(key)
    if t is None:
        tbl = torch.arange(n, dtype=torch.int32, device=device).view(1, n)
        t = _TABLES[key] = AG.Table(tbl, n, n, symmetric=True)
    return t


def _tconv3d_tables(device, B, H, W, D):
    """ConvTranspose3d(k=4, s=2, p=1) from the (B,H,W,D) grid to (B,2H,2W,2D), rows in (b, y, x, z) order.
    -> (classes [(taps (kz,ky,kx), tbl i32 [8, n_in], out_rows i32 [n_in])], adjoint Table fine -> coarse, K = 64)."""
    key = (str(device), "tconv3d", B, H, W, D)
    hit = _TABLES.get(key)
    if hit is not None:
        return hit
    ar = lambda n: torch.arange(n, device=device, dtype=torch.int64)
    b, y, x, z = (t.reshape(-1) for t in torch.meshgrid(ar(B), ar(H), ar(W), ar(D), indexing="ij"))
    n_in = b.numel()
    H2, W2, D2 = 2 * H, 2 * W, 2 * D
    classes = []
    for pz in range(2):
        for py in range(2):
            for px in range(2):
                taps, cols = [], []
                for az in range(2):
                    for ay in range(2):
                        for ax in range(2):
                            # output o = 2q + p takes tap k = (p + 1) % 2 + 2a from input j = q + p - a
                            taps.append(((pz + 1) % 2 + 2 * az, (py + 1) % 2 + 2 * ay, (px + 1) % 2 + 2 * ax))
                            jz, jy, jx = z + pz - az, y + py - ay, x + px - ax
                            ok = (jz >= 0) & (jz < D) & (jy >= 0) & (jy < H) & (jx >= 0) & (jx < W)
                            idx = ((b * H + jy) * W + jx) * D + jz
                            cols.append(torch.where(ok, idx, torch.full_like(idx, -1)).int())
                rows = (((b * H2 + 2 * y + py) * W2 + 2 * x + px) * D2 + 2 * z + pz).int()
                classes.append((taps, torch.stack(cols, 0).contiguous(), rows.contiguous()))
    cols = []
    for kz in range(4):
        for ky in range(4):
            for kx in range(4):                     # adjoint Conv3d(4, 2, 1): coarse j reads fine 2j - 1 + k
                fz, fy, fx = 2 * z - 1 + kz, 2 * y - 1 + ky, 2 * x - 1 + kx
                ok = (fz >= 0) & (fz < D2) & (fy >= 0) & (fy < H2) & (fx >= 0) & (fx < W2)
                idx = ((b * H2 + fy) * W2 + fx) * D2 + fz
                cols.append(torch.where(ok, idx, torch.full_like(idx, -1)).int())
    adj = AG.Table(torch.stack(cols, 0).contiguous(), 8 * n_in, n_in)
    hit = _TABLES[key] = (classes, adj)
    return hit


def _conv1(x, conv, precision):
    """Conv3d(1x1x1) on rows: identity-table gather-GEMM (bias handled by the caller's norm_act)."""
    w = conv.weight
    kio = w.view(w.shape[0], w.shape[1]).t().unsqueeze(0)                 # [1, Cin, Cout]
    return AG.GatherConv.apply(x, kio, _identity_table(x.device, x.shape[0]), precision)


def _seq_conv_bn_relu(x, conv, bn, precision):
    return AG.norm_act(_conv1(x, conv, precision), bn, None, ACT_RELU, pre_bias=conv.bias)


def _seq_tconv_bn_relu(x, conv, bn, B, H, W, D, precision):
    assert conv.kernel_size == (4, 4, 4) and conv.stride == (2, 2, 2) and conv.padding == (1, 1, 1)
    classes, adj = _tconv3d_tables(x.device, B, H, W, D)
    y = AG.TransposedConv.apply(x, conv.weight, classes, adj, precision)
    return AG.norm_act(y, bn, None, ACT_RELU, pre_bias=conv.bias)


def _head(x, conv, precision):
    return AG.norm_act(_conv1(x, conv, precision), None, conv.bias, ACT_NONE)


def pcr_branch(neck, F_S_b, B, H, W):
    """rpn.py:314-323 on rows -> dict(off4, mask4 @ (B,2H,2W,10); off2, mask2 @ (B,4H,4W,20); dims)."""
    Dn = neck._dense
    prec = Dn.precision
    g, _, _ = Dn.conv("out_conv.0", F_S_b, B, H, W, neck.out_conv[0], neck.out_conv[1], ACT_GELU)      # [BHW, 640], ch = c*5 + d
    g = g.view(B * H * W, 128, 5).transpose(1, 2).reshape(B * H * W * 5, 128)                       # view(N,128,5,H,W) as 3-D rows
    g1 = neck.generator_1
    g = _seq_conv_bn_relu(g, g1[0], g1[1], prec)
    g = _seq_tconv_bn_relu(g, g1[3], g1[4], B, H, W, 5, prec)                                       # (B, 2H, 2W, 10) x 32
    off4, mask4 = _head(g, neck.gen_out_4[0], prec), _head(g, neck.gen_mask_4[0], prec)
    g2 = neck.generator_2
    g = _seq_conv_bn_relu(g, g2[0], g2[1], prec)
    g = _seq_tconv_bn_relu(g, g2[3], g2[4], B, 2 * H, 2 * W, 10, prec)                              # (B, 4H, 4W, 20) x 3
    mask2, off2 = _head(g, neck.gen_mask_2[0], prec), _head(g, neck.gen_out_2[0], prec)
    return dict(off4=off4, mask4=mask4, off2=off2, mask2=mask2, dims=(B, H, W))


def as_ncdhw(p, B, H, W):
    """(gen_offset_2, gen_mask_2, gen_offset_4, gen_mask_4) in the reference's [N,C,D,H,W] layout."""
    def cv(rows, s, D):
        return rows.view(B, s * H, s * W, D, rows.shape[1]).permute(0, 4, 3, 1, 2)
    return cv(p["off2"], 4, 20), cv(p["mask2"], 4, 20), cv(p["off4"], 2, 10), cv(p["mask4"], 2, 10)


class _PcrLoss(torch.autograd.Function):
    """-> float32 [2]: (BCE-with-logits mask loss, L1 offset loss) of one scale."""

    @staticmethod
    def forward(ctx, mask_logits, offset, coors, gt_feats, dims, centre9):
        B, D, H, W = dims
        lib = _lib.load()
        mask_logits, offset = mask_logits.contiguous(), offset.contiguous()
        n = B * D * H * W
        assert mask_logits.numel() == n and tuple(offset.shape) == (n, 3)
        sums = torch.empty((6,), dtype=torch.float64, device=offset.device)
        nb = lib.s2d_loss_workspace_bytes()
        ws = torch.empty((nb,), dtype=torch.uint8, device=offset.device)
        c9 = (ctypes.c_float * 9)(*centre9)
        _lib.check(lib.s2d_pcr_loss(mask_logits.data_ptr(), offset.data_ptr(), B, D, H, W, coors.data_ptr(),
                                    gt_feats.data_ptr(), coors.shape[0], c9, sums.data_ptr(), ws.data_ptr(), nb,
                                    ops._stream()), "s2d_pcr_loss")
        ctx.save_for_backward(mask_logits, offset, coors, gt_feats, sums)
        ctx.meta = (dims, tuple(centre9))
        beta = (n - sums[3]).float() / sums[3].float()                       # count_neg / count_pos (fp32, voxelnet.py:176)
        mask_loss = (sums[0] - sums[2] + beta.double() * sums[1]) / n
        return torch.stack([mask_loss, sums[4] / sums[5]]).float()

    @staticmethod
    def backward(ctx, g):
        mask_logits, offset, coors, gt_feats, sums = ctx.saved_tensors
        (B, D, H, W), centre9 = ctx.meta
        lib = _lib.load()
        dm, do = torch.empty_like(mask_logits), torch.empty_like(offset)
        up = g.float().contiguous()
        c9 = (ctypes.c_float * 9)(*centre9)
        _lib.check(lib.s2d_pcr_loss_bwd(mask_logits.data_ptr(), offset.data_ptr(), B, D, H, W, coors.data_ptr(),
                                        gt_feats.data_ptr(), coors.shape[0], c9, sums.data_ptr(), up.data_ptr(),
                                        dm.data_ptr(), do.data_ptr(), ops._stream()), "s2d_pcr_loss_bwd")
        return dm, do, None, None, None, None


def _centre9(D, H, W):
    """voxelnet.py:231-236: xs*(150.4/W) - 75.2 + (150.4/H)/2 (the H in the x half-cell term is the reference's)."""
    f = lambda v: float(np.float32(v))
    return (f(150.4 / W), f(75.2), f((150.4 / H) / 2), f(150.4 / H), f(75.2), f((150.4 / H) / 2), f(6 / D), f(2), f((6 / D) / 2))


def pcr_losses(detector, example, p, B, H, W):
    """mask_loss = mask_loss_2 + mask_loss_4, comp_loss = offset_loss_2 + offset_loss_4 (voxelnet.py:239-243)."""
    out = []
    for suffix, s, D in (("_2", 4, 20), ("_4", 2, 10)):
        coors = example["reconstruction_coordinates" + suffix].int().contiguous()
        with torch.no_grad():
            feats = detector.reader(example["reconstruction_voxels" + suffix],
                                    example["reconstruction_num_points" + suffix]).contiguous()
        key = "2" if suffix == "_2" else "4"
        out.append(_PcrLoss.apply(p["mask" + key], p["off" + key], coors, feats, (B, D, s * H, s * W),
                                  _centre9(D, s * H, s * W)))
    return out[0][0] + out[1][0], out[0][1] + out[1][1]
