"""GPU parity of the dense BEV stage (S2D neck, RPN pyramid, CenterHead) through the C ABI vs torch-CPU
restatements / the oracle.  Bar: 1e-3 relative (north_star); TF32x3 is the shipped precision."""
import logging
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F
from torch import nn

from oracle import neck_head as NH
from sparse2dense_b200 import dense, ops, registry, synth

from conftest import GOLDEN
from test_oracle_dense import HEAD_CFG, NECK_CFG, bev_input, build_modules

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("k,s,p", [(3, 1, 1), (3, 2, 1), (2, 2, 0), (1, 1, 0)])
def test_conv_table_matches_unfold_geometry(k, s, p):
    B, H, W = 2, 7, 6
    tbl, Ho, Wo = dense.conv_table(torch.device("cuda"), B, H, W, k, s, p)
    t = tbl.cpu().numpy()
    assert (Ho, Wo) == ((H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1)
    for o in range(B * Ho * Wo):
        ox, oy, b = o % Wo, (o // Wo) % Ho, o // (Wo * Ho)
        for ky in range(k):
            for kx in range(k):
                iy, ix = oy * s - p + ky, ox * s - p + kx
                want = (b * H + iy) * W + ix if (0 <= iy < H and 0 <= ix < W) else -1
                assert t[ky * k + kx, o] == want


@pytest.mark.parametrize("precision,tol", [(ops.PRECISION_FP32, 1e-5), (ops.PRECISION_TF32X3, 2e-5), (ops.PRECISION_AUTO, 2e-5),
                                           (ops.PRECISION_TF32, 3e-3)])
@pytest.mark.parametrize("kind", ["c3s1", "c3s2", "c2s2", "c1", "t4", "t2"])
def test_dense_conv_ops_vs_torch(kind, precision, tol):
    torch.manual_seed(3)
    B, H, W, cin, cout = 2, 12, 10, 64, 96
    mk = dict(c3s1=lambda: nn.Conv2d(cin, cout, 3, 1, 1), c3s2=lambda: nn.Conv2d(cin, cout, 3, 2, 1),
              c2s2=lambda: nn.Conv2d(cin, cout, 2, 2), c1=lambda: nn.Conv2d(cin, cout, 1),
              t4=lambda: nn.ConvTranspose2d(cin, cout, 4, 2, 1), t2=lambda: nn.ConvTranspose2d(cin, cout, 2, 2, bias=False))
    conv = mk[kind]().eval()
    bn = nn.BatchNorm2d(cout).eval()
    bn.running_mean.normal_(0, 0.1); bn.running_var.uniform_(0.5, 1.5); bn.weight.data.uniform_(0.5, 1.5); bn.bias.data.normal_(0, 0.1)
    x = torch.randn(B, cin, H, W)
    with torch.no_grad():
        ref = F.gelu(bn(conv(x)))
    D = dense.DenseOps(precision)
    conv_c, bn_c = conv.cuda(), bn.cuda()
    rows = dense.to_rows(x.cuda())
    if kind.startswith("t"):
        y, Ho, Wo = D.tconv("m", rows, B, H, W, conv_c, bn_c, dense.ACT_GELU)
    else:
        y, Ho, Wo = D.conv("m", rows, B, H, W, conv_c, bn_c, dense.ACT_GELU)
    got = dense.to_nchw(y, B, Ho, Wo).cpu().numpy()
    assert got.shape == tuple(ref.shape)
    assert rel_err(got, ref.numpy()) < tol


def test_residual_modes_concat_slices_dwconv_layernorm():
    torch.manual_seed(4)
    B, H, W, C = 2, 9, 11, 64
    x = torch.randn(B, C, H, W)
    conv = nn.Conv2d(C, C, 1).eval()
    res = torch.randn(B, C, H, W)
    D = dense.DenseOps(ops.PRECISION_TF32X3)
    rows, rres = dense.to_rows(x.cuda()), dense.to_rows(res.cuda())
    wide = torch.zeros((B * H * W, 3 * C), device="cuda")
    with torch.no_grad():
        before = F.relu(conv(x) + res)                     # residual before the activation (ResNet style)
        after = F.gelu(conv(x)) + res                      # residual after it (F_S_a fusion, rpn.py:311)
    y1, _, _ = D.conv("a", rows, B, H, W, conv.cuda(), None, dense.ACT_RELU, residual=rres, out=wide[:, C:2 * C])
    y2, _, _ = D.conv("a", rows, B, H, W, conv.cuda(), None, dense.ACT_GELU, residual=rres, res_after_act=True)
    assert rel_err(dense.to_nchw(wide[:, C:2 * C], B, H, W).cpu(), before) < 2e-5
    assert float(wide[:, :C].abs().max()) == 0 and float(wide[:, 2 * C:].abs().max()) == 0     # slice write only
    assert rel_err(dense.to_nchw(y2, B, H, W).cpu(), after) < 2e-5
    dw = nn.Conv2d(C, C, 7, padding=3, groups=C).eval()
    ln = nn.LayerNorm([C, H, W], eps=1e-6).eval()
    ln.weight.data.uniform_(0.5, 1.5); ln.bias.data.normal_(0, 0.1)
    with torch.no_grad():
        ref = ln(dw(x))
    t = D.layernorm(D.dwconv(rows, B, H, W, dw.cuda()), B, H, W, ln.cuda())
    assert rel_err(dense.to_nchw(t, B, H, W).cpu(), ref) < 1e-5
    assert torch.equal(dense.to_nchw(rows, B, H, W).cpu(), x)                                   # transposes round-trip


def test_dense_bev_nhwc_matches_nchw_form():
    rng = np.random.default_rng(2)
    n, C, B, Dz, H, W = 700, 128, 2, 2, 20, 24
    lin = rng.choice(B * Dz * H * W, n, replace=False)
    coors = np.stack([lin // (Dz * H * W), (lin // (H * W)) % Dz, (lin // W) % H, lin % W], 1).astype(np.int32)
    feats = torch.from_numpy(rng.normal(size=(n, C)).astype(np.float32)).cuda()
    c = torch.from_numpy(coors).cuda()
    nchw = ops.dense_bev(feats, c, B, (Dz, H, W))
    rows = ops.dense_bev_rows(feats, c, B, (Dz, H, W))
    assert torch.equal(dense.to_nchw(rows, B, H, W), nchw)


@pytest.mark.parametrize("precision,tol", [(ops.PRECISION_TF32X3, 1e-3), (ops.PRECISION_AUTO, 1e-3),
                                           (ops.PRECISION_TF32, 2e-2)])
def test_s2d_rpn_and_center_head_full_size_vs_oracle_and_reference_samples(precision, tol):
    g = np.load(os.path.join(GOLDEN, "neck_head_s2d.npz"))
    neck, head = build_modules()
    ns = synth.random_module_state(neck, int(g["neck_seed"]))
    hs = synth.random_module_state(head, int(g["head_seed"]))
    neck.load_state_dict({k: torch.from_numpy(v) for k, v in ns.items()}, strict=False)
    head.load_state_dict({k: torch.from_numpy(v) for k, v in hs.items()}, strict=False)
    neck, head = neck.cuda().eval(), head.cuda().eval()
    neck.set_precision(precision); head.set_precision(precision)
    x = torch.from_numpy(bev_input(int(g["input_seed"])))
    with torch.no_grad():
        gx, n1, n2, n3, n4, gfa, gfb = neck(x.cuda())
        gh = head(gx)[0]
        ox, ofa, ofb = NH.s2d_rpn_forward(ns, x)
        oh = NH.center_head_forward(hs, ox)[0]
    assert n1 is None and n4 is None and gx.shape == (1, 512, 188, 188)
    outs = dict(x=(gx, ox), F_S_a=(gfa, ofa), F_S_b=(gfb, ofb), **{h: (gh[h], oh[h]) for h in oh})
    for name, (got, ref) in outs.items():
        got = got.cpu().numpy()
        e = rel_err(got, ref.numpy())
        print(f"{name}: rel err vs oracle {e:.2e}")
        assert e < tol, name
        samp = got.reshape(-1)[g[name + "_idx"]]                      # the reference's own outputs
        assert np.abs(samp - g[name + "_val"]).max() <= tol * float(g[name + "_absmax"]), name
    assert list(gh) == ["reg", "height", "dim", "rot", "hm"]


def test_plain_rpn_forward_vs_torch():
    """RPN.forward keeps the outer F.relu that S2D_RPN.forward drops (rpn.py:156 vs :327-331)."""
    neck = registry.build_neck(dict(type="RPN", logger=logging.getLogger("t"), **NECK_CFG)).eval()
    st = synth.random_module_state(neck, 3)
    neck.load_state_dict({k: torch.from_numpy(v) for k, v in st.items()}, strict=False)
    x = torch.from_numpy(bev_input(8))
    ups, h = [], x
    with torch.no_grad():
        for i, (n_l, s) in enumerate(zip((5, 5), (1, 2))):
            h = F.relu(NH._rpn_block(st, h, i, n_l, s, 1e-3))
            p = f"deblocks.{i}"
            u = NH._tconv(st, h, p + ".0", 2, 0) if i == 1 else NH._conv(st, h, p + ".0", 1, 0)
            ups.append(F.relu(NH._bn(st, u, p + ".1", 1e-3)))
        ref = torch.cat(ups, 1)
        got = neck.cuda()(x.cuda())
    assert rel_err(got.cpu().numpy(), ref.numpy()) < 1e-3


@pytest.mark.parametrize("B,H,W,cin,cout", [(2, 47, 47, 256, 256), (1, 188, 188, 128, 128), (3, 20, 33, 64, 320), (2, 9, 5, 512, 64),
                                            (1, 16, 8, 32, 16)])
def test_dense_grid_tma_conv_equals_table_conv_bitwise(B, H, W, cin, cout):
    """s2d_conv_fwd_grid (TMA boxes, 16 x 8-pixel tiles, zero fill = padding, ragged edge tiles) writes exactly the bits of the
    table-driven launch: every output row accumulates the same (chunk, offset) blocks in the same order."""
    torch.manual_seed(B * 1000 + H)
    n = B * H * W
    x = torch.relu(torch.randn(n, cin, device="cuda"))
    w = torch.randn(9, cin, cout, device="cuda") / (9 * cin) ** 0.5
    sc, sh = torch.rand(cout, device="cuda") + 0.5, torch.randn(cout, device="cuda")
    res = torch.randn(n, cout, device="cuda")
    tbl, Ho, Wo = dense.conv_table(torch.device("cuda"), B, H, W, 3, 1, 1)
    for r, act in ((None, dense.ACT_RELU), (res, dense.ACT_GELU)):
        a = dense.conv_rows(x, w, tbl, n, sc, sh, act, r, False, precision=ops.PRECISION_BF16X2)
        b = dense.conv_rows(x, w, tbl, n, sc, sh, act, r, False, precision=ops.PRECISION_BF16X2, grid=(B, H, W, 3, 1))
        assert torch.equal(a, b)
        if cout % 32 == 0:
            assert torch.equal(ops.get_split(a), ops.get_split(b))
