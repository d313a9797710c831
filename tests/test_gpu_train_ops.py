"""GPU parity of the training-mode operators (csrc/train.cu through sparse2dense_b200/autograd.py) against the
reference's own arithmetic: torch.nn modules in float64 on the CPU with torch autograd (the reference trains exactly those
modules, rpn.py / center_head.py / norm.py), and for the sparse convolution the gather formulation of App. A over the same
neighbour table in float64.  Bar: 1e-4 relative for fp32 / TF32x3 / AUTO arithmetic (gradients are sums over up to 10^4
rows); exact for the integer table transpose."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F
from torch import nn

from sparse2dense_b200 import _lib, dense, ops
from sparse2dense_b200 import autograd as AG

pytestmark = pytest.mark.gpu
DEV = "cuda"


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


def random_table(K, n_in, n_out, fill, seed):
    """Per tap an injective partial map i -> j (as every table of the library is)."""
    rng = np.random.default_rng(seed)
    tbl = np.full((K, n_out), -1, np.int32)
    for k in range(K):
        m = min(n_in, n_out)
        outs = rng.permutation(n_out)[:m]
        ins = rng.permutation(n_in)[:m]
        keep = rng.random(m) < fill
        tbl[k, outs[keep]] = ins[keep]
    return tbl


def test_table_transpose_exact():
    K, n_in, n_out = 27, 1000, 1300
    tbl = random_table(K, n_in, n_out, 0.4, 0)
    t = AG.Table(torch.from_numpy(tbl).to(DEV), n_in, n_out)
    inv = t.transposed().cpu().numpy()
    want = np.full((K, n_in), -1, np.int32)
    for k in range(K):
        i = np.nonzero(tbl[k] >= 0)[0]
        want[k, tbl[k, i]] = i
    assert np.array_equal(inv, want)
    # with an output-row indirection (sub-pixel classes of a transposed convolution)
    rows = np.random.default_rng(1).permutation(5 * n_out)[:n_out].astype(np.int32)
    t2 = AG.Table(torch.from_numpy(tbl).to(DEV), n_in, n_out, out_rows=torch.from_numpy(rows).to(DEV))
    inv2 = t2.transposed().cpu().numpy()
    assert np.array_equal(inv2[want >= 0], rows[want[want >= 0]]) and np.all(inv2[want < 0] == -1)


def gather_conv_ref(x, w, tbl):
    """out[i] = sum_k x[tbl[k][i]] . w[k] in float64 with torch autograd (SURVEY App. A restated on the table)."""
    xp = torch.cat([x, x.new_zeros(1, x.shape[1])], 0)
    idx = torch.from_numpy(np.where(tbl < 0, x.shape[0], tbl).astype(np.int64))
    return sum(xp[idx[k]] @ w[k] for k in range(tbl.shape[0]))


@pytest.mark.parametrize("precision", [ops.PRECISION_FP32, ops.PRECISION_AUTO])
@pytest.mark.parametrize("cin,cout,K", [(5, 16, 27), (16, 16, 27), (32, 64, 27), (128, 128, 27), (128, 128, 3), (64, 3, 9),
                                        (256, 1024, 1)])
def test_gather_conv_forward_backward(cin, cout, K, precision):
    n_in, n_out = 2100, 1777
    tbl = random_table(K, n_in, n_out, 0.35, cin + cout)
    g = torch.Generator().manual_seed(cin * 7 + cout)
    x = torch.randn(n_in, cin, generator=g, dtype=torch.float64, requires_grad=True)
    w = (torch.randn(K, cin, cout, generator=g, dtype=torch.float64) / np.sqrt(cin * K * 0.35)).requires_grad_(True)
    dy = torch.randn(n_out, cout, generator=g, dtype=torch.float64)
    ref = gather_conv_ref(x, w, tbl)
    ref.backward(dy)

    xc = x.detach().float().to(DEV).requires_grad_(True)
    wc = w.detach().float().to(DEV).requires_grad_(True)
    table = AG.Table(torch.from_numpy(tbl).to(DEV), n_in, n_out)
    out = AG.GatherConv.apply(xc, wc, table, precision)
    out.backward(dy.float().to(DEV))
    assert rel(out, ref) < 1e-4
    assert rel(xc.grad, x.grad) < 1e-4
    assert rel(wc.grad, w.grad) < 1e-4


@pytest.mark.parametrize("cg,cd,K", [(3, 16, 8), (5, 16, 27), (32, 3, 1), (32, 1, 1), (3, 1, 1), (3, 3, 1), (16, 3, 8),
                                     (64, 64, 9), (128, 256, 3), (20, 36, 4), (16, 16, 27), (32, 32, 27), (48, 40, 5),
                                     (24, 30, 3), (10, 13, 2),
                                     # tensor-core kernel (wgrad_tc.cu): lone / odd chunk counts, partial channel blocks, many blocks
                                     (32, 64, 2), (96, 160, 4), (256, 512, 1), (320, 32, 9), (512, 64, 2), (128, 128, 27)])
def test_conv_wgrad_shapes(cg, cd, K):
    """Row-parallel small-channel kernel, vectorised and scalar tile loaders, with and without a row indirection on D."""
    n_g, n_rows = 5000, 4321
    tbl = random_table(K, n_g, n_rows, 0.5, cg * 31 + cd)
    rng = np.random.default_rng(cg + cd)
    G, Dm = rng.normal(size=(n_g, cg)), rng.normal(size=(2 * n_rows, cd))
    d_rows = rng.permutation(2 * n_rows)[:n_rows].astype(np.int32)
    for use_rows in (False, True):
        di = d_rows if use_rows else np.arange(n_rows)
        want = np.zeros((K, cg, cd))
        for k in range(K):
            i = np.nonzero(tbl[k] >= 0)[0]
            want[k] = G[tbl[k, i]].T @ Dm[di[i]]
        got = AG.conv_wgrad(torch.from_numpy(G).float().to(DEV), torch.from_numpy(Dm).float().to(DEV),
                            torch.from_numpy(tbl).to(DEV), n_rows, torch.from_numpy(d_rows).to(DEV) if use_rows else None)
        assert float(np.abs(got.double().cpu().numpy() - want).max() / np.abs(want).max()) < 2e-5


def test_subm_table_is_its_own_transpose():
    """The SubM rulebook satisfies tbl[k][i] = j <=> tbl[K-1-k][j] = i, which GatherConv uses for the data gradient."""
    rng = np.random.default_rng(4)
    cells = rng.choice(2 * 9 * 24 * 20, 1500, replace=False)
    b, r = np.divmod(cells, 9 * 24 * 20)
    z, r = np.divmod(r, 24 * 20)
    y, x = np.divmod(r, 20)
    coors = torch.from_numpy(np.stack([b, z, y, x], 1).astype(np.int32)).to(DEV)
    index = ops.build_grid_index(coors, 2, (9, 24, 20))
    tbl = ops.rulebook_subm(coors, index, 3)
    n = coors.shape[0]
    assert torch.equal(AG.Table(tbl, n, n).transposed(), torch.flip(tbl, dims=[0]))


ACTS = {dense.ACT_NONE: lambda z: z, dense.ACT_RELU: F.relu, dense.ACT_GELU: F.gelu}


@pytest.mark.parametrize("act", [dense.ACT_NONE, dense.ACT_RELU, dense.ACT_GELU])
@pytest.mark.parametrize("mode", ["bn", "bn_res", "bn_res_after", "bias", "bias_res"])
@pytest.mark.parametrize("n,C", [(3000, 16), (777, 96), (2048, 640)])
def test_rows_norm_act(n, C, mode, act):
    g = torch.Generator().manual_seed(n + C)
    x = (torch.randn(n, C, generator=g, dtype=torch.float64) * 2 + 0.5).requires_grad_(True)
    res = torch.randn(n, C, generator=g, dtype=torch.float64).requires_grad_(True) if "res" in mode else None
    dy = torch.randn(n, C, generator=g, dtype=torch.float64)
    after = mode.endswith("after")
    if mode.startswith("bn"):
        bn = nn.BatchNorm1d(C, eps=1e-3, momentum=0.01).double().train()
        bn.weight.data.uniform_(0.5, 1.5, generator=g)
        bn.bias.data.normal_(0, 0.2, generator=g)
        z = bn(x)
    else:
        bias = (torch.randn(C, generator=g, dtype=torch.float64) * 0.3).requires_grad_(True)
        z = x + bias
    ref = (ACTS[act](z) + res) if (res is not None and after) else ACTS[act](z if res is None else z + res)
    ref.backward(dy)

    xc = x.detach().float().to(DEV).requires_grad_(True)
    rc = None if res is None else res.detach().float().to(DEV).requires_grad_(True)
    if mode.startswith("bn"):
        bnc = nn.BatchNorm1d(C, eps=1e-3, momentum=0.01).to(DEV).train()
        bnc.weight.data.copy_(bn.weight.data.float())
        bnc.bias.data.copy_(bn.bias.data.float())
        out = AG.norm_act(xc, bnc, None, act, rc, after)
    else:
        bc = bias.detach().float().to(DEV).requires_grad_(True)
        out = AG.norm_act(xc, None, bc, act, rc, after)
    out.backward(dy.float().to(DEV))
    assert rel(out, ref) < 2e-5
    assert rel(xc.grad, x.grad) < 1e-4
    if rc is not None:
        assert rel(rc.grad, res.grad) < 1e-4
    if mode.startswith("bn"):
        assert rel(bnc.weight.grad, bn.weight.grad) < 1e-4 and rel(bnc.bias.grad, bn.bias.grad) < 1e-4
        assert rel(bnc.running_mean, bn.running_mean) < 1e-5 and rel(bnc.running_var, bn.running_var) < 1e-5
        assert int(bnc.num_batches_tracked) == 1
    else:
        assert rel(bc.grad, bias.grad) < 1e-4


def _rows(x):          # NCHW -> rows, differentiable (torch permute: layout plumbing of the test only)
    B, C, H, W = x.shape
    return x.permute(0, 2, 3, 1).reshape(B * H * W, C)


def _nchw(rows, B, H, W):
    return rows.reshape(B, H, W, -1).permute(0, 3, 1, 2)


@pytest.mark.parametrize("precision", [ops.PRECISION_FP32, ops.PRECISION_AUTO])
@pytest.mark.parametrize("kind", ["c3s1", "c3s2", "c2s2", "c1", "t4", "t2", "c3s1_nobn"])
def test_dense_conv_bn_gelu_training_vs_torch(kind, precision):
    torch.manual_seed(5)
    B, H, W, cin, cout = 2, 12, 10, 64, 96
    mk = dict(c3s1=lambda: nn.Conv2d(cin, cout, 3, 1, 1), c3s2=lambda: nn.Conv2d(cin, cout, 3, 2, 1),
              c2s2=lambda: nn.Conv2d(cin, cout, 2, 2), c1=lambda: nn.Conv2d(cin, cout, 1),
              t4=lambda: nn.ConvTranspose2d(cin, cout, 4, 2, 1), t2=lambda: nn.ConvTranspose2d(cin, cout, 2, 2, bias=False),
              c3s1_nobn=lambda: nn.Conv2d(cin, cout, 3, 1, 1))
    conv = mk[kind]().double().train()
    bn = None if kind.endswith("nobn") else nn.BatchNorm2d(cout).double().train()
    if bn is not None:
        bn.weight.data.uniform_(0.5, 1.5); bn.bias.data.normal_(0, 0.1)
    x = torch.randn(B, cin, H, W, dtype=torch.float64, requires_grad=True)
    z = conv(x)
    ref = F.gelu(z if bn is None else bn(z))
    dy = torch.randn_like(ref)
    ref.backward(dy)

    import copy
    conv_c = copy.deepcopy(conv).float().to(DEV)
    bn_c = None if bn is None else nn.BatchNorm2d(cout).to(DEV).train()
    if bn is not None:
        bn_c.weight.data.copy_(bn.weight.data.float()); bn_c.bias.data.copy_(bn.bias.data.float())
    conv_c.zero_grad()
    xc = x.detach().float().to(DEV).requires_grad_(True)
    D = dense.DenseOps(precision)
    D.training = True
    if kind.startswith("t"):
        y, Ho, Wo = D.tconv("m", _rows(xc), B, H, W, conv_c, bn_c, dense.ACT_GELU)
    else:
        y, Ho, Wo = D.conv("m", _rows(xc), B, H, W, conv_c, bn_c, dense.ACT_GELU)
    out = _nchw(y, B, Ho, Wo)
    assert tuple(out.shape) == tuple(ref.shape)
    out.backward(dy.float().to(DEV))
    assert rel(out, ref) < 1e-4
    assert rel(xc.grad, x.grad) < 1e-4
    assert rel(conv_c.weight.grad, conv.weight.grad) < 1e-4
    if bn is not None:
        assert rel(bn_c.weight.grad, bn.weight.grad) < 1e-4 and rel(bn_c.bias.grad, bn.bias.grad) < 1e-4
        assert rel(bn_c.running_mean, bn.running_mean) < 1e-4 and rel(bn_c.running_var, bn.running_var) < 1e-4
        if conv.bias is not None:                          # cancels in the normalisation: exactly zero
            assert conv_c.bias.grad is not None and float(conv_c.bias.grad.abs().max()) == 0.0
    elif conv.bias is not None:
        assert rel(conv_c.bias.grad, conv.bias.grad) < 1e-4


def test_convnext_block_training_vs_torch():
    """dw 7x7 -> LayerNorm([C,H,W]) -> 1x1 -> GELU -> 1x1, residual (rpn.py:204-222,306-308)."""
    torch.manual_seed(6)
    B, C, H, W = 2, 32, 9, 9
    blk = nn.Sequential(nn.Conv2d(C, C, 7, padding=3, groups=C), nn.LayerNorm([C, H, W], eps=1e-6), nn.Conv2d(C, 4 * C, 1),
                        nn.GELU(), nn.Conv2d(4 * C, C, 1)).double().train()
    blk[1].weight.data.uniform_(0.5, 1.5); blk[1].bias.data.normal_(0, 0.1)
    x = torch.randn(B, C, H, W, dtype=torch.float64, requires_grad=True)
    ref = F.gelu(blk(x) + x)
    dy = torch.randn_like(ref)
    ref.backward(dy)

    import copy
    bc = copy.deepcopy(blk).float().to(DEV)
    bc.zero_grad()
    xc = x.detach().float().to(DEV).requires_grad_(True)
    D = dense.DenseOps(ops.PRECISION_AUTO)
    D.training = True
    att = _rows(xc)
    t = D.dwconv(att, B, H, W, bc[0])
    t = D.layernorm(t, B, H, W, bc[1])
    t, _, _ = D.conv("a", t, B, H, W, bc[2], None, dense.ACT_GELU)
    y, _, _ = D.conv("b", t, B, H, W, bc[4], None, dense.ACT_GELU, residual=att)
    out = _nchw(y, B, H, W)
    out.backward(dy.float().to(DEV))
    assert rel(out, ref) < 1e-4
    assert rel(xc.grad, x.grad) < 2e-4
    for (name, p), (_, q) in zip(bc.named_parameters(), blk.named_parameters()):
        assert rel(p.grad, q.grad) < 2e-4, name


def test_sparse_block_training_vs_gather_reference():
    """conv_input-style layer + SparseBasicBlock in training mode (scn.py:42-85,104-112): outputs, running statistics and
    every parameter gradient against a float64 restatement over the same rulebook."""
    from sparse2dense_b200 import spconv
    from sparse2dense_b200.backbones import SparseBasicBlock
    torch.manual_seed(7)
    rng = np.random.default_rng(7)
    shape, B, n = (9, 24, 20), 2, 1500
    cells = rng.choice(B * 9 * 24 * 20, n, replace=False)
    b, r = np.divmod(cells, 9 * 24 * 20)
    z, r = np.divmod(r, 24 * 20)
    y, x = np.divmod(r, 20)
    coors = torch.from_numpy(np.stack([b, z, y, x], 1).astype(np.int32)).to(DEV)
    feats = torch.randn(n, 5)
    net = spconv.SparseSequential(spconv.SubMConv3d(5, 16, 3, bias=False, indice_key="res0"),
                                  nn.BatchNorm1d(16, eps=1e-3, momentum=0.01), nn.ReLU(),
                                  SparseBasicBlock(16, 16, indice_key="res0"),
                                  spconv.SparseConv3d(16, 32, 3, 2, padding=1, bias=False),
                                  nn.BatchNorm1d(32, eps=1e-3, momentum=0.01), nn.ReLU()).to(DEV).train()
    for m in net.modules():
        if isinstance(m, spconv.SparseConvolution):
            m.precision = ops.PRECISION_AUTO
    xin = spconv.SparseConvTensor(feats.to(DEV), coors, shape, B)
    out = net(xin)
    dy = torch.randn(out.features.shape[0], 32)
    out.features.backward(dy.to(DEV))

    # float64 restatement on the CPU over the same tables
    tbl0 = xin.indice_dict["res0"].tbl.cpu().numpy()
    index = xin.index()
    sc = ops.sparse_out_coords(coors, n, B, shape, 3, 2, 1, 1)
    tbl1 = ops.rulebook_sparse(sc.coors, index, 3, 2, 1, 1).cpu().numpy()
    assert torch.equal(sc.coors, out.indices)
    P = {k: v.detach().double().cpu().requires_grad_(True) for k, v in net.named_parameters()}

    def bn(x, pre):
        mean, var = x.mean(0), x.var(0, unbiased=False)
        return (x - mean) / torch.sqrt(var + 1e-3) * P[pre + ".weight"] + P[pre + ".bias"]

    x0 = feats.double()
    h = F.relu(bn(gather_conv_ref(x0, P["0.weight"].view(27, 5, 16), tbl0), "1"))
    t = F.relu(bn(gather_conv_ref(h, P["3.conv1.weight"].view(27, 16, 16), tbl0) + P["3.conv1.bias"], "3.bn1"))
    t = F.relu(bn(gather_conv_ref(t, P["3.conv2.weight"].view(27, 16, 16), tbl0) + P["3.conv2.bias"], "3.bn2") + h)
    ref = F.relu(bn(gather_conv_ref(t, P["4.weight"].view(27, 16, 32), tbl1), "5"))
    ref.backward(dy.double())
    assert rel(out.features, ref) < 1e-4
    for k, v in net.named_parameters():
        if k.endswith("conv1.bias") or k.endswith("conv2.bias"):
            assert float(v.grad.abs().max()) == 0.0          # bias in front of a training BN
            continue
        assert rel(v.grad, P[k].grad) < 5e-4, k


def test_adam_step_and_grad_clip_vs_torch():
    """fastai OptimWrapper.step with true_wd (fastai_optim.py:158-174): p *= 1 - wd*lr, then torch Adam; and
    clip_grad_norm_(35) (hooks/optimizer.py:15-21)."""
    torch.manual_seed(8)
    n = 100003
    p = torch.randn(n)
    ref_p = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([ref_p], lr=0.0, betas=(0.9, 0.99), eps=1e-8)
    pc = p.to(DEV)
    m, v = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    lib = _lib.load()
    ws = torch.empty((lib.s2d_grad_norm_workspace_bytes(),), dtype=torch.uint8, device=DEV)
    out2 = torch.empty(2, device=DEV)
    for step in range(1, 6):
        g = torch.randn(n) * (3.0 if step % 2 else 0.01)
        lr, wd, b1 = 1e-3 * step, 0.01, 0.95 - 0.02 * step
        ref_p.grad = g.clone()
        total = torch.nn.utils.clip_grad_norm_([ref_p], 35.0)
        with torch.no_grad():
            ref_p.mul_(1 - wd * lr)
        for grp in opt.param_groups:
            grp["lr"], grp["betas"] = lr, (b1, 0.99)
        opt.step()
        gc = g.to(DEV)
        _lib.check(lib.s2d_grad_norm_clip(gc.data_ptr(), n, 35.0, out2.data_ptr(), ws.data_ptr(), ws.numel(),
                                          torch.cuda.current_stream().cuda_stream))
        _lib.check(lib.s2d_adam_step(pc.data_ptr(), gc.data_ptr(), m.data_ptr(), v.data_ptr(), n, lr, b1, 0.99, 1e-8, wd,
                                     step, out2[1:].data_ptr(), torch.cuda.current_stream().cuda_stream))
        assert abs(float(out2[0]) - float(total)) < 1e-3 * float(total)
        assert rel(pc, ref_p) < 1e-5
