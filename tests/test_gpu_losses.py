"""GPU parity of the loss reductions (csrc/losses.cu) against the oracle and the reference-generated values; maps are
passed both as NCHW tensors and as NHWC row slices (how the heads of this package hold them)."""
import os

import numpy as np
import pytest
import torch

from loss_common import loss_inputs
from oracle import ref_ops as R
from sparse2dense_b200 import losses as L
from test_oracle_losses import G, oracle_losses

pytestmark = pytest.mark.gpu


def rows_of(x):
    """NCHW numpy -> Rows view over a wider buffer (column offset 3) to exercise strides."""
    B, C, H, W = x.shape
    buf = torch.zeros((B * H * W, C + 5), dtype=torch.float32, device="cuda")
    buf[:, 3:3 + C] = torch.from_numpy(np.ascontiguousarray(x.transpose(0, 2, 3, 1)).reshape(B * H * W, C)).cuda()
    return L.Rows(buf[:, 3:3 + C], B, H * W)


@pytest.mark.parametrize("as_rows", [False, True])
def test_losses_vs_reference_values_and_oracle(as_rows):
    g = np.load(G)
    for seed in (40, 41):
        d = loss_inputs(seed)
        want = oracle_losses(d)
        m = (lambda a: rows_of(a)) if as_rows else (lambda a: torch.from_numpy(a).cuda())
        ind, mask, cat = (torch.from_numpy(d[k]).cuda() for k in ("ind", "mask", "cat"))
        got = dict(
            hm=L.fastfocalloss(m(d["hm_logits"]), m(d["gt_hm"]), ind, mask, cat, out_is_logits=True),
            kd=L.fastfocalloss(m(d["hm_logits"]), m(d["t_logits"]), ind, mask, cat, out_is_logits=True, target_is_logits=True),
            reg=L.RegLoss()(m(d["box"]), mask, ind, torch.from_numpy(d["anno"]).cuda()),
            dl=L.distill_reg_loss(m(d["box"]), m(d["t_box"]), mask, ind),
            s2d=L.sparse2dense_loss(torch.from_numpy(d["box"]).cuda(), torch.from_numpy(d["t_box"]).cuda(),
                                    torch.from_numpy(d["hm_logits"]).cuda(), torch.from_numpy(d["t_logits"]).cuda()))
        for k, v in got.items():
            v = v.cpu().numpy()
            np.testing.assert_allclose(v, np.asarray(want[k]), rtol=2e-5, atol=1e-6, err_msg=k)
            np.testing.assert_allclose(v, g[f"{seed}_{k}"], rtol=3e-5, atol=1e-6, err_msg=k)


def test_masked_mse_full_size_deterministic_and_exact_counts():
    """[4,256,188,188] maps (the F_S_a / F_D_a of a batch of 4): counts exact, sums vs float64 numpy, bitwise repeatable."""
    rng = np.random.default_rng(0)
    shape = (4, 256, 188, 188)
    fd = (np.abs(rng.normal(0, 1, shape)) * (rng.uniform(size=shape) < 0.3)).astype(np.float32)
    fs = (fd + rng.normal(0, 0.1, shape)).astype(np.float32)
    a = L.masked_mse_terms(torch.from_numpy(fs).cuda(), torch.from_numpy(fd).cuda())
    b = L.masked_mse_terms(torch.from_numpy(fs).cuda(), torch.from_numpy(fd).cuda())
    assert torch.equal(a, b)
    q = (fs.astype(np.float64) - fd.astype(np.float64)) ** 2
    p = fd > 0
    a = a.cpu().numpy()
    assert a[1] == p.sum() and a[3] == (~p).sum()
    np.testing.assert_allclose(a[0], q[p].sum(), rtol=1e-6)
    np.testing.assert_allclose(a[2], q[~p].sum(), rtol=1e-6)
    want = R.sparse2dense_loss(fs, fd, fs, fd)
    got = L.sparse2dense_loss(torch.from_numpy(fs).cuda(), torch.from_numpy(fd).cuda(), torch.from_numpy(fs).cuda(),
                              torch.from_numpy(fd).cuda()).item()
    assert abs(got - want) < 1e-5 * abs(want)


def test_center_head_loss_matches_oracle():
    from sparse2dense_b200 import registry
    d = loss_inputs(40)
    head = registry.build_head(dict(type="CenterHead", in_channels=512, tasks=[dict(num_class=3, class_names=["a", "b", "c"])],
                                    dataset="waymo", weight=2, code_weights=[1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 0.5, 0.5],
                                    common_heads={"reg": (2, 2), "height": (1, 2), "dim": (3, 2), "rot": (2, 2)})).cuda()
    box = torch.from_numpy(d["box"]).cuda()
    preds = [dict(reg=box[:, 0:2], height=box[:, 2:3], dim=box[:, 3:6], rot=box[:, 6:8], hm=torch.from_numpy(d["hm_logits"]).cuda())]
    anno10 = np.concatenate([d["anno"][..., :6], np.zeros_like(d["anno"][..., :2]), d["anno"][..., 6:8]], -1)   # [.., vel x2, rot x2]
    example = dict(hm=[torch.from_numpy(d["gt_hm"]).cuda()], ind=[torch.from_numpy(d["ind"]).cuda()],
                   mask=[torch.from_numpy(d["mask"]).cuda()], cat=[torch.from_numpy(d["cat"]).cuda()],
                   anno_box=[torch.from_numpy(anno10).cuda()])
    ret = head.loss(example, preds)
    w = oracle_losses(d)
    loc = float((w["reg"] * np.array([1, 1, 1, 1, 1, 1, 0.5, 0.5], np.float32)).sum())
    assert abs(float(ret["hm_loss"][0]) - float(w["hm"])) < 1e-4 * abs(float(w["hm"]))
    assert abs(float(ret["loc_loss"][0]) - loc) < 1e-4 * abs(loc)
    assert abs(float(ret["loss"][0]) - (float(w["hm"]) + 2 * loc)) < 1e-4 * abs(float(w["hm"]) + 2 * loc)
    assert float(ret["num_positive"][0]) == d["mask"].sum()


@pytest.mark.parametrize("seed", [60, 61])
def test_pcr_loss_kernel_matches_reference_golden_and_autograd(seed):
    """s2d_pcr_loss / s2d_pcr_loss_bwd (sparse voxel list, rows in (b,y,x,z) order) against the values of the reference's own
    KD_VoxelNet.mask_offset_loss on dense tensors (tests/golden/pcr_loss.npz) and its torch-autograd gradients."""
    import os
    from conftest import GOLDEN
    from oracle import train_ref as TR
    from sparse2dense_b200 import pcr
    g = np.load(os.path.join(GOLDEN, "pcr_loss.npz"))
    B, D, H, W = (int(v) for v in g[f"{seed}_dims"])
    off = torch.from_numpy(g[f"{seed}_gen_offset"]).requires_grad_(True)
    msk = torch.from_numpy(g[f"{seed}_gen_mask"]).requires_grad_(True)
    gt = TR.dense_from_voxels(torch.from_numpy(g[f"{seed}_feats"]), torch.from_numpy(g[f"{seed}_coors"]), B, (D, H, W))
    m_ref, o_ref = TR.mask_offset_loss(off, msk, gt, TR.voxel_grid(B, D, H, W, gt))
    (0.7 * m_ref + 1.3 * o_ref).backward()

    to_rows = lambda t: t.detach().permute(0, 3, 4, 2, 1).reshape(-1, t.shape[1]).contiguous().cuda()   # [B,C,D,H,W] -> (b,y,x,z) rows
    off_r, msk_r = to_rows(off).requires_grad_(True), to_rows(msk).requires_grad_(True)
    out = pcr._PcrLoss.apply(msk_r, off_r, torch.from_numpy(g[f"{seed}_coors"]).cuda(), torch.from_numpy(g[f"{seed}_feats"]).cuda(),
                             (B, D, H, W), pcr._centre9(D, H, W))
    assert abs(float(out[0]) - float(g[f"{seed}_mask_loss"])) < 2e-6 * max(1.0, float(g[f"{seed}_mask_loss"]))
    assert abs(float(out[1]) - float(g[f"{seed}_offset_loss"])) < 2e-6
    (0.7 * out[0] + 1.3 * out[1]).backward()
    assert torch.allclose(off_r.grad.cpu(), to_rows(off.grad).cpu(), rtol=1e-5, atol=1e-9)
    assert torch.allclose(msk_r.grad.cpu(), to_rows(msk.grad).cpu(), rtol=1e-4, atol=1e-9)
