"""TEST INFRASTRUCTURE ONLY (never imported by the product path): torch restatement of the reference's distillation
training step for the CenterPoint student, used by tests/test_gpu_train_step.py as the gradient oracle.

The reference trains plain ``torch.nn`` modules (cuDNN) plus spconv; this file runs the SAME module objects
(``sparse2dense_b200`` keeps the reference's module tree, so ``neck.encoder_1`` etc. are the reference's
``nn.Sequential``s) through torch's own operators and autograd:

* sparse backbone (det3d/models/backbones/scn.py:42-85,156-185): spconv is not vendored ("parity
  unpinned"), so the convolution is the gather formulation of App. A over given neighbour tables,
  ``out[i] = sum_k in[tbl[k][i]] @ W[k]``, with train-mode ``F.batch_norm`` over the active rows;
* ``S2D_RPN.forward`` incl. the PCR branch (det3d/models/necks/rpn.py:300-337);
* ``CenterHead.forward`` / ``SepHead.forward`` (det3d/models/bbox_heads/center_head.py:100-110,236-244);
* ``CenterHead.loss`` (:250-291) with ``FastFocalLoss`` / ``RegLoss`` (det3d/models/losses/centernet_loss.py:6-54);
* ``fastfocalloss`` / ``distill_reg_loss`` / the sparse2dense terms (det3d/torchie/trainer/trainer.py:38-76,775-799);
* ``KD_VoxelNet.mask_offset_loss`` and the grid construction (det3d/models/detectors/voxelnet.py:171-185,229-249).
"""
import numpy as np
import torch
import torch.nn.functional as F


def gather_conv(x, w, tbl):
    """x [n_in, Cin], w [K, Cin, Cout], tbl int64 [K, n_out] with n_in marking "no neighbour"."""
    xp = torch.cat([x, x.new_zeros(1, x.shape[1])], 0)
    out = None
    for k in range(tbl.shape[0]):
        y = xp[tbl[k]] @ w[k]
        out = y if out is None else out + y
    return out


def _bn_rows(x, bn):
    return F.batch_norm(x, None, None, bn.weight, bn.bias, True, 0.0, bn.eps)


def _tbl(t, n_in):
    t = t.long()
    return torch.where(t < 0, torch.full_like(t, n_in), t)


def backbone_forward(bb, feats, tables, coors_out, batch, out_shape):
    """SpMiddleResNetFHD.forward in training mode.  tables: dict res0..res3 (SubM) and down1..down4 (strided), int32
    device tensors of the library; coors_out: coordinates [N4,4] (b,z,y,x) of the last layer -> NCHW [B, C*D, H, W]."""
    def conv(x, m, tbl):
        K = m.weight.shape[0] * m.weight.shape[1] * m.weight.shape[2]
        y = gather_conv(x, m.weight.view(K, m.in_channels, m.out_channels), _tbl(tbl, x.shape[0]))
        return y if m.bias is None else y + m.bias

    def block(x, blk, tbl):
        out = F.relu(_bn_rows(conv(x, blk.conv1, tbl), blk.bn1))
        out = _bn_rows(conv(out, blk.conv2, tbl), blk.bn2)
        return F.relu(out + x)

    x = F.relu(_bn_rows(conv(feats, bb.conv_input[0], tables["res0"]), bb.conv_input[1]))
    for blk in bb.conv1:
        x = block(x, blk, tables["res0"])
    for i, stage in enumerate((bb.conv2, bb.conv3, bb.conv4), 1):
        x = F.relu(_bn_rows(conv(x, stage[0], tables[f"down{i}"]), stage[1]))
        x = block(x, stage[3], tables[f"res{i}"])
        x = block(x, stage[4], tables[f"res{i}"])
    x = F.relu(_bn_rows(conv(x, bb.extra_conv[0], tables["down4"]), bb.extra_conv[1]))
    D, H, W = out_shape
    C = x.shape[1]
    return dense_from_voxels(x, coors_out, batch, (D, H, W)).view(batch, C * D, H, W)       # scn.py:173-176


def s2d_rpn_forward(neck, x, train_pcr=True):
    """rpn.py:300-337 (training mode: PCR branch on)."""
    y_1 = neck.encoder_1(x)
    y_2 = neck.encoder_2(y_1)
    att = neck.convnext_block_1(y_2) + y_2
    att = neck.convnext_block_2(att) + att
    att = F.gelu(neck.convnext_block_3(att) + att)
    y_3 = torch.cat([neck.decoder_1(att), y_1], 1)
    F_S_b = neck.decoder_2(y_3)
    F_S_a = neck.fusion_dense(F_S_b) + neck.fusion_sparse(x)
    gen = (None, None, None, None)
    if train_pcr:
        N, _, H, W = x.shape
        g = neck.out_conv(F_S_b).view(N, 128, 5, H, W)
        g = neck.generator_1(g)
        off4, mask4 = neck.gen_out_4(g), neck.gen_mask_4(g)
        g = neck.generator_2(g)
        mask2, off2 = neck.gen_mask_2(g), neck.gen_out_2(g)
        gen = (off2, mask2, off4, mask4)
    ups, h = [], F_S_a
    for i in range(len(neck.blocks)):
        h = neck.blocks[i](h)
        if i - neck._upsample_start_idx >= 0:
            ups.append(neck.deblocks[i - neck._upsample_start_idx](h))
    return torch.cat(ups, 1), gen, F_S_a, F_S_b


def center_head_forward(head, x):
    x = head.shared_conv(x)
    return [{h: getattr(task, h)(x) for h in task.heads} for task in head.tasks]


def _gather_feat(m, ind):
    B, C, H, W = m.shape
    m = m.permute(0, 2, 3, 1).reshape(B, H * W, C)
    return m.gather(1, ind.unsqueeze(2).expand(B, ind.shape[1], C))


def fast_focal(out, target, ind, mask, cat):
    """FastFocalLoss.forward == trainer.fastfocalloss."""
    mask = mask.float()
    gt = torch.pow(1 - target, 4)
    neg = (torch.log(1 - out) * torch.pow(out, 2) * gt).sum()
    pos_pred = _gather_feat(out, ind).gather(2, cat.unsqueeze(2))
    num = mask.sum()
    pos = (torch.log(pos_pred) * torch.pow(1 - pos_pred, 2) * mask.unsqueeze(2)).sum()
    return -neg if num == 0 else -(pos + neg) / num


def reg_loss(output, mask, ind, target):
    pred = _gather_feat(output, ind)
    m = mask.float().unsqueeze(2)
    loss = F.l1_loss(pred * m, target * m, reduction="none") / (m.sum() + 1e-4)
    return loss.transpose(2, 0).sum(dim=2).sum(dim=1)


def distill_reg(output, target, mask, ind):
    pred, gt = _gather_feat(output, ind), _gather_feat(target, ind)
    m = mask.float().unsqueeze(2)
    loss = F.mse_loss(pred * m, gt * m, reduction="none") / (m.sum() + 1e-4)
    return loss.transpose(2, 0).sum(dim=2).sum(dim=1)


def clamp_sigmoid(x):
    return torch.clamp(torch.sigmoid(x), min=1e-4, max=1 - 1e-4)


def center_head_loss(head, example, preds):
    p = preds[0]
    hm = clamp_sigmoid(p["hm"])
    hm_loss = fast_focal(hm, example["hm"][0], example["ind"][0], example["mask"][0], example["cat"][0])
    anno = torch.cat((p["reg"], p["height"], p["dim"], p["rot"]), 1)
    target = example["anno_box"][0][..., [0, 1, 2, 3, 4, 5, -2, -1]]
    box = reg_loss(anno, example["mask"][0], example["ind"][0], target)
    loc = (box * box.new_tensor(head.code_weights)).sum()
    return hm_loss + head.weight * loc, hm, anno, hm_loss, loc


def mask_offset_loss(gen_offset, gen_mask, gt, grid):
    """voxelnet.py:171-185."""
    gt_mask = gt.sum(1) != 0
    count_pos, count_neg = gt_mask.sum(), (~gt_mask).sum()
    beta = count_neg / count_pos
    loss = F.binary_cross_entropy_with_logits(gen_mask[:, 0], gt_mask.float(), pos_weight=beta)
    grid = grid * gt_mask[:, None]
    gt3 = gt[:, :3] - grid
    sel = gt3 != 0
    return loss, F.l1_loss(gen_offset[sel], gt3[sel])


def voxel_grid(N, D, H, W, like):
    """voxelnet.py:231-236 (note the H in the x half-cell term)."""
    zs, ys, xs = torch.meshgrid([torch.arange(0, D), torch.arange(0, H), torch.arange(0, W)], indexing="ij")
    ys = ys * (150.4 / H) - 75.2 + (150.4 / H) / 2
    xs = xs * (150.4 / W) - 75.2 + (150.4 / H) / 2
    zs = zs * (6 / D) - 2 + (6 / D) / 2
    return torch.cat([xs[None], ys[None], zs[None]], 0)[None].repeat(N, 1, 1, 1, 1).to(like)


def dense_from_voxels(feats, coors, batch, shape):
    """spconv.SparseConvTensor(...).dense() -> [B, C, D, H, W]."""
    D, H, W = shape
    out = feats.new_zeros(batch, D, H, W, feats.shape[1])
    b, z, y, x = (coors[:, i].long() for i in range(4))
    out = out.index_put((b, z, y, x), feats)
    return out.permute(0, 4, 1, 2, 3).contiguous()


def distill_total(head, loss_head, F_S_a, F_S_b, F_D_a, F_D_b, hm_s, anno_s, T_preds, example, mask_loss=0, comp_loss=0):
    """trainer.py:775-799 (CenterPoint branch)."""
    inds = F_D_a > 0
    s2d = F.mse_loss(F_S_a[inds], F_D_a[inds]) * 10 + F.mse_loss(F_S_a[~inds], F_D_a[~inds]) * 20
    inds = F_D_b > 0
    s2d = s2d + F.mse_loss(F_S_b[inds], F_D_b[inds]) * 5 + F.mse_loss(F_S_b[~inds], F_D_b[~inds]) * 20
    T = T_preds[0]
    kd_hm = fast_focal(hm_s, torch.sigmoid(T["hm"]), example["ind"][0], example["mask"][0], example["cat"][0])
    anno_t = torch.cat((T["reg"], T["height"], T["dim"], T["rot"]), 1)
    kd_reg = distill_reg(anno_s, anno_t, example["mask"][0], example["ind"][0])
    kd_reg = (kd_reg * kd_reg.new_tensor(head.code_weights)).sum() * head.weight
    total = loss_head + kd_hm + kd_reg + s2d + (mask_loss + comp_loss)
    return total, dict(sparse2dense_loss=s2d, kd_hm_loss=kd_hm, kd_reg_loss=kd_reg)
