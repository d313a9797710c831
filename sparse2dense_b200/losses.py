"""Losses of the detection / distillation training step:
``FastFocalLoss`` / ``RegLoss`` (det3d/models/losses/centernet_loss.py:6-54), ``fastfocalloss`` / ``distill_reg_loss`` and the
``sparse2dense_loss`` expression of ``TS_Trainer.batch_processor_inline`` (det3d/torchie/trainer/trainer.py:38-76,783-799),
each as ONE deterministic reduction launch (csrc/losses.cu).  When the student's map requires grad the value is produced
by a ``torch.autograd.Function`` whose backward is the matching ``s2d_*_bwd`` kernel (teacher maps / targets are constants).

A *map* argument is either a torch NCHW tensor ``[B,C,H,W]`` or a ``Rows`` view (NHWC rows ``[B*H*W, c]`` that may be a
column slice of a wider buffer, as the heads of this package produce them)."""
from collections import namedtuple

import torch
from torch import nn

from . import _lib, ops

Rows = namedtuple("Rows", "rows B HW")     # rows: 2-D tensor with stride(1) == 1


def _view(m):
    """-> (tensor keeping the storage alive, ptr, sb, sc, scell, B, C, HW)."""
    if isinstance(m, Rows):
        r = m.rows
        assert r.dim() == 2 and r.stride(1) == 1 and r.shape[0] == m.B * m.HW and r.dtype == torch.float32
        return r, r.data_ptr(), m.HW * r.stride(0), 1, r.stride(0), m.B, r.shape[1], m.HW
    t = m.contiguous().float()
    B, C, H, W = t.shape
    return t, t.data_ptr(), C * H * W, H * W, 1, B, C, H * W


def _ws(dev):
    n = _lib.load().s2d_loss_workspace_bytes()
    return torch.empty((n,), dtype=torch.uint8, device=dev), n


def masked_mse_terms(f_student, f_teacher):
    """-> float64 [4]: sum / count over teacher > 0, sum / count over teacher <= 0 of (student - teacher)^2."""
    ops._need_cuda(f_student, f_teacher)
    fs, fd = f_student.contiguous().float(), f_teacher.contiguous().float()
    assert fs.numel() == fd.numel()
    out = torch.empty((4,), dtype=torch.float64, device=fs.device)
    ws, n = _ws(fs.device)
    _lib.check(_lib.load().s2d_masked_mse(fs.data_ptr(), fd.data_ptr(), fs.numel(), out.data_ptr(), ws.data_ptr(), n,
                                          ops._stream()), "s2d_masked_mse")
    return out


class _MaskedMSE(torch.autograd.Function):
    """w_pos * MSE(s | t > 0) + w_neg * MSE(s | t <= 0).

    Empty partition (a map that is all-positive or all-non-positive): the VALUE is NaN (0 / 0), exactly what the reference's
    ``F.mse_loss`` returns for an empty boolean selection (trainer.py:783-789), and the GRADIENT of that partition is zero --
    also what the reference produces: autograd routes the NaN term's gradient through ``index`` of an empty selection, i.e. to
    no element, so the student's gradient stays finite while the logged loss is NaN.  Value and gradient are therefore
    consistent with the reference, not with each other."""

    @staticmethod
    def forward(ctx, fs, ft, w_pos, w_neg):
        fs_c, ft_c = fs.contiguous().float(), ft.contiguous().float()
        out4 = masked_mse_terms(fs_c, ft_c)
        ctx.save_for_backward(fs_c, ft_c, out4)
        ctx.w = (float(w_pos), float(w_neg))
        return (w_pos * out4[0] / out4[1] + w_neg * out4[2] / out4[3]).float()

    @staticmethod
    def backward(ctx, g):
        fs, ft, out4 = ctx.saved_tensors
        d = torch.empty_like(fs)
        up = g.float().contiguous()
        _lib.check(_lib.load().s2d_masked_mse_bwd(fs.data_ptr(), ft.data_ptr(), fs.numel(), out4.data_ptr(), ctx.w[0],
                                                  ctx.w[1], up.data_ptr(), d.data_ptr(), ops._stream()), "s2d_masked_mse_bwd")
        return d, None, None, None


def sparse2dense_loss(F_S_a, F_D_a, F_S_b, F_D_b, w=(10.0, 20.0, 5.0, 20.0)):
    """trainer.py:783-789: 10*MSE(a | F_D_a>0) + 20*MSE(a | <=0) + 5*MSE(b | F_D_b>0) + 20*MSE(b | <=0) (fp32 scalar)."""
    if torch.is_grad_enabled() and (F_S_a.requires_grad or F_S_b.requires_grad):
        return _MaskedMSE.apply(F_S_a, F_D_a, w[0], w[1]) + _MaskedMSE.apply(F_S_b, F_D_b, w[2], w[3])
    a, b = masked_mse_terms(F_S_a, F_D_a), masked_mse_terms(F_S_b, F_D_b)
    return (w[0] * a[0] / a[1] + w[1] * a[2] / a[3] + w[2] * b[0] / b[1] + w[3] * b[2] / b[3]).float()


def _tensor_of(m):
    return m.rows if isinstance(m, Rows) else m


def _rewrap(m, t):
    return Rows(t, m.B, m.HW) if isinstance(m, Rows) else t


def _grad_buffer(m, t):
    """Contiguous gradient buffer shaped like the student's tensor and its (sb, sc, scell) strides."""
    if isinstance(m, Rows):
        g = torch.empty((t.shape[0], t.shape[1]), dtype=torch.float32, device=t.device)
        return g, m.HW * t.shape[1], 1, t.shape[1]
    B, C, H, W = t.shape
    g = torch.empty((B, C, H, W), dtype=torch.float32, device=t.device)
    return g, C * H * W, H * W, 1


class _Focal(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t, m_out, target, ind, mask8, cat, out_is_logits, target_is_logits):
        m = _rewrap(m_out, t)
        ko, po, osb, osc, oscell, B, C, HW = _view(m)
        kt, pt, tsb, tsc, tscell, Bt, Ct, HWt = _view(target)
        assert (B, C, HW) == (Bt, Ct, HWt)
        res = torch.empty((3,), dtype=torch.float64, device=ko.device)
        ws, n = _ws(ko.device)
        _lib.check(_lib.load().s2d_focal_loss(po, osb, osc, oscell, int(out_is_logits), pt, tsb, tsc, tscell,
                                              int(target_is_logits), B, C, HW, ind.data_ptr(), mask8.data_ptr(),
                                              cat.data_ptr(), ind.shape[1], res.data_ptr(), ws.data_ptr(), n, ops._stream()),
                   "s2d_focal_loss")
        ctx.save_for_backward(ko, kt, ind, mask8, cat, res)
        ctx.meta = (m_out, (osb, osc, oscell), (tsb, tsc, tscell), (B, C, HW), bool(out_is_logits), bool(target_is_logits))
        neg, pos, num = res[0], res[1], res[2]
        return torch.where(num == 0, -neg, -(pos + neg) / torch.clamp(num, min=1.0)).float()

    @staticmethod
    def backward(ctx, g):
        ko, kt, ind, mask8, cat, res = ctx.saved_tensors
        m_out, (osb, osc, oscell), (tsb, tsc, tscell), (B, C, HW), ol, tl = ctx.meta
        d, dsb, dsc, dscell = _grad_buffer(m_out, ko)
        up = g.float().contiguous()
        _lib.check(_lib.load().s2d_focal_loss_bwd(ko.data_ptr(), osb, osc, oscell, int(ol), kt.data_ptr(), tsb, tsc, tscell,
                                                  int(tl), B, C, HW, ind.data_ptr(), mask8.data_ptr(), cat.data_ptr(),
                                                  ind.shape[1], res.data_ptr(), up.data_ptr(), d.data_ptr(), dsb, dsc, dscell,
                                                  ops._stream()), "s2d_focal_loss_bwd")
        return d, None, None, None, None, None, None, None


class _Reg(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t, m_out, mask8, ind, target_rows, target_map, squared):
        m = _rewrap(m_out, t)
        ko, po, osb, osc, oscell, B, D, HW = _view(m)
        res = torch.empty((17,), dtype=torch.float64, device=ko.device)
        ws, n = _ws(ko.device)
        if target_map is not None:
            kt, pt, tsb, tsc, tscell, _, Dt, _ = _view(target_map)
            assert Dt == D
            tr = None
        else:
            kt, pt, tsb, tsc, tscell = None, None, 0, 0, 0
            tr = target_rows.contiguous().float()
            assert tuple(tr.shape) == (B, ind.shape[1], D)
        _lib.check(_lib.load().s2d_gather_reg_loss(po, osb, osc, oscell, None if tr is None else tr.data_ptr(), pt, tsb,
                                                   tsc, tscell, B, ind.shape[1], D, int(squared), ind.data_ptr(),
                                                   mask8.data_ptr(), res.data_ptr(), ws.data_ptr(), n, ops._stream()),
                   "s2d_gather_reg_loss")
        ctx.save_for_backward(ko, kt, tr, ind, mask8, res)
        ctx.meta = (m_out, (osb, osc, oscell), (tsb, tsc, tscell), (B, D), bool(squared))
        return (res[:D] / (res[16] + 1e-4)).float()

    @staticmethod
    def backward(ctx, g):
        ko, kt, tr, ind, mask8, res = ctx.saved_tensors
        m_out, (osb, osc, oscell), (tsb, tsc, tscell), (B, D), squared = ctx.meta
        d, dsb, dsc, dscell = _grad_buffer(m_out, ko)
        up = g.float().contiguous()
        _lib.check(_lib.load().s2d_gather_reg_loss_bwd(ko.data_ptr(), osb, osc, oscell, None if tr is None else tr.data_ptr(),
                                                       None if kt is None else kt.data_ptr(), tsb, tsc, tscell, B,
                                                       ind.shape[1], D, int(squared), ind.data_ptr(), mask8.data_ptr(),
                                                       res.data_ptr(), up.data_ptr(), d.data_ptr(), d.numel(), dsb, dsc,
                                                       dscell, ops._stream()), "s2d_gather_reg_loss_bwd")
        return d, None, None, None, None, None, None


def _needs_grad(m):
    return torch.is_grad_enabled() and _tensor_of(m).requires_grad


def _peaks(ind, mask, cat=None):
    ind = ind.contiguous().long()
    mask8 = mask.contiguous().to(torch.uint8)
    return ind, mask8, (None if cat is None else cat.contiguous().long())


def fastfocalloss(out, target, ind, mask, cat, out_is_logits=False, target_is_logits=False):
    """FastFocalLoss.forward.  ``out_is_logits`` fuses CenterHead._sigmoid, ``target_is_logits`` fuses F.sigmoid."""
    ind, mask8, cat = _peaks(ind, mask, cat)
    if _needs_grad(out):
        meta = Rows(None, out.B, out.HW) if isinstance(out, Rows) else None
        return _Focal.apply(_tensor_of(out), meta, target, ind, mask8, cat, out_is_logits, target_is_logits)
    ko, po, osb, osc, oscell, B, C, HW = _view(out)
    kt, pt, tsb, tsc, tscell, Bt, Ct, HWt = _view(target)
    assert (B, C, HW) == (Bt, Ct, HWt)
    res = torch.empty((3,), dtype=torch.float64, device=ko.device)
    ws, n = _ws(ko.device)
    _lib.check(_lib.load().s2d_focal_loss(po, osb, osc, oscell, int(out_is_logits), pt, tsb, tsc, tscell,
                                          int(target_is_logits), B, C, HW, ind.data_ptr(), mask8.data_ptr(), cat.data_ptr(),
                                          ind.shape[1], res.data_ptr(), ws.data_ptr(), n, ops._stream()), "s2d_focal_loss")
    neg, pos, num = res[0], res[1], res[2]
    return torch.where(num == 0, -neg, -(pos + neg) / torch.clamp(num, min=1.0)).float()


def _reg(output, mask, ind, target_rows, target_map, squared):
    ind, mask8, _ = _peaks(ind, mask)
    if _needs_grad(output):
        meta = Rows(None, output.B, output.HW) if isinstance(output, Rows) else None
        return _Reg.apply(_tensor_of(output), meta, mask8, ind, target_rows, target_map, squared)
    ko, po, osb, osc, oscell, B, D, HW = _view(output)
    res = torch.empty((17,), dtype=torch.float64, device=ko.device)
    ws, n = _ws(ko.device)
    if target_map is not None:
        kt, pt, tsb, tsc, tscell, _, Dt, _ = _view(target_map)
        assert Dt == D
        tr = None
    else:
        kt, pt, tsb, tsc, tscell = None, None, 0, 0, 0
        tr = target_rows.contiguous().float()
        assert tuple(tr.shape) == (B, ind.shape[1], D)
    _lib.check(_lib.load().s2d_gather_reg_loss(po, osb, osc, oscell, None if tr is None else tr.data_ptr(), pt, tsb, tsc,
                                               tscell, B, ind.shape[1], D, int(squared), ind.data_ptr(), mask8.data_ptr(),
                                               res.data_ptr(), ws.data_ptr(), n, ops._stream()), "s2d_gather_reg_loss")
    return (res[:D] / (res[16] + 1e-4)).float()


def distill_reg_loss(output, target, mask, ind):
    """trainer.py:68-76: squared error between the student's and the teacher's anno_box maps gathered at ``ind``."""
    return _reg(output, mask, ind, None, target, True)


class RegLoss(nn.Module):
    def forward(self, output, mask, ind, target):
        return _reg(output, mask, ind, target, None, False)


class FastFocalLoss(nn.Module):
    def forward(self, out, target, ind, mask, cat):
        return fastfocalloss(out, target, ind, mask, cat)
