"""CPU restatement (plain torch.nn.functional, fp32) of the eval-mode forward of the reference's dense BEV
stage.  TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

  s2d_rpn_forward     det3d/models/necks/rpn.py:300-337 (S2D_RPN.forward, eval: PCR branch skipped :324-325),
                      modules defined :186-259; block / deblock pyramid :126-145,72-115
  center_head_forward det3d/models/bbox_heads/center_head.py:236-244 (CenterHead.forward), SepHead :65-110

PINNED: checked bit-for-bit-level (<= 1e-5) against the reference's own torch modules imported through the
shim of SURVEY.md App. G in the authoring container (tests/golden/make_golden.py -> neck_head_*.npz).
``state`` is a reference-format state dict (key -> torch tensor or numpy array).
"""
import numpy as np
import torch
import torch.nn.functional as F


def _t(state, key):
    v = state[key]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v))


def _bn(state, x, prefix, eps):
    return F.batch_norm(x, _t(state, prefix + ".running_mean"), _t(state, prefix + ".running_var"),
                        _t(state, prefix + ".weight"), _t(state, prefix + ".bias"), False, 0.0, eps)


def _conv(state, x, prefix, stride=1, padding=0):
    b = prefix + ".bias"
    return F.conv2d(x, _t(state, prefix + ".weight"), _t(state, b) if b in state else None, stride, padding)


def _tconv(state, x, prefix, stride, padding):
    b = prefix + ".bias"
    return F.conv_transpose2d(x, _t(state, prefix + ".weight"), _t(state, b) if b in state else None, stride, padding)


def _rpn_block(state, x, i, n_layers, stride, eps):
    """rpn.py:126-145: ZeroPad2d(1)+Conv3x3(stride)+BN+ReLU, then n_layers x (Conv3x3+BN), ReLU between."""
    p = f"blocks.{i}"
    x = F.relu(_bn(state, _conv(state, F.pad(x, (1, 1, 1, 1)), f"{p}.1", stride, 0), f"{p}.2", eps))
    idx = 4
    for j in range(n_layers):
        x = _bn(state, _conv(state, x, f"{p}.{idx}", 1, 1), f"{p}.{idx + 1}", eps)
        idx += 2
        if j < n_layers - 1:
            x = F.relu(x)
            idx += 1
    return x


def s2d_rpn_forward(state, x, layer_nums=(5, 5), ds_layer_strides=(1, 2), us_layer_strides=(1, 2), rpn_eps=1e-3):
    """-> (x [B,512,188,188], F_S_a, F_S_b).  BatchNorm2d of the S2D module uses the torch default eps 1e-5
    (rpn.py:188-248), the RPN pyramid the norm_cfg eps 1e-3 (rpn.py:48)."""
    e = 1e-5
    g = F.gelu
    y_1 = g(_bn(state, _conv(state, x, "encoder_1.0", 2, 0), "encoder_1.1", e))
    y_1 = g(_bn(state, _conv(state, y_1, "encoder_1.3", 1, 1), "encoder_1.4", e))
    y_2 = g(_bn(state, _conv(state, y_1, "encoder_2.0", 2, 1), "encoder_2.1", e))
    y_2 = g(_bn(state, _conv(state, y_2, "encoder_2.3", 1, 1), "encoder_2.4", e))

    def convnext(att, p):
        t = F.conv2d(att, _t(state, p + ".0.weight"), _t(state, p + ".0.bias"), 1, 3, 1, att.shape[1])
        t = F.layer_norm(t, tuple(t.shape[1:]), _t(state, p + ".1.weight"), _t(state, p + ".1.bias"), 1e-6)
        t = g(_conv(state, t, p + ".2"))
        return _conv(state, t, p + ".4")

    att = convnext(y_2, "convnext_block_1") + y_2
    att = convnext(att, "convnext_block_2") + att
    att = g(convnext(att, "convnext_block_3") + att)
    d1 = g(_bn(state, _tconv(state, att, "decoder_1.0", 2, 1), "decoder_1.1", e))
    y_3 = torch.cat([d1, y_1], 1)
    t = g(_bn(state, _conv(state, y_3, "decoder_2.0", 1, 1), "decoder_2.1", e))
    F_S_b = g(_bn(state, _tconv(state, t, "decoder_2.3", 2, 1), "decoder_2.4", e))
    F_S_a = g(_bn(state, _conv(state, F_S_b, "fusion_dense.0"), "fusion_dense.1", e)) + \
        g(_bn(state, _conv(state, x, "fusion_sparse.0"), "fusion_sparse.1", e))

    ups = []
    start = len(layer_nums) - len(us_layer_strides)
    h = F_S_a
    for i in range(len(layer_nums)):
        h = _rpn_block(state, h, i, layer_nums[i], ds_layer_strides[i], rpn_eps)        # no outer ReLU (rpn.py:327-331)
        d = i - start
        if d >= 0:
            s = us_layer_strides[d]
            p = f"deblocks.{d}"
            if s > 1:
                u = _tconv(state, h, p + ".0", s, 0)
            else:
                k = int(round(1 / s))
                u = _conv(state, h, p + ".0", k, 0)
            ups.append(F.relu(_bn(state, u, p + ".1", rpn_eps)))
    return torch.cat(ups, 1), F_S_a, F_S_b


def center_head_forward(state, x, heads=("reg", "height", "dim", "rot", "hm"), n_tasks=1):
    """-> list of dict head -> [B,classes,H,W] (center_head.py:236-244; SepHead with bn=True, final_kernel=3)."""
    e = 1e-5
    s = F.relu(_bn(state, _conv(state, x, "shared_conv.0", 1, 1), "shared_conv.1", e))
    ret = []
    for t in range(n_tasks):
        d = {}
        for h in heads:
            p = f"tasks.{t}.{h}"
            y = F.relu(_bn(state, _conv(state, s, p + ".0", 1, 1), p + ".1", e))
            d[h] = _conv(state, y, p + ".3", 1, 1)
        ret.append(d)
    return ret
