"""CenterPoint head (det3d/models/bbox_heads/center_head.py:65-110 ``SepHead``, :167-244
``CenterHead.__init__/forward``, :250-291 ``loss``).  Same constructor arguments and state-dict keys
(``shared_conv.0.weight``, ``tasks.0.hm.3.bias`` …).  The shared 3x3 conv and the five 64->64 branch convs
(fused into one 64->320 launch) run on the tcgen05 gather-GEMM; the tiny output convs (Cout <= 3) on the fp32
kernel.  ``predict`` (:293-495: decode, masks, top-4096, rotated NMS, top-500) runs as four launches for the whole
batch on the device (csrc/detect.cu); ``loss`` runs the reduction kernels of csrc/losses.cu and is differentiable; in
training mode every conv is its own launch on the autograd operators of ``autograd.py`` (batch-statistics BatchNorm)."""
import copy
import math
import logging

import torch
from torch import nn

from . import ops
from .dense import ACT_NONE, ACT_RELU, DenseOps, conv_rows, conv_table, fold_bn, to_nchw, to_rows
from .registry import HEADS


class SepHead(nn.Module):
    def __init__(self, in_channels, heads, head_conv=64, final_kernel=1, bn=False, init_bias=-2.19, **kwargs):
        super(SepHead, self).__init__(**kwargs)
        self.heads = heads
        for head in self.heads:
            classes, num_conv = self.heads[head]
            mods = []
            for _ in range(num_conv - 1):
                mods.append(nn.Conv2d(in_channels, head_conv, kernel_size=final_kernel, stride=1,
                                      padding=final_kernel // 2, bias=True))
                if bn:
                    mods.append(nn.BatchNorm2d(head_conv))
                mods.append(nn.ReLU())
            mods.append(nn.Conv2d(head_conv, classes, kernel_size=final_kernel, stride=1, padding=final_kernel // 2,
                                  bias=True))
            fc = nn.Sequential(*mods)
            if "hm" in head:
                fc[-1].bias.data.fill_(init_bias)
            else:
                for m in fc.modules():
                    if isinstance(m, nn.Conv2d):
                        nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                        if m.bias is not None:
                            nn.init.constant_(m.bias, 0)
            self.__setattr__(head, fc)


@HEADS.register_module
class CenterHead(nn.Module):
    def __init__(self, in_channels=[128, ], tasks=[], dataset="nuscenes", weight=0.25, code_weights=[],
                 common_heads=dict(), logger=None, init_bias=-2.19, share_conv_channel=64, num_hm_conv=2,
                 dcn_head=False):
        super(CenterHead, self).__init__()
        if dcn_head:
            raise NotImplementedError("dcn_head=True is not used by the Waymo configs and is out of scope")
        num_classes = [len(t["class_names"]) for t in tasks]
        self.class_names = [t["class_names"] for t in tasks]
        self.code_weights = code_weights
        self.weight = weight
        self.dataset = dataset
        self.in_channels = in_channels
        self.num_classes = num_classes
        self.box_n_dim = 9 if "vel" in common_heads else 7
        self.use_direction_classifier = False
        self.logger = logger or logging.getLogger("CenterHead")
        self.shared_conv = nn.Sequential(
            nn.Conv2d(in_channels, share_conv_channel, kernel_size=3, padding=1, bias=True),
            nn.BatchNorm2d(share_conv_channel), nn.ReLU(inplace=True))
        self.tasks = nn.ModuleList()
        for num_cls in num_classes:
            heads = copy.deepcopy(common_heads)
            heads.update(dict(hm=(num_cls, num_hm_conv)))
            self.tasks.append(SepHead(share_conv_channel, heads, bn=True, init_bias=init_bias, final_kernel=3))
        self._dense = DenseOps()

    def set_precision(self, precision):
        self._dense = DenseOps(precision)

    def forward_rows(self, x, B, H, W):
        """x rows [B*H*W, C] -> list (per task) of dict head -> rows [B*H*W, classes]."""
        D = self._dense
        D.training = self.training      # training: batch-statistics BN + autograd, one launch per conv (autograd.py)
        s, _, _ = D.conv("shared_conv.0", x, B, H, W, self.shared_conv[0], self.shared_conv[1], ACT_RELU)
        tbl, _, _ = conv_table(x.device, B, H, W, 3, 1, 1)
        n = B * H * W
        ret = []
        for ti, task in enumerate(self.tasks):
            names = list(task.heads)
            two_conv = (not self.training) and all(task.heads[h][1] == 2 for h in names)
            out = {}
            if two_conv:
                # the first conv of every branch reads the same 64-channel map: one 64 -> 64*len(heads) launch
                firsts = [getattr(task, h)[0] for h in names]
                bns = [getattr(task, h)[1] for h in names]

                def build():
                    w = torch.cat([c.weight.detach().float() for c in firsts], 0)               # [320,64,3,3]
                    kio = w.permute(2, 3, 1, 0).reshape(9, w.shape[1], w.shape[0]).contiguous()
                    packed = {}                                   # effective precision -> image (filled by conv_rows)
                    sc, sh = zip(*[fold_bn(b, c.bias, c.weight.shape[0], c.weight.device) for c, b in zip(firsts, bns)])
                    return kio, packed, torch.cat(sc).contiguous(), torch.cat(sh).contiguous()
                src = [c.weight for c in firsts] + [c.bias for c in firsts] + \
                      [t for b in bns for t in (b.weight, b.bias, b.running_mean, b.running_var)]
                kio, packed, sc, sh = D.cache.get(("heads", ti, D.precision), src, build)
                mid = conv_rows(s, kio, tbl, n, sc, sh, ACT_RELU, precision=D.precision, packed=packed)
                hc = firsts[0].weight.shape[0]
                # the five output convs (Cout 2/1/3/2/3) as ONE block-diagonal 320 -> 32 launch on the tensor cores
                lasts = [getattr(task, h)[-1] for h in names]
                classes = [c.weight.shape[0] for c in lasts]
                cpad = -(-sum(classes) // 32) * 32

                def build_last():
                    kio = torch.zeros((9, hc * len(names), cpad), dtype=torch.float32, device=s.device)
                    bias = torch.zeros((cpad,), dtype=torch.float32, device=s.device)
                    off = 0
                    for hi, c in enumerate(lasts):
                        w = c.weight.detach().float().permute(2, 3, 1, 0).reshape(9, hc, c.weight.shape[0])
                        kio[:, hi * hc:(hi + 1) * hc, off:off + c.weight.shape[0]] = w
                        bias[off:off + c.weight.shape[0]] = c.bias.detach().float()
                        off += c.weight.shape[0]
                    return kio, {}, bias
                kio_l, packed_l, bias_l = D.cache.get(("lasts", ti, D.precision),
                                                      [c.weight for c in lasts] + [c.bias for c in lasts], build_last)
                y = conv_rows(mid, kio_l, tbl, n, None, bias_l, ACT_NONE, precision=D.precision, packed=packed_l)
                off = 0
                for h, c in zip(names, classes):
                    out[h] = y[:, off:off + c]
                    off += c
            else:
                for h in names:
                    y = s
                    mods = list(getattr(task, h))
                    j = 0
                    while j < len(mods):
                        m = mods[j]
                        if isinstance(m, nn.Conv2d):
                            bn = mods[j + 1] if j + 1 < len(mods) and isinstance(mods[j + 1], nn.modules.batchnorm._BatchNorm) else None
                            k = j + 1 + int(bn is not None)
                            relu = k < len(mods) and isinstance(mods[k], nn.ReLU)
                            y, _, _ = D.conv(f"tasks.{ti}.{h}.{j}", y, B, H, W, m, bn, ACT_RELU if relu else ACT_NONE)
                            j = k + int(relu)
                        else:
                            j += 1
                    out[h] = y
            ret.append(out)
        return ret

    # -- predict (center_head.py:293-495) ---------------------------------------------------------
    @staticmethod
    def _cfg(cfg, name, default=None):
        if isinstance(cfg, dict):
            return cfg.get(name, default)
        return getattr(cfg, name, default)

    @torch.no_grad()
    def predict_rows(self, preds_rows, B, H, W, test_cfg, metadata=None):
        """preds_rows: list (per task) of dict head -> rows [B*H*W, c] (what ``forward_rows`` returns).
        Returns the reference's list (per sample) of dict(box3d_lidar, scores, label_preds, metadata)."""
        g = self._cfg
        if g(test_cfg, "double_flip", False):
            raise NotImplementedError("double_flip testing is not used by the Waymo configs and is not built")
        if g(test_cfg, "circular_nms", False) or g(test_cfg, "per_class_nms", False):
            raise NotImplementedError("only rotated NMS (nms.use_rotate_nms) is built")
        per_task = self.select_rows(preds_rows, B, H, W, test_cfg)
        flat = ops.read_ints(torch.stack([t[4] for t in per_task]).reshape(-1))   # the one host sync
        counts = [flat[i * B:(i + 1) * B] for i in range(len(per_task))]          # [task][sample]
        meta = metadata if metadata else [None] * B
        ret = []
        for i in range(B):
            bx, sc, lb, flag = [], [], [], 0
            for ti, (ob, os_, ol, _, _) in enumerate(per_task):
                n = counts[ti][i]
                bx.append(ob[i, :n]); sc.append(os_[i, :n]); lb.append(ol[i, :n].long() + flag)
                flag += self.num_classes[ti]
            ret.append(dict(box3d_lidar=torch.cat(bx), scores=torch.cat(sc), label_preds=torch.cat(lb), metadata=meta[i]))
        return ret

    @torch.no_grad()
    def select_rows(self, preds_rows, B, H, W, test_cfg):
        """Device part of predict: per task (boxes [B,post_max,7], scores, labels i32, cells i32, counts i32 [B])."""
        g = self._cfg
        nms = g(test_cfg, "nms")
        rng = list(g(test_cfg, "post_center_limit_range"))
        assert len(rng) == 6, "post_center_limit_range must have 6 entries"
        per_task = []
        for preds in preds_rows:
            boxes, scores, labels, keys = ops.centerhead_decode(
                preds, B, H, W, g(test_cfg, "out_size_factor"), g(test_cfg, "voxel_size"), g(test_cfg, "pc_range"),
                g(test_cfg, "score_threshold"), rng)
            per_task.append(ops.centerhead_select(keys, boxes, scores, labels, B, H * W, int(g(nms, "nms_pre_max_size")),
                                                  float(g(nms, "nms_iou_threshold")), int(g(nms, "nms_post_max_size"))))
        return per_task

    @torch.no_grad()
    def predict(self, example, preds_dicts, test_cfg, **kwargs):
        """Reference signature: preds_dicts = list of dict head -> NCHW map (``forward`` output)."""
        B, _, H, W = preds_dicts[0]["hm"].shape
        rows = [{h: to_rows(v.contiguous()) for h, v in d.items()} for d in preds_dicts]
        meta = example.get("metadata") if isinstance(example, dict) else None
        return self.predict_rows(rows, B, H, W, test_cfg, meta if meta else None)

    # -- losses (center_head.py:250-291) --------------------------------------------------------------
    def loss(self, example, preds_dicts, **kwargs):
        """Same return structure as the reference (dict of per-task lists).  ``preds_dicts``: list of dict head -> NCHW
        logits map or ``losses.Rows`` view.  Differentiable w.r.t. the prediction maps (losses.py)."""
        from collections import defaultdict
        from . import losses as L
        crit_reg = L.RegLoss()
        merged = defaultdict(list)
        for task_id, preds in enumerate(preds_dicts):
            hm_loss = L.fastfocalloss(preds["hm"], example["hm"][task_id], example["ind"][task_id],
                                      example["mask"][task_id], example["cat"][task_id], out_is_logits=True)
            target_box = example["anno_box"][task_id]
            names = ["reg", "height", "dim"] + (["vel"] if "vel" in preds else []) + ["rot"]
            if "vel" not in preds:
                target_box = target_box[..., [0, 1, 2, 3, 4, 5, -2, -1]]                 # remove the velocity target
            parts, off = [], 0
            mask, ind = example["mask"][task_id], example["ind"][task_id]
            for n in names:                                  # anno_box = cat(reg, height, dim, [vel], rot): per-head gathers
                c = preds[n].rows.shape[1] if isinstance(preds[n], L.Rows) else preds[n].shape[1]
                parts.append(crit_reg(preds[n], mask, ind, target_box[..., off:off + c].contiguous()))
                off += c
            box_loss = torch.cat(parts)
            loc_loss = (box_loss * box_loss.new_tensor(self.code_weights)).sum()
            ret = {"loss": hm_loss + self.weight * loc_loss, "hm_loss": hm_loss.detach(), "loc_loss": loc_loss,
                   "loc_loss_elem": box_loss.detach(), "num_positive": mask.float().sum()}
            for k, v in ret.items():
                merged[k].append(v)
        return merged

    def forward(self, x, *kwargs):
        """x NCHW [B,C,H,W] -> list of dicts of NCHW maps, like center_head.py:236-244."""
        B, _, H, W = x.shape
        rows = self.forward_rows(to_rows(x), B, H, W)
        return [{h: to_nchw(v, B, H, W) for h, v in d.items()} for d in rows]


class Head(nn.Module):
    """mg_head.py:199-231: the 1x1 box / class / direction convolutions of one task."""

    def __init__(self, num_input, num_pred, num_cls, use_dir=False, num_dir=0, header=True, name="", focal_loss_init=False,
                 **kwargs):
        super(Head, self).__init__(**kwargs)
        self.use_dir = use_dir
        self.conv_box = nn.Conv2d(num_input, num_pred, 1)
        self.conv_cls = nn.Conv2d(num_input, num_cls, 1)
        if self.use_dir:
            self.conv_dir = nn.Conv2d(num_input, num_dir, 1)


@HEADS.register_module
class MultiGroupHead(nn.Module):
    """Anchor-based SECOND head (det3d/models/bbox_heads/mg_head.py:386-533 ``__init__/forward``, :697-1086 ``predict``):
    same constructor arguments and state-dict keys (``tasks.0.conv_box.weight`` ...).  Inference only.

    forward: the three 1x1 convolutions of a task are ONE gather-GEMM launch (weights concatenated on Cout, padded to a
    multiple of 16); predict: box decoding against the anchors (``box_coder.decode_torch``), sigmoid scores, per-anchor
    best class, score threshold, top ``nms_pre_max_size``, rotated NMS, direction flip and the centre-range mask.
    The rotated NMS is the library's exact polygon-clipping IoU (csrc/detect.cu, pinned to the reference's iou3d_cpu.cpp); the
    reference routes this head through ``box_torch_ops.rotate_nms`` -> ``rotate_nms_cc`` (boost::geometry, not buildable
    here): that one stage is therefore PARITY-UNPINNED.  ``use_multi_class_nms`` and the non-rotated
    NMS are not built (the Waymo SECOND configs use neither)."""

    def __init__(self, mode="3d", in_channels=[128, ], norm_cfg=None, tasks=[], weights=[], num_classes=[1, ], box_coder=None,
                 with_cls=True, with_reg=True, reg_class_agnostic=False, encode_background_as_zeros=True,
                 loss_norm=None, loss_cls=None, use_sigmoid_score=True, loss_bbox=None, encode_rad_error_by_sin=True,
                 loss_aux=None, direction_offset=0.0, name="rpn", logger=None):
        super(MultiGroupHead, self).__init__()
        assert with_cls or with_reg
        assert box_coder is not None, "MultiGroupHead needs a box_coder (det3d.builder.build_box_coder)"
        num_classes = [len(t["class_names"]) for t in tasks]
        self.class_names = [t["class_names"] for t in tasks]
        self.num_anchor_per_locs = [2 * n for n in num_classes]
        self.box_coder = box_coder
        self.in_channels, self.num_classes = in_channels, num_classes
        self.encode_background_as_zeros, self.use_sigmoid_score = encode_background_as_zeros, use_sigmoid_score
        self.box_n_dim, self.anchor_dim = box_coder.code_size, box_coder.n_dim
        self.use_direction_classifier = loss_aux is not None
        self.direction_offset = direction_offset
        self.bev_only = mode == "bev"
        self.loss_cfg = dict(loss_norm=loss_norm, loss_cls=loss_cls, loss_bbox=loss_bbox, loss_aux=loss_aux, weights=weights,
                             encode_rad_error_by_sin=encode_rad_error_by_sin)
        self.logger = logger or logging.getLogger("MultiGroupHead")
        self.tasks = nn.ModuleList()
        for num_c, num_a in zip(num_classes, self.num_anchor_per_locs):
            num_cls = num_a * num_c if encode_background_as_zeros else num_a * (num_c + 1)
            num_pred = num_a * (self.box_n_dim - 2) if self.bev_only else num_a * self.box_n_dim
            self.tasks.append(Head(in_channels, num_pred, num_cls, use_dir=self.use_direction_classifier,
                                   num_dir=num_a * 2 if self.use_direction_classifier else None, header=False))
        self._dense = DenseOps()

    def set_precision(self, precision):
        self._dense = DenseOps(precision)

    def init_weights(self, pretrained=None):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    def forward_rows(self, x, B, H, W):
        """x rows [B*H*W, C] -> list (per task) of dict(box_preds, cls_preds[, dir_cls_preds]) as rows [B*H*W, c]
        (= the reference's NHWC ``permute(0, 2, 3, 1)`` tensors flattened over B, H, W)."""
        if self.training:
            raise NotImplementedError("MultiGroupHead: the training branch (anchor losses) is not built")
        D = self._dense
        tbl, _, _ = conv_table(x.device, B, H, W, 1, 1, 0)
        n = B * H * W
        ret = []
        for ti, task in enumerate(self.tasks):
            convs = [("box_preds", task.conv_box), ("cls_preds", task.conv_cls)]
            if task.use_dir:
                convs.append(("dir_cls_preds", task.conv_dir))
            couts = [c.weight.shape[0] for _, c in convs]
            cpad = -(-sum(couts) // 16) * 16

            def build():
                cin = convs[0][1].weight.shape[1]
                kio = torch.zeros((1, cin, cpad), dtype=torch.float32, device=x.device)
                bias = torch.zeros((cpad,), dtype=torch.float32, device=x.device)
                off = 0
                for (_, c), co in zip(convs, couts):
                    kio[0, :, off:off + co] = c.weight.detach().float().reshape(co, cin).t()
                    bias[off:off + co] = c.bias.detach().float()
                    off += co
                return kio, {}, bias
            kio, packed, bias = D.cache.get(("mg", ti, D.precision), [c.weight for _, c in convs] + [c.bias for _, c in convs],
                                            build)
            y = conv_rows(x, kio, tbl, n, None, bias, ACT_NONE, precision=D.precision, packed=packed)
            out, off = {}, 0
            for (name, _), co in zip(convs, couts):
                out[name] = y[:, off:off + co]
                off += co
            ret.append(out)
        return ret

    def forward(self, x):
        """mg_head.py:528-533: NCHW feature map -> list of dicts of NHWC tensors [B, H, W, c]."""
        B, _, H, W = x.shape
        rows = self.forward_rows(to_rows(x.contiguous()), B, H, W)
        return [{k: v.reshape(B, H, W, -1) for k, v in d.items()} for d in rows]

    @torch.no_grad()
    def predict(self, example, preds_dicts, test_cfg, **kwargs):
        """mg_head.py:697-1086 -> list (per sample) of dict(box3d_lidar, scores, label_preds, metadata)."""
        g = CenterHead._cfg
        nms = g(test_cfg, "nms")
        if g(nms, "use_multi_class_nms", False) or not g(nms, "use_rotate_nms", True):
            raise NotImplementedError("MultiGroupHead.predict: only class-agnostic rotated NMS is built")
        assert self.encode_background_as_zeros and self.use_sigmoid_score, "only sigmoid scores without a background column"
        rng = g(test_cfg, "post_center_limit_range")
        thr = float(g(test_cfg, "score_threshold"))
        pre, post, iou = int(g(nms, "nms_pre_max_size")), int(g(nms, "nms_post_max_size")), float(g(nms, "nms_iou_threshold"))
        batch_anchors = example["anchors"]
        B = batch_anchors[0].shape[0]
        meta = example.get("metadata") or [None] * B
        rets = []
        for ti, preds in enumerate(preds_dicts):
            anchors = batch_anchors[ti].view(B, -1, self.anchor_dim)
            box = preds["box_preds"].reshape(B, -1, self.box_n_dim)
            cls = preds["cls_preds"].reshape(B, -1, self.num_classes[ti]).float()
            reg = self.box_coder.decode_torch(box[:, :, :self.box_coder.code_size], anchors).float()
            dirs = preds["dir_cls_preds"].reshape(B, -1, 2) if self.use_direction_classifier else None
            task_out = []
            for b in range(B):
                scores = torch.sigmoid(cls[b])
                top_scores, top_labels = (scores.squeeze(-1), torch.zeros_like(scores[:, 0], dtype=torch.long)) \
                    if scores.shape[1] == 1 else torch.max(scores, dim=-1)
                keep = top_scores >= thr if thr > 0.0 else torch.ones_like(top_scores, dtype=torch.bool)
                bx, sc, lb = reg[b][keep], top_scores[keep], top_labels[keep]
                dl = torch.max(dirs[b], dim=-1)[1][keep] if dirs is not None else None
                if sc.shape[0]:
                    k = min(pre, sc.shape[0])
                    sc_k, idx = torch.topk(sc, k=k)                       # box_torch_ops.rotate_nms :522-526
                    kept, n_kept = ops.nms_sorted(bx[idx][:, [0, 1, 2, 3, 4, 5, -1]].contiguous(), iou)
                    sel = idx[kept[:int(n_kept.item())].long()][:post]
                    bx, sc, lb = bx[sel], sc[sel], lb[sel]
                    if dl is not None:
                        dl = dl[sel]
                        opp = ((bx[..., -1] - self.direction_offset) > 0) ^ dl.bool()
                        bx = bx.clone()
                        bx[..., -1] += torch.where(opp, torch.tensor(math.pi).type_as(bx), torch.tensor(0.0).type_as(bx))
                    if rng is not None and len(rng) > 0:
                        r = torch.tensor(rng, dtype=bx.dtype, device=bx.device)
                        m = (bx[:, :3] >= r[:3]).all(1) & (bx[:, :3] <= r[3:]).all(1)
                        bx, sc, lb = bx[m], sc[m], lb[m]
                task_out.append(dict(box3d_lidar=bx, scores=sc, label_preds=lb, metadata=meta[b]))
            rets.append(task_out)
        out = []
        for b in range(B):
            flag, labels = 0, []
            for ti, nc in enumerate(self.num_classes):
                labels.append(rets[ti][b]["label_preds"] + flag)
                flag += nc
            out.append(dict(box3d_lidar=torch.cat([r[b]["box3d_lidar"] for r in rets]),
                            scores=torch.cat([r[b]["scores"] for r in rets]), label_preds=torch.cat(labels),
                            metadata=rets[0][b]["metadata"]))
        return out

    def loss(self, example, preds_dicts, **kwargs):
        raise NotImplementedError("MultiGroupHead.loss (anchor target losses, mg_head.py:580-695) is not built")
