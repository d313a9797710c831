"""CPU restatement of the anchor head's eval path (MultiGroupHead, det3d/models/bbox_heads/mg_head.py:199-231,528-533,
697-1086 with use_multi_class_nms = False, use_rotate_nms = True).  TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

PINNED to the reference's own ``Head.forward``, ``GroundBox3dCoderTorch.decode_torch`` and ``create_anchors_3d_range`` imported
through the shim (tests/golden/make_golden.py mg -> tests/golden/mg_head.npz) up to and including the score / top-k selection.
PARITY UNPINNED for the NMS stage: the reference calls ``rotate_nms_cc`` (boost::geometry polygons, a compiled extension that
cannot be built here); this restatement uses the iou3d polygon IoU of ref_ops (itself pinned to iou3d_cpu.cpp)."""
import numpy as np
import torch
import torch.nn.functional as F

from . import ref_ops as R


def head_forward(state, x, task=0, use_dir=True):
    """x NCHW -> dict of NHWC tensors (mg_head.py:221-231)."""
    t = lambda k: torch.as_tensor(state[f"tasks.{task}.{k}"])
    out = {"box_preds": F.conv2d(x, t("conv_box.weight"), t("conv_box.bias")).permute(0, 2, 3, 1).contiguous(),
           "cls_preds": F.conv2d(x, t("conv_cls.weight"), t("conv_cls.bias")).permute(0, 2, 3, 1).contiguous()}
    if use_dir:
        out["dir_cls_preds"] = F.conv2d(x, t("conv_dir.weight"), t("conv_dir.bias")).permute(0, 2, 3, 1).contiguous()
    return out


def decode(box_encodings, anchors):
    """second_box_decode (box_torch_ops.py:87-160), 7-dim, log sizes, additive yaw."""
    xa, ya, za, wa, la, ha, ra = np.split(anchors, 7, axis=-1)
    xt, yt, zt, wt, lt, ht, rt = np.split(box_encodings, 7, axis=-1)
    diagonal = np.sqrt(la ** 2 + wa ** 2)
    return np.concatenate([xt * diagonal + xa, yt * diagonal + ya, zt * ha + za, np.exp(wt) * wa, np.exp(lt) * la, np.exp(ht) * ha,
                           rt + ra], -1).astype(np.float32)


def select(preds, anchors, num_cls, score_threshold, pre_max):
    """-> (boxes [k,7], scores [k], labels [k], dir_labels [k]) sorted by descending score (torch.topk order)."""
    box = preds["box_preds"].reshape(-1, 7).numpy()
    cls = preds["cls_preds"].reshape(-1, num_cls)
    reg = decode(box, anchors)
    scores = torch.sigmoid(cls.float())
    top_scores, top_labels = torch.max(scores, dim=-1)
    keep = top_scores >= score_threshold
    dl = torch.max(preds["dir_cls_preds"].reshape(-1, 2), dim=-1)[1][keep]
    bx, sc, lb = torch.from_numpy(reg)[keep], top_scores[keep], top_labels[keep]
    sc, idx = torch.topk(sc, k=min(pre_max, sc.shape[0]))
    return bx[idx].numpy(), sc.numpy(), lb[idx].numpy(), dl[idx].numpy()


def predict(preds, anchors, num_cls, score_threshold, pre_max, post_max, iou_threshold, post_range, direction_offset=0.0):
    bx, sc, lb, dl = select(preds, anchors, num_cls, score_threshold, pre_max)
    keep = R.nms_sorted(bx, iou_threshold)[0][:post_max]
    bx, sc, lb, dl = bx[keep].copy(), sc[keep], lb[keep], dl[keep]
    opp = ((bx[:, -1] - direction_offset) > 0) ^ dl.astype(bool)
    bx[:, -1] += np.where(opp, np.float32(np.pi), np.float32(0.0))
    r = np.asarray(post_range, np.float32)
    m = (bx[:, :3] >= r[:3]).all(1) & (bx[:, :3] <= r[3:]).all(1)
    return bx[m], sc[m], lb[m]
