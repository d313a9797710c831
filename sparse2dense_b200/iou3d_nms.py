"""The reference's native NMS operator API on top of libs2d_b200.so.

``nms_gpu`` has the signature and return convention of ``iou3d_nms_cuda.nms_gpu`` (det3d/ops/iou3d_nms/src/
iou3d_nms_api.cpp:11-17, iou3d_nms.cpp:90-136) and ``rotate_nms_pcdet`` those of det3d/core/bbox/box_torch_ops.py:449-470,
so reference-style callers keep working.  Errors are Python exceptions, never ``exit(-1)``.
"""
import torch

from . import ops


def nms_gpu(boxes, keep, nms_overlap_thresh):
    """boxes: cuda f32 [N,7] contiguous, sorted by descending score; keep: CPU int64 [N] (written); returns the count."""
    if not boxes.is_cuda:
        raise ValueError("boxes must be a CUDA tensor")
    if not boxes.is_contiguous() or not keep.is_contiguous():
        raise ValueError("boxes and keep must be contiguous")
    k, n = ops.nms_sorted(boxes, nms_overlap_thresh)
    num = int(n.item())                                   # the reference API is synchronous too
    keep[:num] = k[:num].to(torch.int64).cpu()
    return num


def rotate_nms_pcdet(boxes, scores, thresh, pre_maxsize=None, post_max_size=None):
    order = scores.sort(0, descending=True)[1]
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    boxes = boxes[order].contiguous()
    k, n = ops.nms_sorted(boxes, thresh)
    selected = order[k[: int(n.item())].long()].contiguous()
    if post_max_size is not None:
        selected = selected[:post_max_size]
    return selected
