"""CPU restatement of the full two-stage forward of one scene (BASELINE configs[2]):
points -> voxelize -> reader mean -> SpMiddleResNetFHD -> S2D_RPN -> CenterHead -> decode + rotated NMS -> BEV RoI
features -> RoIHead -> refined boxes.  TEST INFRASTRUCTURE / CPU BASELINE ONLY -- see oracle/__init__.py.

Assembled from the pinned pieces: ref_ops (voxelizer, sparse conv, decode, NMS, second stage), backbone.py
(scn.py:88-185), neck_head.py (rpn.py:300-337, center_head.py:236-244).  Follows
det3d/models/detectors/two_stage.py:154-199 and voxelnet.py:188-265 (eval branch).
"""
import numpy as np
import torch

from . import backbone as OB
from . import neck_head as NH
from . import ref_ops as R

VOXEL, RANGE = (0.1, 0.1, 0.15), (-75.2, -75.2, -2.0, 75.2, 75.2, 4.0)
POST_RANGE = [-80, -80, -10.0, 80, 80, 10.0]


def scene_forward(states, cloud, second_stage=True, timings=None):
    """states = dict(backbone=, neck=, head=, roi=) of numpy / torch state dicts.  -> dict(boxes [k,7], scores, labels)."""
    import time
    t0 = time.perf_counter()
    v, c, n = R.points_to_voxel(cloud, VOXEL, RANGE, 5, True, 150000)
    feats = R.voxel_mean(v, n)
    coors = np.concatenate([np.zeros((len(c), 1), np.int32), c], 1)
    bev, _ = OB.backbone_forward(states["backbone"], feats, coors, 1, (1504, 1504, 40))
    t1 = time.perf_counter()
    with torch.no_grad():
        x, _, _ = NH.s2d_rpn_forward(states["neck"], torch.from_numpy(np.ascontiguousarray(bev)))
        preds = NH.center_head_forward(states["head"], x)[0]
    t2 = time.perf_counter()
    nhwc = {k: np.ascontiguousarray(p.permute(0, 2, 3, 1).numpy()) for k, p in preds.items()}
    boxes, hm = R.centerhead_decode(nhwc, 8, (0.1, 0.1), (-75.2, -75.2))
    det, _ = R.post_processing(boxes[0], hm[0], 0.1, POST_RANGE, 0.7, 4096, 500)
    out = dict(boxes=det["box3d_lidar"], scores=det["scores"], labels=det["label_preds"])
    t3 = time.perf_counter()
    if second_stage and len(out["boxes"]):
        f = R.roi_features(np.ascontiguousarray(x[0].permute(1, 2, 0).numpy()), out["boxes"], (-75.2, -75.2), (0.1, 0.1), 8)
        cls, reg = R.roi_head_forward(states["roi"], f)
        out["boxes"], out["scores"] = R.roi_refine(out["boxes"], out["scores"], cls, reg)
    t4 = time.perf_counter()
    if timings is not None:
        timings.update(backbone=t1 - t0, neck_head=t2 - t1, decode_nms=t3 - t2, second_stage=t4 - t3)
    return out


PP_VOXEL, PP_RANGE = (0.32, 0.32, 6.0), (-74.88, -74.88, -2, 74.88, 74.88, 4.0)


def pillar_scene_forward(states, cloud, second_stage=True, timings=None):
    """BASELINE configs[3] for one scene: pillar voxelize -> PFN -> scatter + S2D -> RPN[3,5,5] -> CenterHead -> decode + NMS ->
    BEV RoI features -> RoIHead (det3d/models/detectors/point_pillars.py:171-251, two_stage.py:154-199).
    states = dict(reader=, backbone=, neck=, head=, roi=)."""
    import time
    from . import pillars as OP
    t0 = time.perf_counter()
    v, c, n = R.points_to_voxel(cloud, PP_VOXEL, PP_RANGE, 20, True, 32000)
    coors = np.concatenate([np.zeros((len(c), 1), np.int32), c], 1)
    with torch.no_grad():
        f = OP.pfn_forward(states["reader"], v, n, coors, PP_VOXEL, PP_RANGE)
        F_S_a, _ = OP.scatter_s2d_forward(states["backbone"], f, coors, 1, 468, 468)
        t1 = time.perf_counter()
        x = OP.rpn_forward(states["neck"], F_S_a, [3, 5, 5], [1, 2, 2], [1, 2, 4])
        preds = NH.center_head_forward(states["head"], x)[0]
    t2 = time.perf_counter()
    nhwc = {k: np.ascontiguousarray(p.permute(0, 2, 3, 1).numpy()) for k, p in preds.items()}
    boxes, hm = R.centerhead_decode(nhwc, 1, (0.32, 0.32), (-74.88, -74.88))
    det, _ = R.post_processing(boxes[0], hm[0], 0.1, POST_RANGE, 0.7, 4096, 500)
    out = dict(boxes=det["box3d_lidar"], scores=det["scores"], labels=det["label_preds"])
    t3 = time.perf_counter()
    if second_stage and len(out["boxes"]):
        fts = R.roi_features(np.ascontiguousarray(x[0].permute(1, 2, 0).numpy()), out["boxes"], (-74.88, -74.88), (0.32, 0.32), 1)
        cls, reg = R.roi_head_forward(states["roi"], fts)
        out["boxes"], out["scores"] = R.roi_refine(out["boxes"], out["scores"], cls, reg)
    t4 = time.perf_counter()
    if timings is not None:
        timings.update(reader_backbone=t1 - t0, neck_head=t2 - t1, decode_nms=t3 - t2, second_stage=t4 - t3)
    return out
