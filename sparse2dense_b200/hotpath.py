"""The hot path as one callable: points -> voxelize (+ fused reader mean) -> SpMiddleResNetFHD -> BEV.

This is what ``bench.py`` and ``__graft_entry__.smoke()`` drive.  It is assembled from the same
registry entries a reference config names (``VoxelFeatureExtractorV3``, ``SpMiddleResNetFHD``;
configs/waymo/voxelnet/two_stage/waymo_centerpoint_voxelnet_two_stage_distill.py:52-56,170-176).
"""
import numpy as np
import torch

from . import ops, registry, synth
from .voxel_generator import VoxelGenerator


class VoxelBackbonePath:
    def __init__(self, state=None, precision=ops.PRECISION_FP32, device="cuda", num_input_features=5):
        if not torch.cuda.is_available():
            raise RuntimeError("VoxelBackbonePath needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device)
        self.generator = VoxelGenerator(synth.WAYMO_VOXEL, synth.WAYMO_RANGE, synth.WAYMO_MAX_POINTS,
                                        synth.WAYMO_MAX_VOXELS)
        self.reader = registry.build_reader(dict(type="VoxelFeatureExtractorV3",
                                                 num_input_features=num_input_features))
        self.backbone = registry.build_backbone(dict(type="SpMiddleResNetFHD",
                                                     num_input_features=num_input_features, ds_factor=8))
        if state is not None:
            self.backbone.load_state_dict({k: torch.as_tensor(v) for k, v in state.items()}, strict=False)
        self.backbone.to(self.device).eval()
        self.backbone.set_precision(precision)
        self.num_input_features = num_input_features
        self.grid = [int(v) for v in self.generator.grid_size]          # (x, y, z) = example["shape"][0]

    @torch.no_grad()
    def forward_points(self, points, scene_offsets):
        """points: cuda f32 [N,5] (scenes concatenated), scene_offsets: host ints [B+1] -> BEV [B,256,188,188]."""
        batch = len(scene_offsets) - 1
        vb = self.generator.generate_batch(points, scene_offsets, want_voxels=False,
                                           mean_channels=self.num_input_features)
        n = vb.n                                                          # host sync #1 (voxel count)
        bev, _ = self.backbone(vb.mean_buffer[:n], vb.coors_buffer[:n], batch, self.grid)
        return bev

    @torch.no_grad()
    def forward_host(self, points_host, scene_offsets, out_host=None):
        """End-to-end form: pinned host points -> device -> hot path -> pinned host BEV."""
        pts = points_host.to(self.device, non_blocking=True)
        bev = self.forward_points(pts, scene_offsets)
        if out_host is None:
            out_host = torch.empty(bev.shape, dtype=bev.dtype, pin_memory=True)
        out_host.copy_(bev, non_blocking=True)
        return out_host

    @torch.no_grad()
    def forward_host_async(self, points_host, scene_offsets, out_host):
        """Pipelined end-to-end form: the device->host copy of the result runs on a copy stream, so it overlaps the
        next call's compute.  Returns the event that marks ``out_host`` complete (wait on it before reading)."""
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        main = torch.cuda.current_stream(self.device)
        pts = points_host.to(self.device, non_blocking=True)
        bev = self.forward_points(pts, scene_offsets)
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(ready)
            out_host.copy_(bev, non_blocking=True)
            bev.record_stream(self._copy_stream)
            done = torch.cuda.Event()
            done.record(self._copy_stream)
        return done


def concat_clouds(clouds, pin=True):
    offs = np.concatenate([[0], np.cumsum([len(c) for c in clouds])]).astype(np.int64).tolist()
    cat = torch.from_numpy(np.ascontiguousarray(np.concatenate(clouds, 0), dtype=np.float32))
    if pin and torch.cuda.is_available():
        cat = cat.pin_memory()
    return cat, offs


class FullForwardPath(VoxelBackbonePath):
    """BASELINE configs[2]: points -> voxelize -> reader -> SpMiddleResNetFHD -> S2D_RPN -> CenterHead -> decode + rotated
    NMS -> BEV RoI features -> RoIHead -> refined boxes, i.e. the ``TwoStageDetector`` of
    configs/waymo/voxelnet/two_stage/waymo_centerpoint_voxelnet_two_stage_distill.py built through the registry.  The
    dense stage runs on NHWC rows end to end; the results stay on the device as padded [B,500,...] tensors + counts."""

    NECK_CFG = dict(type="S2D_RPN", layer_nums=[5, 5], ds_layer_strides=[1, 2], ds_num_filters=[128, 256],
                    us_layer_strides=[1, 2], us_num_filters=[256, 256], num_input_features=256)
    HEAD_CFG = dict(type="CenterHead", in_channels=512,
                    tasks=[dict(num_class=3, class_names=["VEHICLE", "PEDESTRIAN", "CYCLIST"])], dataset="waymo",
                    weight=2, code_weights=[1.0] * 8,
                    common_heads={"reg": (2, 2), "height": (1, 2), "dim": (3, 2), "rot": (2, 2)})
    TEST_CFG = dict(post_center_limit_range=[-80, -80, -10.0, 80, 80, 10.0], max_per_img=4096,
                    nms=dict(use_rotate_nms=True, use_multi_class_nms=False, nms_pre_max_size=4096,
                             nms_post_max_size=500, nms_iou_threshold=0.7),
                    score_threshold=0.1, pc_range=[-75.2, -75.2], out_size_factor=8, voxel_size=[0.1, 0.1])
    ROI_CFG = dict(type="RoIHead", input_channels=512 * 5, code_size=7,
                   model_cfg=dict(CLASS_AGNOSTIC=True, SHARED_FC=[256, 256], CLS_FC=[256, 256], REG_FC=[256, 256],
                                  DP_RATIO=0.3))

    def __init__(self, state=None, neck_state=None, head_state=None, precision=ops.PRECISION_AUTO, device="cuda",
                 second_stage=True):
        super().__init__(state=state, precision=precision, device=device)
        import logging
        from . import second_stage as SS
        self.neck = registry.build_neck(dict(logger=logging.getLogger("RPN"), **self.NECK_CFG))
        self.head = registry.build_head(dict(**self.HEAD_CFG))
        if neck_state is not None:
            self.neck.load_state_dict({k: torch.as_tensor(v) for k, v in neck_state.items()}, strict=False)
        if head_state is not None:
            self.head.load_state_dict({k: torch.as_tensor(v) for k, v in head_state.items()}, strict=False)
        self.neck.to(self.device).eval().set_precision(precision)
        self.head.to(self.device).eval().set_precision(precision)
        self.extractor = SS.BEVFeatureExtractor([-75.2, -75.2], [0.1, 0.1], 8) if second_stage else None
        self.roi_head = registry.build_roi_head(dict(**self.ROI_CFG)).to(self.device).eval() if second_stage else None
        if self.roi_head is not None:
            self.roi_head.set_precision(precision)

    @torch.no_grad()
    def forward_maps(self, points, scene_offsets):
        """Up to the head outputs: list (per task) of dict head -> NCHW map (CenterHead.forward's return value)."""
        from .dense import to_nchw
        preds, _, (batch, Hu, Wu) = self._heads(points, scene_offsets)
        return [{h: to_nchw(v, batch, Hu, Wu) for h, v in d.items()} for d in preds]

    def _heads(self, points, scene_offsets):
        batch = len(scene_offsets) - 1
        vb = self.generator.generate_batch(points, scene_offsets, want_voxels=False,
                                           mean_channels=self.num_input_features)
        n = vb.n
        rows, _ = self.backbone(vb.mean_buffer[:n], vb.coors_buffer[:n], batch, self.grid, as_rows=True)
        H = W = 188
        ups, (Hu, Wu), F_S_a, F_S_b = self.neck.forward_rows(rows, batch, H, W)
        return self.head.forward_rows(ups, batch, Hu, Wu), ups, (batch, Hu, Wu)

    @torch.no_grad()
    def forward_points(self, points, scene_offsets):
        """-> (boxes [B,500,7], scores [B,500], labels i32 [B,500], counts i32 [B]) on the device."""
        from . import _lib
        preds, ups, (batch, Hu, Wu) = self._heads(points, scene_offsets)
        rois, roi_scores, roi_labels, _, n_boxes = self.head.select_rows(preds, batch, Hu, Wu, self.TEST_CFG)[0]
        if self.roi_head is None:
            return rois, roi_scores, roi_labels, n_boxes
        feats = self.extractor.box_features(ups, batch, Hu, Wu, rois, n_boxes, 5)
        rcnn_cls, rcnn_reg = self.roi_head.forward_rows(feats)
        boxes, scores = torch.empty_like(rois), torch.empty_like(roi_scores)
        _lib.check(_lib.load().s2d_roi_refine(rois.data_ptr(), roi_scores.data_ptr(), n_boxes.data_ptr(), batch,
                                              rois.shape[1], rcnn_cls.data_ptr(), rcnn_cls.stride(0), rcnn_reg.data_ptr(),
                                              rcnn_reg.stride(0), boxes.data_ptr(), scores.data_ptr(), ops._stream()),
                   "s2d_roi_refine")
        return boxes, scores, roi_labels, n_boxes
