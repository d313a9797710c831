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

    @classmethod
    def synthetic(cls, precision=ops.PRECISION_AUTO, device="cuda", seed=0, second_stage=True):
        """The configs[2] model with seeded random weights (no checkpoint exists on the build / GPU boxes); the heat-map bias
        is raised to -1 so that random weights still give every scene a full NMS load (4096 candidates, 500 kept)."""
        path = cls(state=synth.backbone_state(seed), precision=precision, device=device, second_stage=second_stage)
        for m, sd in ((path.neck, 11), (path.head, 12)):
            m.load_state_dict({k: torch.as_tensor(v) for k, v in synth.random_module_state(m, sd + 100 * seed).items()},
                              strict=False)
        with torch.no_grad():
            path.head.tasks[0].hm[-1].bias.fill_(-1.0)
        if path.roi_head is not None:
            path.roi_head.load_state_dict({k: torch.as_tensor(v) for k, v in
                                           synth.random_module_state(path.roi_head, 13 + 100 * seed).items()}, strict=False)
            path.roi_head.to(device).eval()
        return path

    def states_numpy(self):
        """State dicts of the four stages as numpy (what the CPU oracle's ``full_forward.scene_forward`` consumes)."""
        f = lambda m: {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
        return dict(backbone=f(self.backbone), neck=f(self.neck), head=f(self.head),
                    roi=f(self.roi_head) if self.roi_head is not None else None)

    @torch.no_grad()
    def forward_host_async(self, points_host, scene_offsets, out_host):
        """Pipelined end-to-end form: pinned host points in, detections out into the pinned host tensors ``out_host`` =
        (boxes [B,500,7], scores [B,500], labels i32 [B,500], counts i32 [B]); the device->host copies run on a copy
        stream under the next call's compute.  Returns the event that marks ``out_host`` complete."""
        if getattr(self, "_copy_stream", None) is None:
            self._copy_stream = torch.cuda.Stream(device=self.device)
        main = torch.cuda.current_stream(self.device)
        pts = points_host.to(self.device, non_blocking=True)
        outs = self.forward_points(pts, scene_offsets)
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(self._copy_stream):
            self._copy_stream.wait_event(ready)
            for h, d in zip(out_host, outs):
                h.copy_(d, non_blocking=True)
                d.record_stream(self._copy_stream)
            done = torch.cuda.Event()
            done.record(self._copy_stream)
        return done

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
    def forward_points(self, points, scene_offsets, return_first_stage=False):
        """-> (boxes [B,500,7], scores [B,500], labels i32 [B,500], counts i32 [B]) on the device (+ the first stage's boxes and
        scores and the BEV feature rows when ``return_first_stage``)."""
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
        if return_first_stage:
            return boxes, scores, roi_labels, n_boxes, rois, roi_scores, ups.view(batch, Hu, Wu, -1)
        return boxes, scores, roi_labels, n_boxes


class PillarForwardPath:
    """BASELINE configs[3]: the CenterPoint-Pillar + S2D student of
    configs/waymo/pp/two_stage/waymo_centerpoint_pp_two_pfn_stride1_two_stage_bev_distill_interval_5.py:17-96 --
    pillar voxelize (0.32 m, 20 points) -> PillarFeatureNet[64,64] -> PointPillarsScatter_S2D -> RPN[3,5,5] -> CenterHead ->
    decode + rotated NMS -> BEV RoI features (stride 1) -> RoIHead; det3d/models/detectors/point_pillars.py:171-251 and
    two_stage.py:154-199 (eval branch).  Built through the registry from the config's own dicts."""

    VOXEL, RANGE = (0.32, 0.32, 6.0), (-74.88, -74.88, -2, 74.88, 74.88, 4.0)
    TEST_CFG = dict(post_center_limit_range=[-80, -80, -10.0, 80, 80, 10.0], max_per_img=4096,
                    nms=dict(use_rotate_nms=True, use_multi_class_nms=False, nms_pre_max_size=4096,
                             nms_post_max_size=500, nms_iou_threshold=0.7),
                    score_threshold=0.1, pc_range=[-74.88, -74.88], out_size_factor=1, voxel_size=[0.32, 0.32])

    def __init__(self, precision=ops.PRECISION_AUTO, device="cuda", seed=0, second_stage=True):
        import logging
        from . import second_stage as SS
        if not torch.cuda.is_available():
            raise RuntimeError("PillarForwardPath needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device)
        self.generator = VoxelGenerator(self.VOXEL, self.RANGE, 20, 32000)
        self.reader = registry.build_reader(dict(type="PillarFeatureNet", num_filters=[64, 64], num_input_features=5,
                                                 with_distance=False, voxel_size=self.VOXEL, pc_range=self.RANGE))
        self.backbone = registry.build_backbone(dict(type="PointPillarsScatter_S2D", ds_factor=1))
        self.neck = registry.build_neck(dict(type="RPN", layer_nums=[3, 5, 5], ds_layer_strides=[1, 2, 2],
                                             ds_num_filters=[64, 128, 256], us_layer_strides=[1, 2, 4],
                                             us_num_filters=[128, 128, 128], num_input_features=64,
                                             logger=logging.getLogger("RPN")))
        self.head = registry.build_head(dict(type="CenterHead", in_channels=128 * 3,
                                             tasks=[dict(num_class=3, class_names=["VEHICLE", "PEDESTRIAN", "CYCLIST"])],
                                             dataset="waymo", weight=2, code_weights=[1.0] * 8,
                                             common_heads={"reg": (2, 2), "height": (1, 2), "dim": (3, 2), "rot": (2, 2)}))
        self.extractor = SS.BEVFeatureExtractor([-74.88, -74.88], [0.32, 0.32], 1) if second_stage else None
        self.roi_head = registry.build_roi_head(dict(type="RoIHead", input_channels=128 * 3 * 5, code_size=7,
                                                     model_cfg=dict(CLASS_AGNOSTIC=True, SHARED_FC=[256, 256], CLS_FC=[256, 256],
                                                                    REG_FC=[256, 256], DP_RATIO=0.3))) if second_stage else None
        mods = [self.reader, self.backbone, self.neck, self.head] + ([self.roi_head] if second_stage else [])
        for i, m in enumerate(mods):
            m.load_state_dict({k: torch.as_tensor(v) for k, v in synth.random_module_state(m, 31 + i + 100 * seed).items()},
                              strict=False)
            m.to(self.device).eval()
            if hasattr(m, "set_precision"):
                m.set_precision(precision)
        with torch.no_grad():
            self.head.tasks[0].hm[-1].bias.fill_(-1.0)
        self.grid = [int(v) for v in self.generator.grid_size]

    @torch.no_grad()
    def forward_points(self, points, scene_offsets, return_first_stage=False):
        """-> (boxes [B,500,7], scores [B,500], labels i32 [B,500], counts i32 [B]) on the device."""
        from . import _lib
        B = len(scene_offsets) - 1
        vb = self.generator.generate_batch(points, scene_offsets, want_voxels=True)
        f = self.reader(vb.voxels, vb.num_points, vb.coors)
        F_S_a, _, (H, W) = self.backbone.forward_rows(f, vb.coors, B, self.grid)
        ups, (Hu, Wu) = self.neck.forward_rows(F_S_a, B, H, W)
        preds = self.head.forward_rows(ups, B, Hu, Wu)
        rois, roi_scores, roi_labels, _, n_boxes = self.head.select_rows(preds, B, Hu, Wu, self.TEST_CFG)[0]
        if self.roi_head is None:
            return rois, roi_scores, roi_labels, n_boxes
        feats = self.extractor.box_features(ups, B, Hu, Wu, rois, n_boxes, 5)
        rcnn_cls, rcnn_reg = self.roi_head.forward_rows(feats)
        boxes, scores = torch.empty_like(rois), torch.empty_like(roi_scores)
        _lib.check(_lib.load().s2d_roi_refine(rois.data_ptr(), roi_scores.data_ptr(), n_boxes.data_ptr(), B,
                                              rois.shape[1], rcnn_cls.data_ptr(), rcnn_cls.stride(0), rcnn_reg.data_ptr(),
                                              rcnn_reg.stride(0), boxes.data_ptr(), scores.data_ptr(), ops._stream()),
                   "s2d_roi_refine")
        if return_first_stage:
            return boxes, scores, roi_labels, n_boxes, rois, roi_scores, ups.view(B, Hu, Wu, -1)
        return boxes, scores, roi_labels, n_boxes

    forward_host_async = FullForwardPath.forward_host_async

    def states_numpy(self):
        f = lambda m: {k: v.detach().cpu().numpy() for k, v in m.state_dict().items()}
        return dict(reader=f(self.reader), backbone=f(self.backbone), neck=f(self.neck), head=f(self.head),
                    roi=f(self.roi_head) if self.roi_head is not None else None)
