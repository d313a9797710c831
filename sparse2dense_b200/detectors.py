"""Detectors (det3d/models/detectors/{base,single_stage,voxelnet}.py): the glue that picks the example-dict keys and
calls reader -> backbone -> neck -> head -> predict.  Same registry names, constructor arguments, ``forward`` flags and
return shapes as the reference for the inference path (``return_loss=False``); the training branches (losses, PCR
targets) are not built and raise.

Inside, the dense stage runs on NHWC rows end to end (backbone densifies straight into rows, the head output buffer
feeds the decode kernel), so no NCHW<->NHWC copy of a 512-channel map is made unless a caller asks for the feature maps.
"""
import logging

import torch
from torch import nn

from . import ops, registry
from .dense import to_nchw
from .registry import DETECTORS


class BaseDetector(nn.Module):
    def __init__(self):
        super(BaseDetector, self).__init__()
        self.fp16_enabled = False

    @property
    def with_reader(self):
        return hasattr(self, "reader") and self.reader is not None

    @property
    def with_neck(self):
        return hasattr(self, "neck") and self.neck is not None

    @property
    def with_bbox(self):
        return hasattr(self, "bbox_head") and self.bbox_head is not None

    def init_weights(self, pretrained=None):
        if pretrained is not None:
            logging.getLogger().info("load model from: {}".format(pretrained))

    def set_precision(self, precision):
        """Arithmetic of every conv of the detector: ops.PRECISION_{FP32,TF32,TF32X3,TF32_BF16C,AUTO}."""
        for m in self.modules():
            if m is not self and hasattr(m, "set_precision"):
                m.set_precision(precision)
        return self


@DETECTORS.register_module
class SingleStageDetector(BaseDetector):
    def __init__(self, reader, backbone, neck=None, bbox_head=None, train_cfg=None, test_cfg=None, pretrained=None):
        super(SingleStageDetector, self).__init__()
        self.reader = registry.build_reader(reader)
        self.backbone = registry.build_backbone(backbone)
        if neck is not None:
            self.neck = registry.build_neck(neck)
        self.bbox_head = registry.build_head(bbox_head)
        self.train_cfg = train_cfg
        self.test_cfg = test_cfg
        self.init_weights(pretrained=pretrained)

    def init_weights(self, pretrained=None):
        """single_stage.py:33-40: a missing checkpoint file is reported, not fatal."""
        if pretrained is None:
            return
        from .checkpoint import load_checkpoint
        try:
            load_checkpoint(self, pretrained, map_location="cpu", strict=False)
            print("init weight from {}".format(pretrained))
        except Exception:
            print("no pretrained model at {}".format(pretrained))

    def freeze(self):
        """single_stage.py:58-61 (BatchNorm2d in eval mode with frozen statistics is what the eval path uses anyway)."""
        for p in self.parameters():
            p.requires_grad = False
        return self.eval()


def _example_data(example):
    return dict(features=example["voxels"], num_voxels=example["num_points"], coors=example["coordinates"],
                batch_size=len(example["num_voxels"]), input_shape=example["shape"][0])


@DETECTORS.register_module
class VoxelNet(SingleStageDetector):
    """The teacher / plain CenterPoint-VoxelNet (voxelnet.py:21-142): neck is ``RPN`` returning one map."""

    def extract_feat(self, data):
        input_features = self.reader(data["features"], data["num_voxels"])
        x, voxel_feature = self.backbone(input_features, data["coors"], data["batch_size"], data["input_shape"])
        if self.with_neck:
            x = self.neck(x)
        return x, voxel_feature

    def _rows_forward(self, example):
        data = _example_data(example)
        B = data["batch_size"]
        feats = self.reader(data["features"], data["num_voxels"])
        rows, voxel_feature = self.backbone(feats, data["coors"], B, data["input_shape"], as_rows=True)
        H, W = self.backbone.bev_hw(data["input_shape"])
        return rows, voxel_feature, B, H, W

    def forward(self, example, return_loss=True, return_feature=False, **kwargs):
        if return_loss or self.training:
            raise NotImplementedError("the training branch (losses) is not built; call .eval() and return_loss=False")
        rows, _, B, H, W = self._rows_forward(example)
        backbone_rows = rows
        ups, (Hu, Wu) = self.neck.forward_rows(rows, B, H, W)
        preds = self.bbox_head.forward_rows(ups, B, Hu, Wu)
        dets = self.bbox_head.predict_rows(preds, B, Hu, Wu, self.test_cfg, example.get("metadata"))
        if return_feature:
            return dets, to_nchw(backbone_rows, B, H, W)          # F_D_a: the backbone's dense BEV map (voxelnet.py:66-72)
        return dets


@DETECTORS.register_module
class KD_VoxelNet(VoxelNet):
    """The student (voxelnet.py:144-301): neck is ``S2D_RPN`` returning the 7-tuple."""

    def extract_feat(self, data, train_pcm=True):
        input_features = self.reader(data["features"], data["num_voxels"])
        x, voxel_feature = self.backbone(input_features, data["coors"], data["batch_size"], data["input_shape"])
        if self.with_neck:
            x, gen_offset_2, gen_mask_2, gen_offset_4, gen_mask_4, F_S_a, F_S_b = self.neck(x)
        return x, gen_offset_2, gen_mask_2, gen_offset_4, gen_mask_4, F_S_a, F_S_b, voxel_feature

    def _first_stage_rows(self, example):
        rows, voxel_feature, B, H, W = self._rows_forward(example)
        ups, (Hu, Wu), F_S_a, F_S_b = self.neck.forward_rows(rows, B, H, W)
        preds = self.bbox_head.forward_rows(ups, B, Hu, Wu)
        dets = self.bbox_head.predict_rows(preds, B, Hu, Wu, self.test_cfg, example.get("metadata"))
        return dets, ups, voxel_feature, F_S_a, F_S_b, B, H, W, Hu, Wu

    def forward(self, example, return_loss=True, return_feature=False, **kwargs):
        if return_loss or self.training:
            raise NotImplementedError("the distillation training branch is not built; call .eval() and return_loss=False")
        dets, _, _, F_S_a, F_S_b, B, H, W, _, _ = self._first_stage_rows(example)
        if return_feature:
            return dets, to_nchw(F_S_a, B, H, W), to_nchw(F_S_b, B, H, W)
        return dets

    def forward_two_stage(self, example, return_loss=True, **kwargs):
        """voxelnet.py:266-301 -> (boxes, bev_feature [B,512,H,W], voxel_feature, loss|None, F_S_a, F_S_b)."""
        if return_loss or self.training:
            raise NotImplementedError("the training branch is not built")
        dets, ups, voxel_feature, F_S_a, F_S_b, B, H, W, Hu, Wu = self._first_stage_rows(example)
        return dets, to_nchw(ups, B, Hu, Wu), voxel_feature, None, to_nchw(F_S_a, B, H, W), to_nchw(F_S_b, B, H, W)

    def first_stage_raw(self, example):
        """First stage for ``TwoStageDetector``: the padded device-side detections of every task
        (boxes [B,500,7], scores, labels, cells, counts -- no host sync), the BEV feature as NHWC rows, F_S_a/F_S_b rows."""
        rows, voxel_feature, B, H, W = self._rows_forward(example)
        ups, (Hu, Wu), F_S_a, F_S_b = self.neck.forward_rows(rows, B, H, W)
        preds = self.bbox_head.forward_rows(ups, B, Hu, Wu)
        raw = self.bbox_head.select_rows(preds, B, Hu, Wu, self.test_cfg)
        return raw, ups, (B, Hu, Wu), voxel_feature, F_S_a, F_S_b, (H, W)


@DETECTORS.register_module
class PointPillars(SingleStageDetector):
    """Teacher / plain CenterPoint-Pillar (det3d/models/detectors/point_pillars.py:13-124): reader ``PillarFeatureNet``,
    backbone ``PointPillarsScatter``-style canvas, neck ``RPN``."""

    def extract_feat(self, data):
        input_features = self.reader(data["features"], data["num_voxels"], data["coors"])
        x_fea = self.backbone(input_features, data["coors"], data["batch_size"], data["input_shape"])
        x = self.neck(x_fea) if self.with_neck else x_fea
        return x, x_fea

    def forward(self, example, return_loss=True, **kwargs):
        """point_pillars.py:37-89, ``return_loss=False``: the teacher's role in distillation ->
        (preds, F_D_a, F_D_b): head maps on the dense pillars, the dense canvas and the reconstruction canvas."""
        if return_loss or self.training:
            raise NotImplementedError("the training branch (losses) is not built; call .eval() and return_loss=False")
        pre = "dense_" if "dense_voxels" in example else ""
        B = len(example[pre + "num_voxels"])
        shape = example["shape"][0]
        feats = self.reader(example[pre + "voxels"], example[pre + "num_points"], example[pre + "coordinates"])
        rows, (H, W) = self.backbone.forward_rows(feats, example[pre + "coordinates"], B, shape)
        ups, (Hu, Wu) = self.neck.forward_rows(rows, B, H, W)
        preds = self.bbox_head.forward_rows(ups, B, Hu, Wu)
        preds = [{h: to_nchw(v, B, Hu, Wu) for h, v in d.items()} for d in preds]
        F_D_a = to_nchw(rows, B, H, W)
        F_D_b = None
        if "reconstruction_voxels" in example:
            f2 = self.reader(example["reconstruction_voxels"], example["reconstruction_num_points"],
                             example["reconstruction_coordinates"])
            F_D_b = self.backbone(f2, example["reconstruction_coordinates"], B, shape)
        return preds, F_D_a, F_D_b


@DETECTORS.register_module
class KD_PointPillars(PointPillars):
    """The pillar student (point_pillars.py:127-251): backbone ``PointPillarsScatter_S2D`` returns (F_S_a, F_S_b, ..),
    the RPN runs on F_S_a.  Inference path only."""

    def extract_feat(self, data):
        input_features = self.reader(data["features"], data["num_voxels"], data["coors"])
        F_S_a, F_S_b, gen_offset, gen_mask = self.backbone(input_features, data["coors"], data["batch_size"],
                                                           data["input_shape"])
        x = self.neck(F_S_a) if self.with_neck else F_S_a
        return x, F_S_a, F_S_b, gen_offset, gen_mask

    def _rows(self, example):
        data = _example_data(example)
        B = data["batch_size"]
        feats = self.reader(data["features"], data["num_voxels"], data["coors"])
        F_S_a, F_S_b, (H, W) = self.backbone.forward_rows(feats, data["coors"], B, data["input_shape"])
        ups, (Hu, Wu) = self.neck.forward_rows(F_S_a, B, H, W)
        preds = self.bbox_head.forward_rows(ups, B, Hu, Wu)
        return preds, ups, F_S_a, F_S_b, B, H, W, Hu, Wu

    def forward(self, example, return_loss=True, **kwargs):
        if return_loss or self.training:
            raise NotImplementedError("the distillation training branch is not built; call .eval() and return_loss=False")
        preds, _, _, _, B, _, _, Hu, Wu = self._rows(example)
        return self.bbox_head.predict_rows(preds, B, Hu, Wu, self.test_cfg, example.get("metadata"))

    def forward_two_stage(self, example, return_loss=True, **kwargs):
        """point_pillars.py:215-251 -> (boxes, bev_feature, None, None, F_S_a, F_S_b)."""
        if return_loss or self.training:
            raise NotImplementedError("the training branch is not built")
        preds, ups, F_S_a, F_S_b, B, H, W, Hu, Wu = self._rows(example)
        boxes = self.bbox_head.predict_rows(preds, B, Hu, Wu, self.test_cfg, example.get("metadata"))
        return boxes, to_nchw(ups, B, Hu, Wu), None, None, to_nchw(F_S_a, B, H, W), to_nchw(F_S_b, B, H, W)

    def first_stage_raw(self, example):
        preds, ups, F_S_a, F_S_b, B, H, W, Hu, Wu = self._rows(example)
        raw = self.bbox_head.select_rows(preds, B, Hu, Wu, self.test_cfg)
        return raw, ups, (B, Hu, Wu), None, F_S_a, F_S_b, (H, W)
