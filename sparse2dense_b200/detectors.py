"""Detectors (det3d/models/detectors/{base,single_stage,voxelnet}.py): the glue that picks the example-dict keys and
calls reader -> backbone -> neck -> head -> predict.  Same registry names, constructor arguments, ``forward`` flags and
return shapes as the reference.  ``VoxelNet`` / ``KD_VoxelNet`` also carry the training branches (``return_loss=True``:
CenterHead losses, the student's feature maps and PCR losses) on the autograd operators of
``autograd.py``; the pillar detectors are inference only.

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


def _predict(head, preds, B, Hu, Wu, test_cfg, example):
    """Head-specific decoding of row-form predictions: CenterHead decodes on the device from the rows; the anchor head
    (MultiGroupHead) needs ``example["anchors"]`` and takes its reference-style dicts (rows are already NHWC-flattened)."""
    if hasattr(head, "box_coder"):
        return head.predict(example, preds, test_cfg)
    return head.predict_rows(preds, B, Hu, Wu, test_cfg, example.get("metadata"))


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


def _example_data(example, prefix=""):
    return dict(features=example[prefix + "voxels"], num_voxels=example[prefix + "num_points"],
                coors=example[prefix + "coordinates"], batch_size=len(example[prefix + "num_voxels"]),
                input_shape=example["shape"][0])


def _as_loss_maps(preds_rows, B, H, W):
    from .losses import Rows
    return [{h: Rows(v, B, H * W) for h, v in d.items()} for d in preds_rows]


def _nchw(rows, B, H, W, differentiable):
    if rows is None:
        return None
    if differentiable:
        return rows.reshape(B, H, W, rows.shape[1]).permute(0, 3, 1, 2)
    return to_nchw(rows, B, H, W)


@DETECTORS.register_module
class VoxelNet(SingleStageDetector):
    """The teacher / plain CenterPoint-VoxelNet (voxelnet.py:21-142): neck is ``RPN`` returning one map."""

    def extract_feat(self, data):
        input_features = self.reader(data["features"], data["num_voxels"])
        x, voxel_feature = self.backbone(input_features, data["coors"], data["batch_size"], data["input_shape"])
        if self.with_neck:
            x = self.neck(x)
        return x, voxel_feature

    def _rows_forward(self, example, prefix=""):
        data = _example_data(example, prefix)
        B = data["batch_size"]
        feats = self.reader(data["features"], data["num_voxels"])
        rows, voxel_feature = self.backbone(feats, data["coors"], B, data["input_shape"], as_rows=True)
        H, W = self.backbone.bev_hw(data["input_shape"])
        return rows, voxel_feature, B, H, W

    def teacher_rows(self, example, recon=True):
        """The teacher's part of a distillation step (voxelnet.py:47-92 with return_feature / return_recon_feature), all on
        NHWC rows: head logits on the DENSE (multi-sweep) voxels, F_D_a = backbone BEV map of the dense voxels,
        F_D_b = backbone BEV map of the reconstruction voxels."""
        prefix = "dense_" if "dense_voxels" in example else ""
        rows, _, B, H, W = self._rows_forward(example, prefix)
        F_D_b = None
        if recon and "reconstruction_voxels" in example:
            F_D_b = self._rows_forward(example, "reconstruction_")[0]
        ups, (Hu, Wu) = self.neck.forward_rows(rows, B, H, W)
        preds = self.bbox_head.forward_rows(ups, B, Hu, Wu)
        return preds, rows, F_D_b, (B, H, W, Hu, Wu)

    def forward(self, example, return_loss=True, return_feature=False, return_recon_feature=False, **kwargs):
        """voxelnet.py:47-104."""
        grad = torch.is_grad_enabled() and self.training
        preds, F_D_a, F_D_b, (B, H, W, Hu, Wu) = self.teacher_rows(example, recon=return_recon_feature)
        if return_loss:
            loss = self.bbox_head.loss(example, _as_loss_maps(preds, B, Hu, Wu))
            if not return_feature:
                return loss
            return loss, _nchw(F_D_a, B, H, W, grad), _nchw(F_D_b, B, H, W, grad)
        if return_feature and return_recon_feature:
            preds_nchw = [{h: _nchw(v, B, Hu, Wu, grad) for h, v in d.items()} for d in preds]
            return preds_nchw, _nchw(F_D_a, B, H, W, grad), _nchw(F_D_b, B, H, W, grad)
        dets = _predict(self.bbox_head, preds, B, Hu, Wu, self.test_cfg, example)
        if return_feature:
            # F_D_a: the backbone's dense BEV map (voxelnet.py:66-72); F_D_b only exists with return_recon_feature
            return dets, _nchw(F_D_a, B, H, W, grad)
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
        dets = _predict(self.bbox_head, preds, B, Hu, Wu, self.test_cfg, example)
        return dets, ups, voxel_feature, F_S_a, F_S_b, B, H, W, Hu, Wu

    def student_rows(self, example, with_loss=True):
        """The student's part of a distillation step (voxelnet.py:188-260) on NHWC rows ->
        dict(loss, preds, F_S_a, F_S_b, mask_loss, comp_loss, dims).  The PCR losses (mask_offset_loss, :171-185) come from
        ``pcr.pcr_losses`` over the neck's generator outputs when the neck ran its PCR branch."""
        rows, voxel_feature, B, H, W = self._rows_forward(example)
        ups, (Hu, Wu), F_S_a, F_S_b = self.neck.forward_rows(rows, B, H, W)
        mask_loss = comp_loss = 0
        if self.training and with_loss and getattr(self.neck, "pcr_rows", None) is not None:
            from . import pcr
            mask_loss, comp_loss = pcr.pcr_losses(self, example, self.neck.pcr_rows, B, H, W)
        preds = self.bbox_head.forward_rows(ups, B, Hu, Wu)
        loss = self.bbox_head.loss(example, _as_loss_maps(preds, B, Hu, Wu)) if with_loss else None
        return dict(loss=loss, preds=preds, F_S_a=F_S_a, F_S_b=F_S_b, mask_loss=mask_loss, comp_loss=comp_loss,
                    dims=(B, H, W, Hu, Wu), ups=ups, voxel_feature=voxel_feature)

    def forward(self, example, return_loss=True, return_feature=False, **kwargs):
        """voxelnet.py:188-265."""
        if return_loss:
            r = self.student_rows(example)
            B, H, W, Hu, Wu = r["dims"]
            grad = torch.is_grad_enabled()
            preds = [{h: _nchw(v, B, Hu, Wu, grad) for h, v in d.items()} for d in r["preds"]]
            if not return_feature:
                return r["loss"], preds
            return (r["loss"], _nchw(r["F_S_a"], B, H, W, grad), _nchw(r["F_S_b"], B, H, W, grad), preds, r["mask_loss"],
                    r["comp_loss"])
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
        return _predict(self.bbox_head, preds, B, Hu, Wu, self.test_cfg, example)

    def forward_two_stage(self, example, return_loss=True, **kwargs):
        """point_pillars.py:215-251 -> (boxes, bev_feature, None, None, F_S_a, F_S_b)."""
        if return_loss or self.training:
            raise NotImplementedError("the training branch is not built")
        preds, ups, F_S_a, F_S_b, B, H, W, Hu, Wu = self._rows(example)
        boxes = _predict(self.bbox_head, preds, B, Hu, Wu, self.test_cfg, example)
        return boxes, to_nchw(ups, B, Hu, Wu), None, None, to_nchw(F_S_a, B, H, W), to_nchw(F_S_b, B, H, W)

    def first_stage_raw(self, example):
        preds, ups, F_S_a, F_S_b, B, H, W, Hu, Wu = self._rows(example)
        raw = self.bbox_head.select_rows(preds, B, Hu, Wu, self.test_cfg)
        return raw, ups, (B, Hu, Wu), None, F_S_a, F_S_b, (H, W)
