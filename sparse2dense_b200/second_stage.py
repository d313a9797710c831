"""Second stage of the two-stage CenterPoint: ``BEVFeatureExtractor``
(det3d/models/second_stage/bird_eye_view.py:9-40), ``RoIHead`` (det3d/models/roi_heads/roi_head.py:16-106,
roi_head_template.py:18-40,153-183) and ``TwoStageDetector`` (det3d/models/detectors/two_stage.py:8-199) with the
reference's registry names, constructor arguments and state-dict keys; inference path only (``ProposalTargetLayer`` and
the RoI losses are train-time).

Device side: one launch builds the [B*500, 2560] RoI feature matrix from the NHWC BEV rows (box centre + four face
mid-points, bilinear), the 1x1-Conv1d MLP runs on the tcgen05 gather-GEMM with an identity table, one launch decodes
the refined boxes and scores."""
import torch
from torch import nn

from . import _lib, ops, registry
from .dense import ACT_NONE, ACT_RELU, DenseOps, conv_rows, fold_bn, to_nchw, to_rows
from .detectors import BaseDetector
from .registry import DETECTORS, ROI_HEAD, SECOND_STAGE


def _get(cfg, name, default=None):
    if isinstance(cfg, dict):
        return cfg.get(name, default)
    return getattr(cfg, name, default)


@SECOND_STAGE.register_module
class BEVFeatureExtractor(nn.Module):
    def __init__(self, pc_start, voxel_size, out_stride):
        super().__init__()
        self.pc_start = pc_start
        self.voxel_size = voxel_size
        self.out_stride = out_stride

    def box_features(self, bev_rows, B, H, W, boxes, n_boxes, num_point):
        """bev_rows [B*H*W, C] (stride(1) == 1), boxes [B,P,7], n_boxes i32 [B] (device) -> [B*P, num_point*C]."""
        P, C = boxes.shape[1], bev_rows.shape[1]
        out = torch.empty((B * P, num_point * C), dtype=torch.float32, device=bev_rows.device)
        _lib.check(_lib.load().s2d_bev_box_features(
            bev_rows.data_ptr(), bev_rows.stride(0), B, H, W, C, boxes.data_ptr(), n_boxes.data_ptr(), P, num_point,
            _lib.floats(self.pc_start), _lib.floats(self.voxel_size), float(self.out_stride), out.data_ptr(),
            ops._stream()), "s2d_bev_box_features")
        return out

    def forward(self, example, batch_centers, num_point):
        """Reference signature (bird_eye_view.py:24-40): example['bev_feature'] is NHWC [B,H,W,C]; batch_centers is the
        list of per-sample point sets from ``TwoStageDetector.get_box_center``.  Kept for API compatibility; it samples
        arbitrary points (num_point sections concatenated along dim 1)."""
        bev = example["bev_feature"]
        B, H, W, C = bev.shape
        rows = bev.reshape(B * H * W, C)
        ret = []
        for b in range(B):
            pts = batch_centers[b]
            n = pts.shape[0]
            if n == 0:
                ret.append(bev.new_zeros((0, C * num_point)))
                continue
            boxes = torch.zeros((1, n, 7), dtype=torch.float32, device=bev.device)
            boxes[0, :, :2] = pts[:, :2]
            cnt = torch.full((1,), n, dtype=torch.int32, device=bev.device)
            f = self.box_features(rows[b * H * W:(b + 1) * H * W], 1, H, W, boxes, cnt, 1)      # centre point only
            if num_point > 1:
                s = n // num_point
                f = torch.cat([f[i * s:(i + 1) * s] for i in range(num_point)], dim=1)
            ret.append(f)
        return ret


@ROI_HEAD.register_module
class RoIHead(nn.Module):
    def __init__(self, input_channels, model_cfg, num_class=1, code_size=7, test_cfg=None):
        super().__init__()
        self.model_cfg = model_cfg
        self.num_class = num_class
        self.test_cfg = test_cfg
        self.code_size = code_size
        g = _get
        pre = input_channels
        shared, fcs = [], g(model_cfg, "SHARED_FC")
        for k in range(len(fcs)):
            shared.extend([nn.Conv1d(pre, fcs[k], kernel_size=1, bias=False), nn.BatchNorm1d(fcs[k]), nn.ReLU()])
            pre = fcs[k]
            if k != len(fcs) - 1 and g(model_cfg, "DP_RATIO") > 0:
                shared.append(nn.Dropout(g(model_cfg, "DP_RATIO")))
        self.shared_fc_layer = nn.Sequential(*shared)
        self.cls_layers = self.make_fc_layers(pre, self.num_class, g(model_cfg, "CLS_FC"))
        self.reg_layers = self.make_fc_layers(pre, code_size, g(model_cfg, "REG_FC"))
        self.init_weights()
        self._dense = DenseOps()
        self._ident = {}

    def make_fc_layers(self, input_channels, output_channels, fc_list):
        layers, pre = [], input_channels
        for k in range(len(fc_list)):
            layers.extend([nn.Conv1d(pre, fc_list[k], kernel_size=1, bias=False), nn.BatchNorm1d(fc_list[k]), nn.ReLU()])
            pre = fc_list[k]
            if _get(self.model_cfg, "DP_RATIO") >= 0 and k == 0:
                layers.append(nn.Dropout(_get(self.model_cfg, "DP_RATIO")))
        layers.append(nn.Conv1d(pre, output_channels, kernel_size=1, bias=True))
        return nn.Sequential(*layers)

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv1d):
                nn.init.xavier_normal_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
        nn.init.normal_(self.reg_layers[-1].weight, mean=0, std=0.001)

    def set_precision(self, precision):
        self._dense = DenseOps(precision)

    def _mlp(self, name, seq, x):
        """Conv1d(k=1) [+ BatchNorm1d] [+ ReLU] chains as K = 1 gather-GEMMs over the rows of x (Dropout: eval no-op)."""
        D = self._dense
        n = x.shape[0]
        key = (n, str(x.device))
        if key not in self._ident:
            self._ident[key] = ops.alloc_table(1, n, x.device)
            self._ident[key][0].copy_(torch.arange(n, dtype=torch.int32, device=x.device))
        tbl = self._ident[key]
        mods = list(seq)
        j = 0
        while j < len(mods):
            m = mods[j]
            if isinstance(m, nn.Conv1d):
                bn = mods[j + 1] if j + 1 < len(mods) and isinstance(mods[j + 1], nn.BatchNorm1d) else None
                k = j + 1 + int(bn is not None)
                relu = k < len(mods) and isinstance(mods[k], nn.ReLU)

                def build(m=m, bn=bn):
                    kio = m.weight.detach().float()[:, :, 0].t().contiguous().unsqueeze(0)       # [1, Cin, Cout]
                    packed = {}                                   # effective precision -> image (filled by conv_rows)
                    sc, sh = fold_bn(bn, m.bias, m.weight.shape[0], m.weight.device)
                    return kio, packed, sc, sh
                src = [m.weight, m.bias] + ([bn.weight, bn.bias, bn.running_mean, bn.running_var] if bn is not None else [])
                kio, packed, sc, sh = D.cache.get((name, j, D.precision), src, build)
                x = conv_rows(x, kio, tbl, n, sc, sh, ACT_RELU if relu else ACT_NONE, precision=D.precision, packed=packed)
                j = k + int(relu)
            else:
                j += 1
        return x

    def forward_rows(self, roi_features):
        """[B*P, input_channels] -> (rcnn_cls [B*P, num_class], rcnn_reg [B*P, code_size])."""
        if self.training:
            raise NotImplementedError("RoIHead is inference only in this build (call .eval())")
        shared = self._mlp("shared", self.shared_fc_layer, roi_features)
        return self._mlp("cls", self.cls_layers, shared), self._mlp("reg", self.reg_layers, shared)

    def forward(self, batch_dict, training=True):
        """Reference signature (roi_head.py:70-106), inference branch."""
        if training:
            raise NotImplementedError("RoI target assignment / losses are train-time and not built")
        rois, feats = batch_dict["rois"], batch_dict["roi_features"]
        B, P = rois.shape[0], rois.shape[1]
        batch_dict["batch_size"] = B
        cls, reg = self.forward_rows(feats.reshape(B * P, -1).contiguous())
        n = torch.full((B,), P, dtype=torch.int32, device=rois.device)
        boxes = torch.empty((B, P, 7), dtype=torch.float32, device=rois.device)
        scores = torch.empty((B, P), dtype=torch.float32, device=rois.device)
        dummy = torch.ones((B, P), dtype=torch.float32, device=rois.device)
        _lib.check(_lib.load().s2d_roi_refine(rois.contiguous().data_ptr(), dummy.data_ptr(), n.data_ptr(), B, P,
                                              cls.data_ptr(), cls.stride(0), reg.data_ptr(), reg.stride(0),
                                              boxes.data_ptr(), scores.data_ptr(), ops._stream()), "s2d_roi_refine")
        batch_dict["batch_cls_preds"] = cls.view(B, P, -1)
        batch_dict["batch_box_preds"] = boxes
        batch_dict["cls_preds_normalized"] = False
        return batch_dict


@DETECTORS.register_module
class TwoStageDetector(BaseDetector):
    def __init__(self, first_stage_cfg, second_stage_modules, roi_head, NMS_POST_MAXSIZE, num_point=1, freeze=False,
                 **kwargs):
        super(TwoStageDetector, self).__init__()
        self.single_det = registry.build_detector(first_stage_cfg, **kwargs)
        self.NMS_POST_MAXSIZE = NMS_POST_MAXSIZE
        if freeze:
            print("Freeze First Stage Network")
            self.single_det = self.single_det.freeze()
        self.bbox_head = self.single_det.bbox_head
        self.second_stage = nn.ModuleList()
        for module in second_stage_modules:
            self.second_stage.append(registry.build_second_stage_module(module))
        self.roi_head = registry.build_roi_head(roi_head)
        self.num_point = num_point

    @torch.no_grad()
    def forward(self, example, return_loss=True, return_feature=False, **kwargs):
        if return_loss or self.training:
            raise NotImplementedError("the two-stage training branch is not built; call .eval() and return_loss=False")
        assert len(self.second_stage) == 1 and isinstance(self.second_stage[0], BEVFeatureExtractor), \
            "only the BEV second-stage stream of the Waymo configs is built"
        det = self.single_det
        raw, ups, (B, Hu, Wu), voxel_feature, F_S_a, F_S_b, (H, W) = det.first_stage_raw(example)
        assert len(raw) == 1, "the second stage is built for single-task heads (Waymo)"
        rois, roi_scores, roi_labels, _, n_boxes = raw[0]
        P = rois.shape[1]
        assert P == self.NMS_POST_MAXSIZE, "first-stage nms_post_max_size must equal NMS_POST_MAXSIZE"
        feats = self.second_stage[0].box_features(ups, B, Hu, Wu, rois, n_boxes, self.num_point)
        rcnn_cls, rcnn_reg = self.roi_head.forward_rows(feats)
        boxes = torch.empty((B, P, 7), dtype=torch.float32, device=rois.device)
        scores = torch.empty((B, P), dtype=torch.float32, device=rois.device)
        _lib.check(_lib.load().s2d_roi_refine(rois.data_ptr(), roi_scores.data_ptr(), n_boxes.data_ptr(), B, P,
                                              rcnn_cls.data_ptr(), rcnn_cls.stride(0), rcnn_reg.data_ptr(),
                                              rcnn_reg.stride(0), boxes.data_ptr(), scores.data_ptr(), ops._stream()),
                   "s2d_roi_refine")
        counts = ops.read_ints(n_boxes)                                     # the one host sync
        meta = example.get("metadata") or [None] * B
        ret = [dict(box3d_lidar=boxes[i, :counts[i]], scores=scores[i, :counts[i]],
                    label_preds=roi_labels[i, :counts[i]].long(), metadata=meta[i]) for i in range(B)]
        if return_feature:
            return ret, to_nchw(F_S_a, B, H, W), to_nchw(F_S_b, B, H, W)
        return ret
