"""GPU parity of the second stage (RoI feature gather, RoI MLP on the gather-GEMM, box refinement) against the
oracle and the reference-generated fixture; plus the two-stage detector built from the unchanged reference config."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_ops as R
from sparse2dense_b200 import _lib, ops, registry, second_stage, synth

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
G = os.path.join(HERE, "golden", "two_stage.npz")
PC, VS, STRIDE = [-75.2, -75.2], [0.1, 0.1], 8
ROI_CFG = dict(CLASS_AGNOSTIC=True, SHARED_FC=[256, 256], CLS_FC=[256, 256], REG_FC=[256, 256], DP_RATIO=0.3)


def pad(arrs, P, shape_tail, dtype):
    out = np.zeros((len(arrs), P) + shape_tail, dtype)
    for i, a in enumerate(arrs):
        out[i, : len(a)] = a
    return out


@pytest.mark.parametrize("precision,tol", [(ops.PRECISION_FP32, 1e-5), (ops.PRECISION_AUTO, 1e-4)])
def test_second_stage_vs_reference_fixture(precision, tol):
    d = np.load(G)
    B, C, H, W = d["bev"].shape
    P = 500
    bev_rows = torch.from_numpy(np.ascontiguousarray(d["bev"].transpose(0, 2, 3, 1)).reshape(B * H * W, C)).cuda()
    boxes = [d[f"in_boxes_{b}"] for b in range(B)]
    n = torch.tensor([len(b) for b in boxes], dtype=torch.int32, device="cuda")
    rois = torch.from_numpy(pad(boxes, P, (7,), np.float32)).cuda()
    scores = torch.from_numpy(pad([d[f"in_scores_{b}"] for b in range(B)], P, (), np.float32)).cuda()
    ext = second_stage.BEVFeatureExtractor(PC, VS, STRIDE)
    feats = ext.box_features(bev_rows, B, H, W, rois, n, 5)
    for b in range(B):
        ref = d[f"roi_features_{b}"]
        got = feats[b * P: b * P + len(ref)].cpu().numpy()
        assert np.abs(got - ref).max() <= 2e-5 * np.abs(ref).max()
        assert (feats[b * P + len(ref): (b + 1) * P] == 0).all()                       # padded RoI slots
    head = second_stage.RoIHead(input_channels=C * 5, model_cfg=ROI_CFG, code_size=7)
    head.load_state_dict({k[4:]: torch.from_numpy(d[k]) for k in d.files if k.startswith("roi.")})
    head = head.cuda().eval()
    head.set_precision(precision)
    cls, reg = head.forward_rows(feats)
    out_b = torch.empty((B, P, 7), device="cuda"); out_s = torch.empty((B, P), device="cuda")
    _lib.check(_lib.load().s2d_roi_refine(rois.data_ptr(), scores.data_ptr(), n.data_ptr(), B, P, cls.data_ptr(),
                                          cls.stride(0), reg.data_ptr(), reg.stride(0), out_b.data_ptr(), out_s.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream))
    for b in range(B):
        k = len(boxes[b])
        scale = max(1.0, np.abs(d[f"boxes_{b}"]).max())
        assert np.abs(out_b[b, :k].cpu().numpy() - d[f"boxes_{b}"]).max() <= 50 * tol * scale
        np.testing.assert_allclose(out_s[b, :k].cpu().numpy(), d[f"scores_{b}"], rtol=0, atol=20 * tol)
        assert (out_b[b, k:] == 0).all()


def test_bev_feature_extractor_reference_signature():
    d = np.load(G)
    bev = torch.from_numpy(np.ascontiguousarray(d["bev"].transpose(0, 2, 3, 1))).cuda()
    ext = second_stage.BEVFeatureExtractor(PC, VS, STRIDE)
    centers = []
    for b in range(2):
        pts = R.box_sample_points(d[f"in_boxes_{b}"], 5)                       # [5, n, 2]
        centers.append(torch.from_numpy(np.concatenate([pts.reshape(-1, 2), np.zeros((pts.shape[0] * pts.shape[1], 1), np.float32)], 1)).cuda())
    out = ext.forward({"bev_feature": bev}, centers, 5)
    for b in range(2):
        ref = d[f"roi_features_{b}"]
        assert np.abs(out[b].cpu().numpy() - ref).max() <= 2e-5 * np.abs(ref).max()


def test_two_stage_detector_from_reference_config_runs_and_matches_oracle_second_stage():
    """The unchanged reference config -> TwoStageDetector; first-stage outputs feed the oracle second stage."""
    cfg_src = '''
import itertools, logging
from det3d.utils.config_tool import get_downsample_factor
tasks = [dict(num_class=3, class_names=['VEHICLE', 'PEDESTRIAN', 'CYCLIST'])]
S_model = dict(type='TwoStageDetector',
    first_stage_cfg=dict(type="KD_VoxelNet", pretrained=None,
        reader=dict(type="VoxelFeatureExtractorV3", num_input_features=5),
        backbone=dict(type="SpMiddleResNetFHD", num_input_features=5, ds_factor=8),
        neck=dict(type="S2D_RPN", layer_nums=[5, 5], ds_layer_strides=[1, 2], ds_num_filters=[128, 256],
                  us_layer_strides=[1, 2], us_num_filters=[256, 256], num_input_features=256, logger=logging.getLogger("RPN")),
        bbox_head=dict(type="CenterHead", in_channels=sum([256, 256]), tasks=tasks, dataset='waymo', weight=2,
                       code_weights=[1.0] * 8, common_heads={'reg': (2, 2), 'height': (1, 2), 'dim': (3, 2), 'rot': (2, 2)})),
    second_stage_modules=[dict(type="BEVFeatureExtractor", pc_start=[-75.2, -75.2], voxel_size=[0.1, 0.1], out_stride=8)],
    roi_head=dict(type="RoIHead", input_channels=512 * 5,
                  model_cfg=dict(CLASS_AGNOSTIC=True, SHARED_FC=[256, 256], CLS_FC=[256, 256], REG_FC=[256, 256], DP_RATIO=0.3),
                  code_size=7),
    NMS_POST_MAXSIZE=500, num_point=5, freeze=True)
test_cfg = dict(post_center_limit_range=[-80, -80, -10.0, 80, 80, 10.0], max_per_img=4096,
    nms=dict(use_rotate_nms=True, use_multi_class_nms=False, nms_pre_max_size=4096, nms_post_max_size=500, nms_iou_threshold=0.7),
    score_threshold=0.1, pc_range=[-75.2, -75.2], out_size_factor=get_downsample_factor(S_model), voxel_size=[0.1, 0.1])
'''
    import tempfile
    from sparse2dense_b200 import Config
    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "two_stage_cfg.py")
        with open(path, "w") as f:
            f.write(cfg_src)
        cfg = Config.fromfile(path)
    assert cfg.test_cfg.out_size_factor == 8
    torch.manual_seed(0)
    model = registry.build_detector(cfg.S_model, train_cfg=None, test_cfg=cfg.test_cfg)
    det = model.single_det
    det.backbone.load_state_dict({k: torch.as_tensor(v) for k, v in synth.backbone_state(0).items()}, strict=False)
    det.neck.load_state_dict({k: torch.as_tensor(v) for k, v in synth.random_module_state(det.neck, 11).items()}, strict=False)
    det.bbox_head.load_state_dict({k: torch.as_tensor(v) for k, v in synth.random_module_state(det.bbox_head, 12).items()}, strict=False)
    with torch.no_grad():
        det.bbox_head.tasks[0].hm[-1].bias.fill_(-1.0)                 # enough confident cells for a real NMS load
    model = model.cuda().eval()
    model.set_precision(ops.PRECISION_AUTO)
    clouds = [synth.lidar_scene(1000), synth.lidar_scene(1001)]
    gen = det_voxels(clouds)
    example = dict(voxels=gen.voxels, coordinates=gen.coors, num_points=gen.num_points,
                   num_voxels=torch.tensor([o2 - o1 for o1, o2 in zip(gen.offsets_host()[:-1], gen.offsets_host()[1:])]),
                   shape=[np.array([1504, 1504, 40])] * 2, metadata=[{"token": "a"}, {"token": "b"}])
    before = ops.kernel_launches()
    out, F_S_a, F_S_b = model(example, return_loss=False, return_feature=True)
    assert ops.kernel_launches() > before and F_S_a.shape == (2, 256, 188, 188) and F_S_b.shape == (2, 256, 188, 188)
    raw, ups, (B, Hu, Wu), _, _, _, _ = det.first_stage_raw(example)
    rois, roi_scores, roi_labels, _, n_boxes = raw[0]
    bev = ups.view(B, Hu, Wu, -1).cpu().numpy()
    state = {k: v.cpu().numpy() for k, v in model.roi_head.state_dict().items()}
    assert sum(n_boxes.cpu().tolist()) > 20, "the synthetic head should produce some detections"
    for b in range(B):
        k = int(n_boxes[b])
        assert len(out[b]["scores"]) == k and out[b]["metadata"]["token"] == "ab"[b]
        f = R.roi_features(bev[b], rois[b, :k].cpu().numpy(), PC, VS, STRIDE)
        cls, reg = R.roi_head_forward(state, f)
        ob, sc = R.roi_refine(rois[b, :k].cpu().numpy(), roi_scores[b, :k].cpu().numpy(), cls, reg)
        got = out[b]["box3d_lidar"].cpu().numpy()
        fin = np.isfinite(ob)                                   # random weights: exp(dim) overflows in a few boxes
        assert np.array_equal(fin, np.isfinite(got)) and fin.mean() > 0.9
        assert np.abs(got[fin] - ob[fin]).max() < 1e-3 * max(1.0, np.abs(ob[fin]).max())
        np.testing.assert_allclose(out[b]["scores"].cpu().numpy(), sc, rtol=0, atol=1e-3)
        assert np.array_equal(out[b]["label_preds"].cpu().numpy(), roi_labels[b, :k].cpu().numpy())


def det_voxels(clouds):
    from sparse2dense_b200.hotpath import concat_clouds
    pts, offs = concat_clouds(clouds)
    return ops.voxelize(pts.cuda(), offs, synth.WAYMO_VOXEL, synth.WAYMO_RANGE, 5, 150000, want_voxels=True)
