"""GPU parity of the PointPillars + S2D student (fused PFN, pillar scatter, S2D encoder/decoder with max-pool / nearest
upsample, three-stage RPN incl. the stride-4 transposed conv) against the reference-generated fixture and the oracle."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F
from torch import nn

from oracle import pillars as OP
from sparse2dense_b200 import dense, ops
from pillar_common import G, PP_RANGE, PP_RPN, PP_VOXEL, check, module_states, pillar_inputs

pytestmark = pytest.mark.gpu


def test_dense_helpers_vs_torch():
    torch.manual_seed(1)
    B, C, H, W = 2, 64, 14, 10
    x = torch.randn(B, C, H, W)
    D = dense.DenseOps(ops.PRECISION_AUTO)
    rows = dense.to_rows(x.cuda())
    y, Ho, Wo = D.maxpool2(rows, B, H, W)
    assert torch.equal(dense.to_nchw(y, B, Ho, Wo).cpu(), F.max_pool2d(x, 2, 2))
    for (ho, wo, sc) in [(27, 19, None), (2 * H, 2 * W, 2)]:
        ref = F.interpolate(x, size=(ho, wo)) if sc is None else F.interpolate(x, scale_factor=2)
        got = dense.to_nchw(D.upsample_nearest(rows, B, H, W, ho, wo, scale=sc), B, ho, wo).cpu()
        assert torch.equal(got, ref)
    for k, s, p in [(4, 4, 0), (2, 2, 0), (4, 2, 1)]:
        conv = nn.ConvTranspose2d(C, 96, k, s, p, bias=False).eval()
        bn = nn.BatchNorm2d(96).eval()
        bn.running_mean.normal_(0, 0.1); bn.running_var.uniform_(0.5, 1.5)
        with torch.no_grad():
            ref = F.relu(bn(conv(x)))
        y, Ho, Wo = D.tconv(f"t{k}{s}", rows, B, H, W, conv.cuda(), bn.cuda(), dense.ACT_RELU)
        got = dense.to_nchw(y, B, Ho, Wo).cpu()
        assert (Ho, Wo) == (s * H, s * W)
        assert float((got - ref).abs().max() / ref.abs().max()) < 2e-5, (k, s, p)


@pytest.mark.parametrize("precision,tol", [(ops.PRECISION_AUTO, 1e-3), (ops.PRECISION_FP32, 1e-4)])
def test_pillar_student_vs_reference_fixture_and_oracle(precision, tol):
    d = np.load(G)
    v, c, n = pillar_inputs(int(d["scene_seed"]))
    (reader, rs), (bb, bs), (neck, ns) = module_states(d)
    for m, st in ((reader, rs), (bb, bs), (neck, ns)):
        m.load_state_dict({k: torch.from_numpy(val) for k, val in st.items()}, strict=False)
        m.cuda().eval()
        if hasattr(m, "set_precision"):
            m.set_precision(precision)
    vc, cc, nc = torch.from_numpy(v).cuda(), torch.from_numpy(c).cuda(), torch.from_numpy(n).cuda()
    before = ops.kernel_launches()
    with torch.no_grad():
        f = reader(vc, nc, cc)
        fa, fb, (H, W) = bb.forward_rows(f, cc, 1, [468, 468, 1])
        ups, (Hu, Wu) = neck.forward_rows(fa, 1, H, W)
        fa_n, fb_n, _, _ = bb(f, cc, 1, [468, 468, 1])                       # reference-style NCHW return
    assert ops.kernel_launches() - before > 50 and (Hu, Wu) == (468, 468) and ups.shape[1] == 384
    check(d, "pfn", f.cpu().numpy(), 1e-5)
    check(d, "F_S_a", fa_n.cpu().numpy(), tol)
    check(d, "F_S_b", fb_n.cpu().numpy(), tol)
    err = check(d, "x", dense.to_nchw(ups, 1, Hu, Wu).cpu().numpy(), tol)
    print("pillar student, precision", precision, "rel err of the RPN output vs the reference fixture", err)
    # and against the oracle on the SAME device-side PFN output (isolates the dense stage)
    with torch.no_grad():
        ofa, ofb = OP.scatter_s2d_forward(bs, f.cpu().numpy(), c, 1, 468, 468)
    assert float((fa_n.cpu() - ofa).abs().max() / ofa.abs().max()) < tol
    assert float((fb_n.cpu() - ofb).abs().max() / ofb.abs().max()) < tol


def test_pp_two_stage_detector_runs_from_reference_style_config():
    """KD_PointPillars + BEV second stage built through the registry from the reference's PP config values."""
    import logging
    from sparse2dense_b200 import registry, synth
    tasks = [dict(num_class=3, class_names=["VEHICLE", "PEDESTRIAN", "CYCLIST"])]
    S_model = dict(
        type="TwoStageDetector",
        first_stage_cfg=dict(
            type="KD_PointPillars", pretrained=None,
            reader=dict(type="PillarFeatureNet", num_filters=[64, 64], num_input_features=5, with_distance=False,
                        voxel_size=PP_VOXEL, pc_range=PP_RANGE),
            backbone=dict(type="PointPillarsScatter_S2D", ds_factor=1),
            neck=dict(type="RPN", logger=logging.getLogger("RPN"), **PP_RPN),
            bbox_head=dict(type="CenterHead", in_channels=128 * 3, tasks=tasks, dataset="waymo", weight=2,
                           code_weights=[1.0] * 8, common_heads={"reg": (2, 2), "height": (1, 2), "dim": (3, 2), "rot": (2, 2)})),
        second_stage_modules=[dict(type="BEVFeatureExtractor", pc_start=[-74.88, -74.88], voxel_size=[0.32, 0.32], out_stride=1)],
        roi_head=dict(type="RoIHead", input_channels=128 * 3 * 5, code_size=7,
                      model_cfg=dict(CLASS_AGNOSTIC=True, SHARED_FC=[256, 256], CLS_FC=[256, 256], REG_FC=[256, 256], DP_RATIO=0.3)),
        NMS_POST_MAXSIZE=500, num_point=5, freeze=True)
    test_cfg = dict(post_center_limit_range=[-80, -80, -10.0, 80, 80, 10.0],
                    nms=dict(use_rotate_nms=True, nms_pre_max_size=4096, nms_post_max_size=500, nms_iou_threshold=0.7),
                    score_threshold=0.1, pc_range=[-74.88, -74.88], out_size_factor=1, voxel_size=[0.32, 0.32])
    torch.manual_seed(0)
    model = registry.build_detector(S_model, train_cfg=None, test_cfg=test_cfg)
    det = model.single_det
    for m, seed in ((det.reader, 31), (det.backbone, 32), (det.neck, 33), (det.bbox_head, 34)):
        m.load_state_dict({k: torch.as_tensor(v) for k, v in synth.random_module_state(m, seed).items()}, strict=False)
    model = model.cuda().eval()
    model.set_precision(ops.PRECISION_AUTO)
    vs, cs, ns, counts = [], [], [], []
    for b, seed in enumerate((2000, 2001)):
        v, c, n = pillar_inputs(seed)
        c[:, 0] = b
        vs.append(v); cs.append(c); ns.append(n); counts.append(len(v))
    example = dict(voxels=torch.from_numpy(np.concatenate(vs)).cuda(), coordinates=torch.from_numpy(np.concatenate(cs)).cuda(),
                   num_points=torch.from_numpy(np.concatenate(ns)).cuda(), num_voxels=torch.tensor(counts),
                   shape=[np.array([468, 468, 1])] * 2, metadata=[None, None])
    out, F_S_a, F_S_b = model(example, return_loss=False, return_feature=True)
    assert len(out) == 2 and F_S_a.shape == (2, 64, 468, 468) and F_S_b.shape == (2, 64, 468, 468)
    for d in out:
        k = len(d["scores"])
        assert d["box3d_lidar"].shape == (k, 7) and d["label_preds"].shape == (k,) and k <= 500
        assert torch.isfinite(d["scores"]).all()


def test_pillar_teacher_forward_shapes_and_exact_scatter():
    import logging
    from sparse2dense_b200 import registry, synth
    tasks = [dict(num_class=3, class_names=["VEHICLE", "PEDESTRIAN", "CYCLIST"])]
    cfg = dict(type="PointPillars", pretrained=None,
               reader=dict(type="PillarFeatureNet", num_filters=[64, 64], num_input_features=5, with_distance=False,
                           voxel_size=PP_VOXEL, pc_range=PP_RANGE),
               backbone=dict(type="PointPillarsScatter", ds_factor=1),
               neck=dict(type="RPN", logger=logging.getLogger("RPN"), **PP_RPN),
               bbox_head=dict(type="CenterHead", in_channels=128 * 3, tasks=tasks, dataset="waymo", weight=2, code_weights=[1.0] * 8,
                              common_heads={"reg": (2, 2), "height": (1, 2), "dim": (3, 2), "rot": (2, 2)}))
    torch.manual_seed(0)
    model = registry.build_detector(cfg, train_cfg=None, test_cfg=None)
    for m, seed in ((model.reader, 31), (model.neck, 33), (model.bbox_head, 34)):
        m.load_state_dict({k: torch.as_tensor(v) for k, v in synth.random_module_state(m, seed).items()}, strict=False)
    model = model.cuda().eval()
    v, c, n = pillar_inputs(2000)
    vc, cc, nc = torch.from_numpy(v).cuda(), torch.from_numpy(c).cuda(), torch.from_numpy(n).cuda()
    example = dict(voxels=vc, coordinates=cc, num_points=nc, num_voxels=torch.tensor([len(v)]), shape=[np.array([468, 468, 1])],
                   reconstruction_voxels=vc, reconstruction_coordinates=cc, reconstruction_num_points=nc,
                   reconstruction_num_voxels=torch.tensor([len(v)]))
    preds, F_D_a, F_D_b = model(example, return_loss=False)
    assert preds[0]["hm"].shape == (1, 3, 468, 468) and preds[0]["dim"].shape == (1, 3, 468, 468)
    assert F_D_a.shape == (1, 64, 468, 468) and torch.equal(F_D_a, F_D_b)
    f = model.reader(vc, nc, cc).cpu().numpy()
    canvas = np.zeros((64, 468 * 468), np.float32)
    canvas[:, c[:, 2] * 468 + c[:, 3]] = f.T                                  # pillar_encoder.py:357-363
    assert np.array_equal(F_D_a.cpu().numpy().reshape(64, -1), canvas)
