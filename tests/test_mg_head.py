"""SECOND-style anchor head: anchors, box coder and MultiGroupHead against the fixture produced by
the reference's own box_np_ops / box_torch_ops / mg_head modules (tests/golden/make_golden.py mg), the oracle restatement
against the same fixture (CPU), and the CUDA forward + predict against the oracle (GPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import mg_head as OM
from sparse2dense_b200 import anchors as A
from sparse2dense_b200 import registry, synth

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mg_head.npz")
TASKS = [dict(num_class=3, class_names=["VEHICLE", "PEDESTRIAN", "CYCLIST"])]
ASSIGNER = dict(tasks=TASKS, anchor_generators=[
    dict(type="anchor_generator_range", sizes=[2.08, 4.73, 1.77], anchor_ranges=[-74.88, -74.88, 0, 74.88, 74.88, 0],
         rotations=[0, 1.57], matched_threshold=0.55, unmatched_threshold=0.4, class_name="VEHICLE"),
    dict(type="anchor_generator_range", sizes=[0.84, 0.91, 1.74], anchor_ranges=[-74.88, -74.88, 0, 74.88, 74.88, 0],
         rotations=[0, 1.57], matched_threshold=0.5, unmatched_threshold=0.35, class_name="PEDESTRIAN"),
    dict(type="anchor_generator_range", sizes=[0.84, 1.81, 1.77], anchor_ranges=[-74.88, -74.88, 0, 74.88, 74.88, 0],
         rotations=[0, 1.57], matched_threshold=0.5, unmatched_threshold=0.3, class_name="CYCLIST")])
TEST_CFG = dict(post_center_limit_range=[-80, -80, -10.0, 80, 80, 10.0],
                nms=dict(use_rotate_nms=True, use_multi_class_nms=False, nms_pre_max_size=4096, nms_post_max_size=500,
                         nms_iou_threshold=0.7), score_threshold=0.3)


def mg_input(seed, H, W):
    rng = np.random.default_rng(seed)
    return (rng.normal(size=(2, 128, H, W)) * (rng.random((2, 1, H, W)) < 0.4)).astype(np.float32)


def build_head():
    head = registry.build_head(dict(type="MultiGroupHead", mode="3d", in_channels=128, tasks=TASKS, weights=[1],
                                    box_coder=A.build_box_coder(dict(type="ground_box3d_coder", n_dim=7, linear_dim=False,
                                                                     encode_angle_vector=False)),
                                    loss_aux=dict(type="WeightedSoftmaxClassificationLoss"), direction_offset=0.0))
    return head


def check(d, name, got, tol):
    got = np.asarray(got).reshape(-1)[d[name + "_idx"]]
    err = np.abs(got - d[name + "_val"]).max() / float(d[name + "_absmax"])
    assert err <= tol, (name, err)


def test_anchors_coder_and_oracle_match_the_reference_fixture():
    d = np.load(G)
    H, W = int(d["H"]), int(d["W"])
    anchors = A.task_anchors(ASSIGNER, [1, H, W])[0]
    assert anchors.shape == (H * W * 6, 7)
    check(d, "anchors", anchors, 0.0)                                      # bit-exact (same numpy arithmetic)
    head = build_head()
    st = synth.random_module_state(head, int(d["head_seed"]))
    assert set(st) == {f"tasks.0.conv_{n}.{p}" for n in ("box", "cls", "dir") for p in ("weight", "bias")}
    x = torch.from_numpy(mg_input(int(d["input_seed"]), H, W))
    preds = OM.head_forward(st, x)
    for k in ("box_preds", "cls_preds", "dir_cls_preds"):
        check(d, k, preds[k].numpy(), 1e-6)
    dec = OM.decode(preds["box_preds"].reshape(2, -1, 7).numpy(), anchors[None])
    check(d, "decoded", dec, 1e-6)
    mine = head.box_coder.decode_torch(preds["box_preds"].reshape(2, -1, 7), torch.from_numpy(anchors)[None].repeat(2, 1, 1))
    check(d, "decoded", mine.numpy(), 1e-6)
    back = head.box_coder.encode_torch(mine, torch.from_numpy(anchors)[None].repeat(2, 1, 1))
    assert float((back - preds["box_preds"].reshape(2, -1, 7)).abs().max()) < 1e-3      # encode inverts decode


@pytest.mark.gpu
def test_multi_group_head_forward_and_predict_on_gpu():
    from oracle import ref_ops as R
    from sparse2dense_b200 import ops
    R.build()
    d = np.load(G)
    H, W = int(d["H"]), int(d["W"])
    head = build_head()
    st = synth.random_module_state(head, int(d["head_seed"]))
    head.load_state_dict({k: torch.from_numpy(v) for k, v in st.items()})
    head = head.cuda().eval()
    head.set_precision(ops.PRECISION_AUTO)
    x = torch.from_numpy(mg_input(int(d["input_seed"]), H, W))
    before = ops.kernel_launches()
    preds = head(x.cuda())
    assert ops.kernel_launches() > before
    for k in ("box_preds", "cls_preds", "dir_cls_preds"):
        check(d, k, preds[0][k].cpu().numpy(), 2e-5)
    anchors = A.task_anchors(ASSIGNER, [1, H, W])[0]
    ex = dict(anchors=[torch.from_numpy(anchors)[None].repeat(2, 1, 1).cuda()], metadata=[{"token": "a"}, {"token": "b"}])
    out = head.predict(ex, preds, TEST_CFG)
    ref_preds = OM.head_forward(st, x)
    for b in range(2):
        one = {k: v[b:b + 1] for k, v in ref_preds.items()}
        bx, sc, lb = OM.predict(one, anchors, 3, 0.3, 4096, 500, 0.7, TEST_CFG["post_center_limit_range"])
        assert out[b]["metadata"]["token"] == "ab"[b]
        got_b, got_s, got_l = (out[b][k].cpu().numpy() for k in ("box3d_lidar", "scores", "label_preds"))
        assert len(sc) > 50 and abs(len(got_s) - len(sc)) <= max(2, 0.02 * len(sc)), (len(got_s), len(sc))
        # same detections up to candidates within rounding of the score / IoU thresholds: match by centre + label
        hit = 0
        for i in range(len(sc)):
            ok = (np.abs(got_b[:, :2] - bx[i, :2]).max(1) < 1e-3) & (got_l == lb[i]) & (np.abs(got_s - sc[i]) < 1e-4)
            hit += bool(ok.any())
        assert hit >= 0.98 * len(sc), (hit, len(sc))


@pytest.mark.gpu
def test_second_detector_runs_end_to_end_with_anchors():
    """The SECOND student of configs/waymo/voxelnet/waymo_second_3x_distill_interval_5.py (KD_VoxelNet: SpMiddleResNetFHD ->
    S2D_RPN -> MultiGroupHead) built through the registry, one small scene, anchors as AssignTarget would provide them."""
    import logging
    from sparse2dense_b200 import ops
    from sparse2dense_b200.hotpath import concat_clouds
    cfg = dict(type="KD_VoxelNet", pretrained=None, reader=dict(type="VoxelFeatureExtractorV3", num_input_features=5),
               backbone=dict(type="SpMiddleResNetFHD", num_input_features=5, ds_factor=8),
               neck=dict(type="S2D_RPN", layer_nums=[5, 5], ds_layer_strides=[1, 2], ds_num_filters=[128, 256],
                         us_layer_strides=[1, 2], us_num_filters=[256, 256], num_input_features=256,
                         logger=logging.getLogger("RPN")),
               bbox_head=dict(type="MultiGroupHead", mode="3d", in_channels=512, tasks=TASKS, weights=[1],
                              box_coder=A.build_box_coder(dict(type="ground_box3d_coder", n_dim=7, linear_dim=False,
                                                               encode_angle_vector=False)),
                              loss_aux=dict(type="WeightedSoftmaxClassificationLoss"), direction_offset=0.0))
    model = registry.build_detector(cfg, train_cfg=None, test_cfg=TEST_CFG)
    model.backbone.load_state_dict({k: torch.as_tensor(v) for k, v in synth.backbone_state(0).items()}, strict=False)
    for i, m in enumerate((model.neck, model.bbox_head)):
        m.load_state_dict({k: torch.as_tensor(v) for k, v in synth.random_module_state(m, 51 + i).items()}, strict=False)
    model = model.cuda().eval()
    model.set_precision(ops.PRECISION_AUTO)
    clouds = [synth.small_scene(61), synth.small_scene(62)]
    pts, offs = concat_clouds(clouds)
    vb = ops.voxelize(pts.cuda(), offs, synth.WAYMO_VOXEL, synth.WAYMO_RANGE, 5, 150000, want_voxels=True)
    counts = np.diff(vb.offsets_host())
    anchors = A.task_anchors(ASSIGNER, [1, 188, 188])
    example = dict(voxels=vb.voxels, coordinates=vb.coors, num_points=vb.num_points, num_voxels=torch.as_tensor(counts),
                   shape=[np.array([1504, 1504, 40])] * 2, metadata=[{"token": "a"}, {"token": "b"}],
                   anchors=[torch.from_numpy(a)[None].repeat(2, 1, 1).cuda() for a in anchors])
    before = ops.kernel_launches()
    with torch.no_grad():
        out = model(example, return_loss=False)
    assert ops.kernel_launches() - before > 100 and len(out) == 2
    for b in range(2):
        n = out[b]["box3d_lidar"].shape[0]
        assert out[b]["box3d_lidar"].shape == (n, 7) and out[b]["scores"].shape == (n,) and n <= 500
        assert out[b]["metadata"]["token"] == "ab"[b]
        if n:
            assert float(out[b]["scores"].min()) >= 0.3 and int(out[b]["label_preds"].max()) <= 2
