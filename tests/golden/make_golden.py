"""Generates the committed golden vectors.  Run in the AUTHORING container only
(it imports the reference from /root/reference, which does not exist on the GPU box):

    python tests/golden/make_golden.py

voxelize_*.npz : outputs of the reference's own numba voxelizer
                 (det3d/ops/point_cloud/point_cloud_ops.py:112-184) imported by file path.
spconv_*.npz   : outputs of a dense torch ``F.conv3d`` formulation of spconv's SubMConv3d /
                 SparseConv3d semantics.  spconv itself is not installable
                 here (un-vendored dependency @ 7342772), so these pin the *restatement*, not spconv.
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from sparse2dense_b200 import synth  # noqa: E402


def ref_voxelizer():
    spec = importlib.util.spec_from_file_location("pco", REF + "/det3d/ops/point_cloud/point_cloud_ops.py")
    pco = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(pco)
    return pco.points_to_voxel


def boundary_cloud(rng, vs, rg, n=4000):
    """Points sitting on / next to voxel faces and range edges (fp32 rounding cases)."""
    vs, rg = np.asarray(vs, np.float32), np.asarray(rg, np.float32)
    g = np.round((rg[3:] - rg[:3]) / vs).astype(np.int64)
    idx = np.stack([rng.integers(-1, g[j] + 1, n) for j in range(3)], 1)
    base = (rg[:3] + idx.astype(np.float32) * vs).astype(np.float32)
    ulps = rng.integers(-2, 3, size=(n, 3))
    pts = base.copy()
    for _ in range(2):
        up = np.nextafter(pts, np.float32(np.inf), dtype=np.float32)
        dn = np.nextafter(pts, np.float32(-np.inf), dtype=np.float32)
        pts = np.where(ulps > 0, up, np.where(ulps < 0, dn, pts))
        ulps = ulps - np.sign(ulps)
    feats = rng.uniform(0, 1, (n, 2)).astype(np.float32)
    # repeat some points so voxels overflow max_points
    out = np.concatenate([pts, feats], 1).astype(np.float32)
    out = np.concatenate([out, out[: n // 4], out[: n // 8]], 0)
    return np.ascontiguousarray(out)


def voxel_cases():
    rng = np.random.default_rng(7)
    wv, wr = synth.WAYMO_VOXEL, synth.WAYMO_RANGE
    small = synth.lidar_scene(11, n_beams=10, n_azimuth=900, n_cylinders=20, second_return=0.1)
    dense_small = small.copy()
    dense_small[:, :2] *= 0.15                      # crowd the points so voxels exceed 5 points
    shuf = small.copy()
    rng.shuffle(shuf, axis=0)
    pv, pr = (0.32, 0.32, 6.0), (-74.88, -74.88, -2.0, 74.88, 74.88, 4.0)
    return {
        "waymo_small": (small, wv, wr, 5, 150000),
        "waymo_crowded": (dense_small, wv, wr, 5, 150000),
        "waymo_cap": (shuf, wv, wr, 5, 1500),         # hits max_voxels: later voxels dropped, old ones still fill
        "waymo_boundary": (boundary_cloud(rng, wv, wr), wv, wr, 5, 150000),
        "pillar_small": (small, pv, pr, 20, 32000),
        "empty": (np.zeros((0, 5), np.float32), wv, wr, 5, 150000),
        "all_outside": (small[:1500] + np.float32(500.0), wv, wr, 5, 150000),
    }


def make_voxel_goldens():
    p2v = ref_voxelizer()
    for name, (pts, vs, rg, mp, mv) in voxel_cases().items():
        vox, coors, num = p2v(pts, np.array(vs, np.float32), np.array(rg, np.float32), mp, True, mv)
        np.savez_compressed(os.path.join(HERE, f"voxelize_{name}.npz"), points=pts,
                            voxel_size=np.array(vs, np.float32), coors_range=np.array(rg, np.float32),
                            max_points=mp, max_voxels=mv, voxels=vox, coors=coors, num_points=num)
        print(f"voxelize_{name}: N={len(pts)} M={len(coors)} max_pts={num.max() if len(num) else 0}")


def dense_spconv(feats, coors, shape, batch, weight, kind, stride, pad):
    """R1: scatter to a zero grid -> F.conv3d -> read back at the active output sites."""
    import torch
    import torch.nn.functional as F
    d, h, w = shape
    cin = feats.shape[1]
    grid = torch.zeros(batch, cin, d, h, w, dtype=torch.float64)
    occ = torch.zeros(batch, 1, d, h, w, dtype=torch.float64)
    c = torch.from_numpy(coors).long()
    grid[c[:, 0], :, c[:, 1], c[:, 2], c[:, 3]] = torch.from_numpy(feats).double()
    occ[c[:, 0], 0, c[:, 1], c[:, 2], c[:, 3]] = 1
    wt = torch.from_numpy(weight).double().permute(4, 3, 0, 1, 2).contiguous()   # [Cout,Cin,kd,kh,kw]
    ks = weight.shape[:3]
    if kind == "subm":
        out = F.conv3d(grid, wt, padding=[k // 2 for k in ks])
        oc = c
    else:
        out = F.conv3d(grid, wt, stride=stride, padding=pad)
        act = F.conv3d(occ, torch.ones(1, 1, *ks, dtype=torch.float64), stride=stride, padding=pad) > 0
        oc = act[:, 0].nonzero()                                                  # ascending (b,z,y,x)
    vals = out[oc[:, 0], :, oc[:, 1], oc[:, 2], oc[:, 3]]
    return vals.float().numpy(), oc.int().numpy(), tuple(out.shape[2:])


def make_spconv_goldens():
    rng = np.random.default_rng(3)
    cases = {
        # name: (shape, batch, n_active, cin, cout, kind, ksize, stride, pad)
        "subm_5_16": ((9, 24, 24), 2, 300, 5, 16, "subm", (3, 3, 3), 1, 1),
        "subm_32_32": ((5, 16, 20), 2, 400, 32, 32, "subm", (3, 3, 3), 1, 1),
        "down_16_32_s2p1": ((9, 24, 24), 2, 300, 16, 32, "sparse", (3, 3, 3), (2, 2, 2), (1, 1, 1)),
        "down_16_24_p011": ((11, 20, 20), 2, 350, 16, 24, "sparse", (3, 3, 3), (2, 2, 2), (0, 1, 1)),
        "extra_k311_s211": ((5, 12, 12), 2, 260, 32, 32, "sparse", (3, 1, 1), (2, 1, 1), (0, 0, 0)),
    }
    for name, (shape, batch, n, cin, cout, kind, ks, st, pd) in cases.items():
        vol = batch * shape[0] * shape[1] * shape[2]
        lin = np.sort(rng.choice(vol, size=n, replace=False))
        if kind == "subm":
            lin = rng.permutation(lin)               # SubM must keep an arbitrary input row order
        x = lin % shape[2]; y = (lin // shape[2]) % shape[1]
        z = (lin // (shape[2] * shape[1])) % shape[0]; b = lin // (shape[2] * shape[1] * shape[0])
        coors = np.stack([b, z, y, x], 1).astype(np.int32)
        feats = rng.normal(size=(n, cin)).astype(np.float32)
        weight = (rng.normal(size=(*ks, cin, cout)) / np.sqrt(cin * np.prod(ks))).astype(np.float32)
        out, oc, oshape = dense_spconv(feats, coors, shape, batch, weight, kind, st, pd)
        np.savez_compressed(os.path.join(HERE, f"spconv_{name}.npz"), feats=feats, coors=coors,
                            shape=np.array(shape, np.int32), batch=batch, weight=weight, kind=kind,
                            ksize=np.array(ks, np.int32), stride=np.array(st, np.int32) if kind != "subm" else np.array([1, 1, 1], np.int32),
                            pad=np.array(pd, np.int32) if kind != "subm" else np.array([1, 1, 1], np.int32),
                            out=out, out_coors=oc, out_shape=np.array(oshape, np.int32))
        print(f"spconv_{name}: N_in={n} N_out={len(oc)} out_shape={oshape}")


def reference_dense_modules():
    """Import the reference's own S2D_RPN / CenterHead through the namespace shim of SURVEY.md App. G."""
    import logging
    import types
    R = REF

    def pkg(name, path):
        m = types.ModuleType(name); m.__path__ = [path]; sys.modules[name] = m; return m

    def stub(name, **kw):
        m = types.ModuleType(name); m.__dict__.update(kw); sys.modules[name] = m; return m
    for n, p_ in [("det3d", "det3d"), ("det3d.models", "det3d/models"), ("det3d.torchie", "det3d/torchie"),
                  ("det3d.models.necks", "det3d/models/necks"), ("det3d.models.bbox_heads", "det3d/models/bbox_heads"),
                  ("det3d.models.readers", "det3d/models/readers"), ("det3d.models.losses", "det3d/models/losses"),
                  ("det3d.torchie.cnn", "det3d/torchie/cnn"), ("det3d.utils", "det3d/utils")]:
        pkg(n, f"{R}/{p_}")
    stub("det3d.torchie.trainer", load_checkpoint=lambda *a, **k: None)
    import det3d.torchie.cnn.weight_init as wi
    sys.modules["det3d.torchie.cnn"].__dict__.update({k: getattr(wi, k) for k in dir(wi) if k.endswith("_init")})
    sys.modules["det3d.torchie"].is_str = lambda x: isinstance(x, str)
    import det3d.utils.registry as reg
    sys.modules["det3d.utils"].Registry, sys.modules["det3d.utils"].build_from_cfg = reg.Registry, reg.build_from_cfg
    stub("det3d.utils.dist")
    sys.modules["det3d.utils.dist"].dist_common = stub("det3d.utils.dist.dist_common", get_world_size=lambda: 1)
    stub("det3d.core", box_torch_ops=types.SimpleNamespace())
    stub("det3d.core.utils")
    stub("det3d.core.utils.circle_nms_jit", circle_nms=None)
    stub("det3d.core.utils.center_utils", _transpose_and_gather_feat=None)
    import det3d.models.registry  # noqa: F401
    import det3d.models.utils  # noqa: F401
    import det3d.models.necks.rpn as rpn
    import det3d.models.bbox_heads.center_head as ch
    return rpn, ch, logging.getLogger("ref")


NECK_CFG = dict(layer_nums=[5, 5], ds_layer_strides=[1, 2], ds_num_filters=[128, 256], us_layer_strides=[1, 2],
                us_num_filters=[256, 256], num_input_features=256)
HEAD_CFG = dict(in_channels=512, tasks=[dict(num_class=3, class_names=["VEHICLE", "PEDESTRIAN", "CYCLIST"])],
                dataset="waymo", weight=2, code_weights=[1.0] * 8,
                common_heads={"reg": (2, 2), "height": (1, 2), "dim": (3, 2), "rot": (2, 2)})


def bev_input(seed, batch=1, occupancy=0.25):
    """Sparse-looking BEV map like the backbone's dense() output: ~25 % occupied cells, ReLU-positive values."""
    rng = np.random.default_rng(seed)
    occ = rng.uniform(size=(batch, 1, 188, 188)) < occupancy
    x = np.abs(rng.normal(0, 1.0, size=(batch, 256, 188, 188))) * occ
    return x.astype(np.float32)


def make_neck_head_goldens():
    import logging
    import torch
    from oracle import neck_head as NH
    from sparse2dense_b200 import registry
    rpn, ch, logger = reference_dense_modules()
    torch.manual_seed(0)
    ours_neck = registry.build_neck(dict(type="S2D_RPN", logger=logging.getLogger("x"), **NECK_CFG))
    ours_head = registry.build_head(dict(type="CenterHead", **HEAD_CFG))
    ns, hs = synth.random_module_state(ours_neck, 11), synth.random_module_state(ours_head, 12)
    ref_neck = rpn.S2D_RPN(logger=logger, **NECK_CFG).eval()
    ref_head = ch.CenterHead(logger=logger, **HEAD_CFG).eval()
    for ref, st in ((ref_neck, ns), (ref_head, hs)):
        res = ref.load_state_dict({k: torch.from_numpy(v) for k, v in st.items()}, strict=False)
        assert not res.unexpected_keys, res.unexpected_keys
        assert all(k.endswith("num_batches_tracked") for k in res.missing_keys), res.missing_keys
    x = torch.from_numpy(bev_input(21))
    with torch.no_grad():
        rx, _, _, _, _, rfa, rfb = ref_neck(x)
        rh = ref_head(rx)[0]
        ox, ofa, ofb = NH.s2d_rpn_forward(ns, x)
        oh = NH.center_head_forward(hs, ox)[0]
    rng = np.random.default_rng(5)
    out = {}
    for name, r, o in [("x", rx, ox), ("F_S_a", rfa, ofa), ("F_S_b", rfb, ofb)] + [(h, rh[h], oh[h]) for h in rh]:
        r, o = r.numpy(), o.numpy()
        err = np.abs(r - o).max() / np.abs(r).max()
        print(f"neck/head {name}: shape {r.shape} max|ref| {np.abs(r).max():.3f} oracle-vs-reference rel err {err:.2e}")
        assert err < 1e-5, name
        idx = rng.choice(r.size, size=min(4000, r.size), replace=False)
        out[name + "_idx"] = idx.astype(np.int64)
        out[name + "_val"] = r.reshape(-1)[idx]
        out[name + "_absmax"] = np.float32(np.abs(r).max())
    np.savez_compressed(os.path.join(HERE, "neck_head_s2d.npz"), neck_seed=11, head_seed=12, input_seed=21, **out)


class Cfg(dict):
    """Attribute + .get access like the reference's ConfigDict."""
    __getattr__ = dict.__getitem__


TEST_CFG = Cfg(post_center_limit_range=[-80, -80, -10.0, 80, 80, 10.0], max_per_img=4096,
               nms=Cfg(use_rotate_nms=True, use_multi_class_nms=False, nms_pre_max_size=4096, nms_post_max_size=500,
                       nms_iou_threshold=0.7),
               score_threshold=0.1, pc_range=[-75.2, -75.2], out_size_factor=8, voxel_size=[0.1, 0.1])


def ref_iou_lib():
    import ctypes
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref_iou3d.so"))
    f32p = ctypes.POINTER(ctypes.c_float)

    def matrix(a, b):
        a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
        out = np.empty((len(a), len(b)), np.float32)
        lib.ref_iou_bev_matrix(a.ctypes.data_as(f32p), len(a), b.ctypes.data_as(f32p), len(b), out.ctypes.data_as(f32p))
        return out
    return matrix


def box_pairs(seed, n=160, m=240):
    rng = np.random.default_rng(seed)
    a = np.stack([rng.uniform(-20, 20, n), rng.uniform(-20, 20, n), rng.uniform(-1, 1, n), rng.uniform(0.6, 6, n),
                  rng.uniform(0.5, 2.5, n), rng.uniform(1, 2, n), rng.uniform(-np.pi, np.pi, n)], 1).astype(np.float32)
    b = a[rng.integers(0, n, m)].copy()
    b[:, :2] += rng.normal(0, 0.7, (m, 2)).astype(np.float32)
    b[:, 3:5] *= rng.uniform(0.8, 1.2, (m, 2)).astype(np.float32)
    b[:, 6] += rng.normal(0, 0.4, m).astype(np.float32)
    b[:16] = a[:16]                                    # identical boxes
    b[16:32, 6] = a[16:32, 6] + np.float32(np.pi / 2)  # right angles about the same centre
    b[32:40, 6] = 0; b[32:40, :2] = a[32:40, :2]       # axis aligned
    return a, b.astype(np.float32)


def head_maps(seed, B=2, H=48, W=44, n_obj=70):
    """CenterHead output maps (NCHW) with object-like peaks: 3x3 neighbourhoods of confident, overlapping boxes."""
    rng = np.random.default_rng(seed)
    hm = rng.normal(-4.5, 0.6, (B, 3, H, W))
    reg = rng.uniform(0, 1, (B, 2, H, W))
    height = rng.normal(0.5, 0.3, (B, 1, H, W))
    dim = np.log(rng.uniform(0.5, 2.0, (B, 3, H, W)))
    ang = rng.uniform(-np.pi, np.pi, (B, 1, H, W))
    for b in range(B):
        for _ in range(n_obj):
            cy, cx, cls = rng.integers(1, H - 1), rng.integers(1, W - 1), rng.integers(0, 3)
            size = np.log([(4.6, 2.0, 1.7), (0.9, 0.8, 1.7), (1.8, 0.8, 1.6)][cls])
            yaw = rng.uniform(-np.pi, np.pi)
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    if rng.uniform() < 0.35:
                        continue
                    y, x = cy + dy, cx + dx
                    hm[b, cls, y, x] = rng.normal(1.0 - 1.2 * (abs(dy) + abs(dx)), 0.5)
                    reg[b, :, y, x] = np.clip([0.5 - 0.8 * dx + rng.normal(0, 0.1), 0.5 - 0.8 * dy + rng.normal(0, 0.1)], -1, 2)
                    dim[b, :, y, x] = size + rng.normal(0, 0.05, 3)
                    ang[b, 0, y, x] = yaw + rng.normal(0, 0.05)
    rot = np.concatenate([np.sin(ang), np.cos(ang)], 1)
    f = np.float32
    return dict(reg=reg.astype(f), height=height.astype(f), dim=dim.astype(f), rot=rot.astype(f), hm=hm.astype(f))


def make_predict_goldens():
    """IoU pairs from the reference's CPU rotated-IoU code (oracle/_ref) and CenterHead.predict outputs from the
    reference's own predict / post_processing (imported through the shim) with rotate_nms_pcdet's CUDA-only nms_gpu
    replaced by its protocol (iou3d_nms.cpp:118-131) over the same reference IoU."""
    import torch
    from oracle import ref_ops as R
    iou = ref_iou_lib()
    a, b = box_pairs(3)
    r = iou(a, b)
    o = R.iou_bev_matrix(a, b)
    print(f"iou pairs: {r.size} pairs, {(r > 0).sum()} overlapping, {(r > 0.7).sum()} above 0.7, "
          f"oracle == reference bitwise: {bool((r.view(np.uint32) == o.view(np.uint32)).all())}")
    assert (r.view(np.uint32) == o.view(np.uint32)).all()
    np.savez_compressed(os.path.join(HERE, "iou_bev_pairs.npz"), boxes_a=a, boxes_b=b, iou=r)

    rpn, ch, logger = reference_dense_modules()
    margins = []

    def rotate_nms_pcdet(boxes, scores, thresh, pre_maxsize=None, post_max_size=None):    # box_torch_ops.py:449-470
        order = scores.sort(dim=0, descending=True, stable=True)[1]
        if pre_maxsize is not None:
            order = order[:pre_maxsize]
        bx = boxes[order].contiguous().numpy()
        n = len(bx)
        m = iou(bx, bx) if n else np.zeros((0, 0), np.float32)
        removed, keep = np.zeros(n, bool), []
        for i in range(n):                                                              # iou3d_nms.cpp:118-131
            if removed[i]:
                continue
            keep.append(i)
            removed[i + 1:] |= m[i, i + 1:] > thresh
            if n - i - 1 > 0:
                margins.append(float(np.abs(m[i, i + 1:] - thresh).min()))
        sel = order[torch.as_tensor(keep, dtype=torch.long)]
        return sel[:post_max_size] if post_max_size is not None else sel
    ch.box_torch_ops.rotate_nms_pcdet = rotate_nms_pcdet
    head = ch.CenterHead(logger=logger, **HEAD_CFG).eval()
    maps = head_maps(9)
    preds = [{k: torch.from_numpy(v.copy()) for k, v in maps.items()}]
    out = head.predict({"metadata": []}, preds, TEST_CFG)
    save = {"in_" + k: v for k, v in maps.items()}
    for i, d in enumerate(out):
        save[f"boxes_{i}"] = d["box3d_lidar"].numpy()
        save[f"scores_{i}"] = d["scores"].numpy()
        save[f"labels_{i}"] = d["label_preds"].numpy()
        print(f"predict sample {i}: {len(d['scores'])} detections, top score {float(d['scores'].max()):.3f}")
    save["nms_margin"] = np.float32(min(margins))
    print(f"predict: min |IoU - thr| over the sweep = {min(margins):.2e}")
    # our numpy/C restatement must reproduce it
    nhwc = {k: np.ascontiguousarray(v.transpose(0, 2, 3, 1)) for k, v in maps.items()}
    boxes, hm = R.centerhead_decode(nhwc, 8, [0.1, 0.1], [-75.2, -75.2])
    for i in range(len(out)):
        det, _ = R.post_processing(boxes[i], hm[i], 0.1, TEST_CFG.post_center_limit_range, 0.7, 4096, 500)
        assert np.array_equal(det["label_preds"], save[f"labels_{i}"]), "labels differ"
        assert np.allclose(det["box3d_lidar"], save[f"boxes_{i}"], rtol=1e-5, atol=1e-6)
        assert np.allclose(det["scores"], save[f"scores_{i}"], rtol=1e-6, atol=1e-7)
    print("predict: oracle restatement reproduces the reference predict() output")
    np.savez_compressed(os.path.join(HERE, "centerhead_predict.npz"), **save)


ROI_CFG = Cfg(CLASS_AGNOSTIC=True, SHARED_FC=[256, 256], CLS_FC=[256, 256], REG_FC=[256, 256], DP_RATIO=0.3,
              TARGET_CONFIG=Cfg(ROI_PER_IMAGE=128, FG_RATIO=0.5, SAMPLE_ROI_BY_EACH_CLASS=True, CLS_SCORE_TYPE="roi_iou",
                                CLS_FG_THRESH=0.75, CLS_BG_THRESH=0.25, CLS_BG_THRESH_LO=0.1, HARD_BG_RATIO=0.8,
                                REG_FG_THRESH=0.55),
              LOSS_CONFIG=Cfg(CLS_LOSS="BinaryCrossEntropy", REG_LOSS="L1",
                              LOSS_WEIGHTS={"rcnn_cls_weight": 1.0, "rcnn_reg_weight": 1.0, "code_weights": [1.0] * 7}))


def make_second_stage_goldens():
    """Second stage through the reference's own TwoStageDetector methods / BEVFeatureExtractor / RoIHead on CPU."""
    import types
    import torch
    from oracle import ref_ops as R
    rpn, ch, logger = reference_dense_modules()

    def stub(name, **kw):
        m = types.ModuleType(name); m.__dict__.update(kw); sys.modules[name] = m; return m
    for n, p_ in [("det3d.core", "det3d/core"), ("det3d.core.bbox", "det3d/core/bbox"), ("det3d.core.utils", "det3d/core/utils"),
                  ("det3d.ops", "det3d/ops"), ("det3d.ops.nms", "det3d/ops/nms"), ("det3d.ops.iou3d_nms", "det3d/ops/iou3d_nms"),
                  ("det3d.models.detectors", "det3d/models/detectors"), ("det3d.models.second_stage", "det3d/models/second_stage"),
                  ("det3d.models.roi_heads", "det3d/models/roi_heads"),
                  ("det3d.models.roi_heads.target_assigner", "det3d/models/roi_heads/target_assigner")]:
        m = types.ModuleType(n); m.__path__ = [f"{REF}/{p_}"]; sys.modules[n] = m
    stub("cv2")
    sys.modules.pop("det3d.core.utils.center_utils", None)          # drop the stub of reference_dense_modules()
    stub("det3d.core.utils.circle_nms_jit", circle_nms=None)
    stub("det3d.ops.nms.nms_cpu", rotate_nms_cc=None)
    stub("det3d.ops.nms.nms_gpu", nms_gpu=None, rotate_iou_gpu=None, rotate_nms_gpu=None)
    stub("det3d.ops.iou3d_nms.iou3d_nms_utils", boxes_iou3d_gpu=None)
    stub("det3d.ops.iou3d_nms.iou3d_nms_cuda")
    import det3d.core.bbox.box_torch_ops as bto
    sys.modules["det3d.core.bbox"].box_torch_ops = bto
    sys.modules["det3d.core"].box_torch_ops = bto
    import det3d.core.utils.center_utils  # noqa: F401
    stub("det3d.models.detectors.base", BaseDetector=torch.nn.Module)
    sys.modules["det3d.models"].builder = types.SimpleNamespace()
    import det3d.models.detectors.two_stage as ts
    import det3d.models.second_stage.bird_eye_view as bev_mod
    import det3d.models.roi_heads.roi_head as rh

    rng = np.random.default_rng(17)
    B, C, H, W, P = 2, 64, 24, 20, 500
    pc_start, voxel, stride = [-75.2, -75.2], [0.1, 0.1], 8
    bev = np.abs(rng.normal(0, 1, (B, C, H, W))).astype(np.float32)
    preds = []
    for b in range(B):
        n = [37, 120][b]
        xy = np.stack([rng.uniform(-76.5, -75.2 + W * 0.8 + 1.0, n), rng.uniform(-76.5, -75.2 + H * 0.8 + 1.0, n)], 1)
        boxes = np.concatenate([xy, rng.normal(0.5, 0.3, (n, 1)), rng.uniform(0.5, 5, (n, 3)), rng.uniform(-np.pi, np.pi, (n, 1))], 1)
        preds.append(dict(box3d_lidar=torch.from_numpy(boxes.astype(np.float32)),
                          scores=torch.from_numpy(rng.uniform(0.1, 1, n).astype(np.float32)),
                          label_preds=torch.from_numpy(rng.integers(0, 3, n).astype(np.int64))))
    torch.manual_seed(4)
    roi = rh.RoIHead(input_channels=C * 5, model_cfg=ROI_CFG, code_size=7).eval()
    with torch.no_grad():
        for m in roi.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.normal_(0, 0.1); m.running_var.uniform_(0.5, 1.5); m.weight.uniform_(0.5, 1.5); m.bias.normal_(0, 0.1)
        roi.reg_layers[-1].weight.normal_(0, 0.05); roi.reg_layers[-1].bias.normal_(0, 0.1)
    det = ts.TwoStageDetector.__new__(ts.TwoStageDetector)
    torch.nn.Module.__init__(det)
    det.NMS_POST_MAXSIZE, det.num_point, det.roi_head = P, 5, roi
    det.second_stage = torch.nn.ModuleList([bev_mod.BEVFeatureExtractor(pc_start, voxel, stride)])
    example = {"bev_feature": torch.from_numpy(bev).permute(0, 2, 3, 1).contiguous(), "metadata": [None] * B}
    with torch.no_grad():
        centers = det.get_box_center(preds)
        feats = [m.forward(example, centers, 5) for m in det.second_stage]
        example = det.reorder_first_stage_pred_and_feature(first_pred=preds, example=example, features=feats)
        out = det.post_process(roi(example, training=False))
    save = dict(bev=bev, roi_features_0=feats[0][0].numpy(), roi_features_1=feats[0][1].numpy())
    state = {k: v.numpy() for k, v in roi.state_dict().items()}
    save.update({"roi." + k: v for k, v in state.items()})
    for b in range(B):
        save[f"in_boxes_{b}"] = preds[b]["box3d_lidar"].numpy(); save[f"in_scores_{b}"] = preds[b]["scores"].numpy()
        save[f"in_labels_{b}"] = preds[b]["label_preds"].numpy()
        save[f"boxes_{b}"] = out[b]["box3d_lidar"].numpy(); save[f"scores_{b}"] = out[b]["scores"].numpy()
        save[f"labels_{b}"] = out[b]["label_preds"].numpy()
        # oracle restatement must reproduce the reference modules
        f = R.roi_features(np.ascontiguousarray(bev[b].transpose(1, 2, 0)), save[f"in_boxes_{b}"], pc_start, voxel, stride)
        e1 = np.abs(f - save[f"roi_features_{b}"]).max() / np.abs(save[f"roi_features_{b}"]).max()
        cls, reg = R.roi_head_forward(state, f)
        ob, os_ = R.roi_refine(save[f"in_boxes_{b}"], save[f"in_scores_{b}"], cls, reg)
        e2, e3 = np.abs(ob - save[f"boxes_{b}"]).max(), np.abs(os_ - save[f"scores_{b}"]).max()
        print(f"second stage sample {b}: {len(ob)} rois, oracle-vs-reference: features {e1:.2e}, boxes {e2:.2e}, scores {e3:.2e}")
        assert e1 < 1e-5 and e2 < 1e-4 and e3 < 1e-5 and np.array_equal(save[f"labels_{b}"], save[f"in_labels_{b}"])
    np.savez_compressed(os.path.join(HERE, "two_stage.npz"), **save)


PP_VOXEL, PP_RANGE = (0.32, 0.32, 6.0), (-74.88, -74.88, -2, 74.88, 74.88, 4.0)
PP_RPN = dict(layer_nums=[3, 5, 5], ds_layer_strides=[1, 2, 2], ds_num_filters=[64, 128, 256], us_layer_strides=[1, 2, 4],
              us_num_filters=[128, 128, 128], num_input_features=64)


def make_pillar_goldens():
    """PillarFeatureNet -> PointPillarsScatter_S2D -> RPN([3,5,5]) through the reference's own modules (shim), on the
    pillars of one synthetic scene voxelized by the reference's numba voxelizer; sampled outputs are committed."""
    import logging
    import torch
    from oracle import pillars as OP
    from sparse2dense_b200 import registry
    rpn, ch, logger = reference_dense_modules()
    import det3d.models.readers.pillar_encoder as pe
    pco = ref_voxelizer()
    cloud = synth.lidar_scene(2000)
    voxels, coors, num = pco(cloud, np.array(PP_VOXEL, np.float32), np.array(PP_RANGE, np.float32), 20, True, 32000)
    coors4 = np.concatenate([np.zeros((len(coors), 1), np.int32), coors], 1)
    print(f"pillars: {len(voxels)} of one scene, grid 468 x 468")
    ours_reader = registry.build_reader(dict(type="PillarFeatureNet", num_filters=[64, 64], num_input_features=5,
                                             with_distance=False, voxel_size=PP_VOXEL, pc_range=PP_RANGE))
    ours_bb = registry.build_backbone(dict(type="PointPillarsScatter_S2D", ds_factor=1))
    ours_neck = registry.build_neck(dict(type="RPN", logger=logging.getLogger("x"), **PP_RPN))
    rs, bs, ns = (synth.random_module_state(m, sd) for m, sd in ((ours_reader, 31), (ours_bb, 32), (ours_neck, 33)))
    ref_reader = pe.PillarFeatureNet(num_filters=[64, 64], num_input_features=5, with_distance=False, voxel_size=PP_VOXEL,
                                     pc_range=PP_RANGE).eval()
    ref_bb = pe.PointPillarsScatter_S2D(ds_factor=1).eval()
    ref_neck = rpn.RPN(logger=logger, **PP_RPN).eval()
    for ref, st in ((ref_reader, rs), (ref_bb, bs), (ref_neck, ns)):
        res = ref.load_state_dict({k: torch.from_numpy(v) for k, v in st.items()}, strict=False)
        assert not res.unexpected_keys, res.unexpected_keys
        assert all(k.endswith("num_batches_tracked") for k in res.missing_keys), res.missing_keys
    with torch.no_grad():
        v, c, n = torch.from_numpy(voxels), torch.from_numpy(coors4), torch.from_numpy(num)
        rf = ref_reader(v, n, c)
        rfa, rfb, _, _ = ref_bb(rf, c, 1, [468, 468, 1])
        rx = ref_neck(rfa)
        of = OP.pfn_forward(rs, voxels, num, coors4, PP_VOXEL, PP_RANGE)
        ofa, ofb = OP.scatter_s2d_forward(bs, of, coors4, 1, 468, 468)
        ox = OP.rpn_forward(ns, ofa, PP_RPN["layer_nums"], PP_RPN["ds_layer_strides"], PP_RPN["us_layer_strides"])
    rng = np.random.default_rng(6)
    out = {}
    for name, r, o in [("pfn", rf, of), ("F_S_a", rfa, ofa), ("F_S_b", rfb, ofb), ("x", rx, ox)]:
        r, o = r.numpy(), o.numpy()
        err = np.abs(r - o).max() / np.abs(r).max()
        print(f"pillars {name}: shape {r.shape} max|ref| {np.abs(r).max():.3f} oracle-vs-reference rel err {err:.2e}")
        assert err < 1e-5, name
        idx = rng.choice(r.size, size=min(4000, r.size), replace=False)
        out[name + "_idx"] = idx.astype(np.int64)
        out[name + "_val"] = r.reshape(-1)[idx]
        out[name + "_absmax"] = np.float32(np.abs(r).max())
    np.savez_compressed(os.path.join(HERE, "pillars_s2d.npz"), scene_seed=2000, reader_seed=31, backbone_seed=32,
                        neck_seed=33, n_pillars=len(voxels), coors_checksum=np.int64(coors4.astype(np.int64).sum()), **out)


sys.path.insert(0, os.path.join(ROOT, "tests"))
from loss_common import loss_inputs  # noqa: E402


MG_TASKS = [dict(num_class=3, class_names=["VEHICLE", "PEDESTRIAN", "CYCLIST"])]
MG_ASSIGNER = dict(tasks=MG_TASKS, anchor_generators=[
    dict(type="anchor_generator_range", sizes=[2.08, 4.73, 1.77], anchor_ranges=[-74.88, -74.88, 0, 74.88, 74.88, 0],
         rotations=[0, 1.57], matched_threshold=0.55, unmatched_threshold=0.4, class_name="VEHICLE"),
    dict(type="anchor_generator_range", sizes=[0.84, 0.91, 1.74], anchor_ranges=[-74.88, -74.88, 0, 74.88, 74.88, 0],
         rotations=[0, 1.57], matched_threshold=0.5, unmatched_threshold=0.35, class_name="PEDESTRIAN"),
    dict(type="anchor_generator_range", sizes=[0.84, 1.81, 1.77], anchor_ranges=[-74.88, -74.88, 0, 74.88, 74.88, 0],
         rotations=[0, 1.57], matched_threshold=0.5, unmatched_threshold=0.3, class_name="CYCLIST")])


def reference_mg_modules():
    """The reference's own anchor generator (box_np_ops), box coder (box_torch_ops) and MultiGroupHead through the shim; the
    compiled NMS extensions they import at module level (det3d.ops.nms.*) are stubbed -- they are not called here."""
    import types
    reference_dense_modules()

    def stub(name, **kw):
        m = types.ModuleType(name); m.__dict__.update(kw); sys.modules[name] = m; return m

    def pkg(name, path):
        m = types.ModuleType(name); m.__path__ = [path]; sys.modules[name] = m; return m
    core = pkg("det3d.core", REF + "/det3d/core")
    pkg("det3d.core.bbox", REF + "/det3d/core/bbox")
    pkg("det3d.ops", REF + "/det3d/ops")
    pkg("det3d.ops.nms", REF + "/det3d/ops/nms")
    stub("det3d.ops.nms.nms_cpu", rotate_nms_cc=None)
    stub("det3d.ops.nms.nms_gpu", nms_gpu=None, rotate_iou_gpu=None, rotate_nms_gpu=None)
    import det3d.core.bbox.box_np_ops as bnp
    import det3d.core.bbox.box_torch_ops as bto
    core.box_torch_ops = bto
    import det3d.core.bbox.box_coders as bc
    import det3d.models.bbox_heads.mg_head as mg
    from det3d.models.registry import LOSSES            # the losses are constructed by __init__ but not used by forward
    for name in ("SigmoidFocalLoss", "WeightedSmoothL1Loss", "WeightedSoftmaxClassificationLoss"):
        if name not in LOSSES.module_dict:
            LOSSES.register_module(type(name, (object,), {"__init__": lambda self, **kw: None}))
    return bnp, bc, mg


def mg_input(seed, H, W):
    rng = np.random.default_rng(seed)
    return (rng.normal(size=(2, 128, H, W)) * (rng.random((2, 1, H, W)) < 0.4)).astype(np.float32)


def make_mg_head_goldens():
    """Anchors, Head.forward maps, decoded boxes and the pre-NMS selection of the reference's anchor head on a random
    128-channel BEV map (SECOND configs: RPN[5] 128 -> MultiGroupHead, 2 anchors x 3 classes per cell)."""
    import torch
    from oracle import mg_head as OM
    from sparse2dense_b200 import anchors as A
    from sparse2dense_b200 import registry
    bnp, bc, mg = reference_mg_modules()
    H, W = 40, 36
    # numpy >= 2 returns a tuple from meshgrid and the reference assigns into it (box_np_ops.py:915): hand it a list
    meshgrid = np.meshgrid
    bnp.np.meshgrid = lambda *a, **k: list(meshgrid(*a, **k))
    ref_parts = []
    for g in MG_ASSIGNER["anchor_generators"]:
        a = bnp.create_anchors_3d_range([1, H, W], g["anchor_ranges"], g["sizes"], g["rotations"], None, np.float32)
        ref_parts.append(a.reshape([*a.shape[:3], -1, a.shape[-1]]))
    ref_anchors = np.concatenate(ref_parts, axis=-2).reshape(-1, 7)
    bnp.np.meshgrid = meshgrid
    ours = A.task_anchors(MG_ASSIGNER, [1, H, W])[0]
    assert ours.shape == ref_anchors.shape and (ours.view(np.uint32) == ref_anchors.view(np.uint32)).all()
    coder = bc.GroundBox3dCoderTorch(False, False, n_dim=7)
    ref_head = mg.MultiGroupHead(mode="3d", in_channels=128, tasks=MG_TASKS, weights=[1], box_coder=coder,
                                 encode_background_as_zeros=True, loss_norm=dict(type="NormByNumPositives", pos_cls_weight=1.0,
                                                                                 neg_cls_weight=2.0),
                                 loss_cls=dict(type="SigmoidFocalLoss", alpha=0.25, gamma=2.0, loss_weight=1.0),
                                 use_sigmoid_score=True,
                                 loss_bbox=dict(type="WeightedSmoothL1Loss", sigma=3.0, code_weights=[1.0] * 7, codewise=True,
                                                loss_weight=2.0),
                                 encode_rad_error_by_sin=True,
                                 loss_aux=dict(type="WeightedSoftmaxClassificationLoss", name="direction_classifier",
                                               loss_weight=0.2), direction_offset=0.0).eval()
    mine = registry.build_head(dict(type="MultiGroupHead", mode="3d", in_channels=128, tasks=MG_TASKS, weights=[1],
                                    box_coder=A.build_box_coder(dict(type="ground_box3d_coder", n_dim=7, linear_dim=False,
                                                                     encode_angle_vector=False)),
                                    loss_aux=dict(type="x"), direction_offset=0.0))
    st = synth.random_module_state(mine, 41)
    res = ref_head.load_state_dict({k: torch.from_numpy(v) for k, v in st.items()}, strict=False)
    assert not res.unexpected_keys and not res.missing_keys, res
    x = torch.from_numpy(mg_input(8, H, W))
    with torch.no_grad():
        rp = ref_head(x)[0]
        op = OM.head_forward(st, x)
        dec = coder.decode_torch(rp["box_preds"].view(2, -1, 7), torch.from_numpy(ref_anchors)[None].repeat(2, 1, 1)).numpy()
    out = dict(head_seed=41, input_seed=8, H=H, W=W)
    srng = np.random.default_rng(9)

    def sample(name, arr):
        idx = srng.choice(arr.size, size=min(4000, arr.size), replace=False)
        out[name + "_idx"], out[name + "_val"] = idx.astype(np.int64), arr.reshape(-1)[idx]
        out[name + "_absmax"] = np.float32(np.abs(arr).max())
    sample("anchors", ref_anchors)
    sample("decoded", dec)
    for k in rp:
        e = float((rp[k] - op[k]).abs().max() / rp[k].abs().max())
        print(f"mg_head {k}: shape {tuple(rp[k].shape)} oracle-vs-reference rel err {e:.2e}")
        assert e < 1e-5
        sample(k, rp[k].numpy())
    od = OM.decode(rp["box_preds"].reshape(2, -1, 7).numpy(), ref_anchors[None])
    e = float(np.abs(od - dec).max() / np.abs(dec).max())
    print(f"mg_head decode: oracle-vs-reference rel err {e:.2e}")
    assert e < 1e-6
    np.savez_compressed(os.path.join(HERE, "mg_head.npz"), **out)


def make_loss_goldens():
    """FastFocalLoss / RegLoss of the reference (shim) + the trainer's distill expressions restated on torch CPU."""
    import torch
    import torch.nn.functional as F
    from oracle import ref_ops as R
    rpn, ch, logger = reference_dense_modules()
    sys.modules.pop("det3d.core.utils.center_utils", None)
    import types
    cu = types.ModuleType("det3d.core.utils.center_utils")

    def _tg(feat, ind):                      # center_utils.py:66-80 (the real file needs cv2 + numba)
        feat = feat.permute(0, 2, 3, 1).contiguous()
        feat = feat.view(feat.size(0), -1, feat.size(3))
        return feat.gather(1, ind.unsqueeze(2).expand(ind.size(0), ind.size(1), feat.size(2)))
    cu._transpose_and_gather_feat = _tg
    sys.modules["det3d.core.utils.center_utils"] = cu
    import det3d.models.losses.centernet_loss as cl
    cl._transpose_and_gather_feat = _tg          # the module may have been imported earlier with the stub
    save = {}
    for seed in (40, 41):
        d = loss_inputs(seed)
        t = {k: torch.from_numpy(v) for k, v in d.items()}
        out = torch.clamp(torch.sigmoid(t["hm_logits"]), min=1e-4, max=1 - 1e-4)
        hm = cl.FastFocalLoss()(out, t["gt_hm"], t["ind"], t["mask"], t["cat"])
        kd = cl.FastFocalLoss()(out, torch.sigmoid(t["t_logits"]), t["ind"], t["mask"], t["cat"])
        reg = cl.RegLoss()(t["box"], t["mask"], t["ind"], t["anno"])
        pred, gt = _tg(t["box"], t["ind"]), _tg(t["t_box"], t["ind"])                        # trainer.py:68-76
        m = t["mask"].float().unsqueeze(2)
        dl = (F.mse_loss(pred * m, gt * m, reduction="none") / (m.sum() + 1e-4)).transpose(2, 0).sum(dim=2).sum(dim=1)
        fs, fd = t["box"], t["t_box"]                                                        # trainer.py:783-789 on small maps
        inds = fd > 0
        s2d = F.mse_loss(fs[inds], fd[inds]) * 10 + F.mse_loss(fs[~inds], fd[~inds]) * 20
        inds = t["t_logits"] > 0
        s2d = s2d + F.mse_loss(t["hm_logits"][inds], t["t_logits"][inds]) * 5 + F.mse_loss(t["hm_logits"][~inds], t["t_logits"][~inds]) * 20
        ref = dict(hm=hm.numpy(), kd=kd.numpy(), reg=reg.numpy(), dl=dl.numpy(), s2d=s2d.numpy())
        o_out = np.clip(1 / (1 + np.exp(-d["hm_logits"])), 1e-4, 1 - 1e-4).astype(np.float32)
        o_t = (1 / (1 + np.exp(-d["t_logits"]))).astype(np.float32)
        got = dict(hm=R.fast_focal_loss(o_out, d["gt_hm"], d["ind"], d["mask"], d["cat"]),
                   kd=R.fast_focal_loss(o_out, o_t, d["ind"], d["mask"], d["cat"]),
                   reg=R.reg_loss(d["box"], d["mask"], d["ind"], d["anno"]),
                   dl=R.reg_loss(d["box"], d["mask"], d["ind"], d["t_box"], squared=True),
                   s2d=R.sparse2dense_loss(d["box"], d["t_box"], d["hm_logits"], d["t_logits"]))
        for k in ref:
            err = np.abs(np.asarray(got[k]) - ref[k]).max() / max(np.abs(ref[k]).max(), 1e-12)
            print(f"loss seed {seed} {k}: reference {np.asarray(ref[k]).reshape(-1)[:3]} oracle rel err {err:.1e}")
            assert err < 2e-5, k
            save[f"{seed}_{k}"] = np.asarray(ref[k], np.float32)
    np.savez_compressed(os.path.join(HERE, "losses.npz"), **save)


def random_gt(seed, n=80, num_cls=3):
    """Synthetic Waymo-like annotations: boxes f32 [n,9] (x,y,z,w,l,h,vx,vy,rot) incl. out-of-range centres, degenerate
    sizes and angles far outside [-pi, pi]; classes 1..num_cls in random order."""
    rng = np.random.default_rng(seed)
    xy = rng.uniform(-80, 80, (n, 2))
    z = rng.uniform(-1, 3, (n, 1))
    cls = rng.integers(1, num_cls + 1, n)
    base = np.array([[2.0, 4.6, 1.7], [0.8, 0.9, 1.7], [0.8, 1.8, 1.6]])[cls - 1]
    wlh = base * rng.uniform(0.6, 1.8, (n, 3))
    wlh[rng.integers(0, n, 2), 0] = 0.0                       # degenerate boxes are skipped (preprocess.py:585)
    vel = rng.normal(0, 3, (n, 2))
    rot = rng.uniform(-7, 7, (n, 1))
    return np.concatenate([xy, z, wlh, vel, rot], 1).astype(np.float32), cls.astype(np.int64)


def make_assign_goldens():
    """The reference's own AssignLabel.__call__ (det3d/datasets/pipelines/preprocess.py:479-653) executed from its source
    text with the real center_utils / limit_period, on synthetic annotations; also checks oracle/assign_label.py."""
    import ast
    import types
    from oracle import assign_label as OA

    def pkg(name, path):
        m = types.ModuleType(name); m.__path__ = [path]; sys.modules[name] = m; return m
    for n in [k for k in sys.modules if k == "det3d" or k.startswith("det3d.")]:
        sys.modules.pop(n)
    pkg("det3d", REF + "/det3d"); pkg("det3d.core", REF + "/det3d/core"); pkg("det3d.core.utils", REF + "/det3d/core/utils")
    import det3d.core.utils.center_utils as cu                  # real file: numba circle_nms + cv2 are importable here
    src = open(REF + "/det3d/datasets/pipelines/preprocess.py").read()
    tree = ast.parse(src)
    want = {"flatten", "merge_multi_group_label", "AssignLabel"}
    code = "\n\n".join(ast.get_source_segment(src, n) for n in tree.body
                       if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in want)
    ns = dict(np=np, draw_umich_gaussian=cu.draw_umich_gaussian, gaussian_radius=cu.gaussian_radius,
              box_np_ops=types.SimpleNamespace(limit_period=lambda val, offset=0.5, period=np.pi:
                                               val - np.floor(val / period + offset) * period),      # box_np_ops.py:360-361
              PIPELINES=types.SimpleNamespace(register_module=lambda c: c))
    exec(compile(code, "preprocess.py[AssignLabel]", "exec"), ns)
    tasks = [types.SimpleNamespace(num_class=3, class_names=["VEHICLE", "PEDESTRIAN", "CYCLIST"])]
    cfg = types.SimpleNamespace(out_size_factor=8, target_assigner=types.SimpleNamespace(tasks=tasks), gaussian_overlap=0.1,
                                max_objs=500, min_radius=2)
    assigner = ns["AssignLabel"](cfg=cfg)
    names = np.array(["VEHICLE", "PEDESTRIAN", "CYCLIST"])
    save = {}
    for seed, n in ((50, 80), (51, 200), (52, 0), (53, 499)):
        boxes, cls = random_gt(seed, n) if n else (np.zeros((0, 9), np.float32), np.zeros((0,), np.int64))
        res = dict(mode="train", type="WaymoDataset",
                   lidar=dict(voxels=dict(shape=np.array([1504, 1504, 40]), range=np.array(synth.WAYMO_RANGE, np.float32),
                                          size=np.array(synth.WAYMO_VOXEL, np.float32)),
                              annotations=dict(gt_boxes=boxes.copy(), gt_classes=cls.copy(), gt_names=names[cls - 1])))
        out, _ = assigner(res, {})
        t = out["lidar"]["targets"]
        mine = OA.assign_label(boxes, cls, [3], (1504, 1504), synth.WAYMO_RANGE, synth.WAYMO_VOXEL, 8, 0.1, 500, 2)
        for k in ("hm", "anno_box", "ind", "mask", "cat"):
            ref = np.asarray(t[k][0])
            got = mine[k][0]
            if k == "anno_box":
                assert np.allclose(got, ref, rtol=0, atol=2e-7 * max(1.0, np.abs(ref).max())), (seed, k)
            else:
                assert np.array_equal(got, ref), (seed, k, np.abs(got.astype(np.float64) - ref).max())
            save[f"{seed}_{k}"] = ref
        assert np.array_equal(mine["gt_boxes_and_cls"], t["gt_boxes_and_cls"])
        save[f"{seed}_gt_boxes_and_cls"] = t["gt_boxes_and_cls"]
        save[f"{seed}_boxes"], save[f"{seed}_classes"] = boxes, cls
        print(f"assign seed {seed}: {int(t['mask'][0].sum())} objects drawn, hm max {t['hm'][0].max():.3f}, oracle identical")
    np.savez_compressed(os.path.join(HERE, "assign_label.npz"), **save)


def make_pcr_loss_goldens():
    """KD_VoxelNet.mask_offset_loss (det3d/models/detectors/voxelnet.py:171-185) executed from the reference source on a small
    synthetic reconstruction target; inputs and outputs are stored (the GPU kernel works from the sparse voxel list)."""
    import ast
    import torch
    import torch.nn.functional as F
    src = open(REF + "/det3d/models/detectors/voxelnet.py").read()
    tree = ast.parse(src)
    fn = None
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name == "mask_offset_loss":
            fn = ast.get_source_segment(src, node)
    import textwrap
    ns = dict(F=F, torch=torch)
    exec(compile(textwrap.dedent(fn), "voxelnet.py[mask_offset_loss]", "exec"), ns)
    ref_fn = ns["mask_offset_loss"]
    from oracle import train_ref as TR
    save = {}
    for seed, (B, D, H, W, M) in ((60, (2, 4, 12, 10, 150)), (61, (1, 6, 9, 16, 300))):
        rng = np.random.default_rng(seed)
        cells = rng.choice(B * D * H * W, M, replace=False)
        b, r = np.divmod(cells, D * H * W)
        z, r = np.divmod(r, H * W)
        y, x = np.divmod(r, W)
        coors = np.stack([b, z, y, x], 1).astype(np.int32)
        grid = TR.voxel_grid(B, D, H, W, torch.zeros(1))
        centre = grid[torch.from_numpy(b).long(), :, torch.from_numpy(z).long(), torch.from_numpy(y).long(), torch.from_numpy(x).long()].numpy()
        feats = np.concatenate([centre + rng.uniform(-0.3, 0.3, (M, 3)), rng.uniform(0, 1, (M, 2))], 1).astype(np.float32)
        feats[:5, 0] = centre[:5, 0]                       # some exact-zero offset components (dropped by `gt != 0`)
        gt = TR.dense_from_voxels(torch.from_numpy(feats), torch.from_numpy(coors), B, (D, H, W))
        gen_offset = torch.from_numpy(rng.normal(0, 0.5, (B, 3, D, H, W)).astype(np.float32))
        gen_mask = torch.from_numpy(rng.normal(0, 2.0, (B, 1, D, H, W)).astype(np.float32))
        loss, com = ref_fn(None, gen_offset, gen_mask, gt, grid)
        mine = TR.mask_offset_loss(gen_offset, gen_mask, gt, grid)
        assert abs(float(mine[0]) - float(loss)) < 1e-6 and abs(float(mine[1]) - float(com)) < 1e-6
        save.update({f"{seed}_dims": np.array([B, D, H, W]), f"{seed}_coors": coors, f"{seed}_feats": feats,
                     f"{seed}_gen_offset": gen_offset.numpy(), f"{seed}_gen_mask": gen_mask.numpy(),
                     f"{seed}_mask_loss": np.float32(loss), f"{seed}_offset_loss": np.float32(com)})
        print(f"pcr loss seed {seed}: mask {float(loss):.6f} offset {float(com):.6f}")
    np.savez_compressed(os.path.join(HERE, "pcr_loss.npz"), **save)


def make_optim_goldens():
    """The reference's own OptimWrapper (det3d/solver/fastai_optim.py:118-174, true_wd / bn_wd as build_one_cycle_optimizer
    sets them, apis/train.py:168-186) over torch Adam, driven by the reference's OneCycle
    (det3d/solver/learning_schedules_fastai.py:7-95) with the Waymo config's hyper-parameters, plus clip_grad_norm_(35)
    (hooks/optimizer.py:15-21): parameter values after every step for fixed gradients."""
    import collections
    import collections.abc
    from functools import partial
    import torch
    from torch import nn
    collections.Iterable = collections.abc.Iterable            # the reference predates python 3.10
    def load(name, path):
        spec = importlib.util.spec_from_file_location(name, REF + path)
        m = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(m)
        return m
    fo = load("fastai_optim", "/det3d/solver/fastai_optim.py")
    ls = load("learning_schedules_fastai", "/det3d/solver/learning_schedules_fastai.py")
    torch.manual_seed(3)
    model = nn.Sequential(nn.Linear(7, 16), nn.BatchNorm1d(16), nn.ReLU(), nn.Linear(16, 5))
    opt = fo.OptimWrapper.create(partial(torch.optim.Adam, betas=(0.9, 0.99), amsgrad=0.0), 3e-3, [model], wd=0.01,
                                 true_wd=True, bn_wd=True)
    total = 20
    sched = ls.OneCycle(opt, total, 0.003, [0.95, 0.85], 10.0, 0.3)
    names = [k for k, _ in model.named_parameters()]
    save = {"names": np.array(names), "total_step": np.array(total)}
    for k, p in model.named_parameters():
        save["init_" + k] = p.detach().numpy().copy()
    rng = np.random.default_rng(5)
    lrs, moms, norms = [], [], []
    for step in range(12):
        sched.step(step)
        lrs.append(opt.lr); moms.append(opt.mom)
        scale = 40.0 if step % 3 == 0 else 0.5                  # every third step is clipped
        for k, p in model.named_parameters():
            g = (rng.normal(size=tuple(p.shape)) * scale).astype(np.float32)
            save[f"grad{step}_{k}"] = g
            p.grad = torch.from_numpy(g.copy())
        norms.append(float(torch.nn.utils.clip_grad_norm_(model.parameters(), 35.0)))
        opt.step()
        for k, p in model.named_parameters():
            save[f"param{step}_{k}"] = p.detach().numpy().copy()
    save["lr"], save["mom"], save["norm"] = np.array(lrs), np.array(moms), np.array(norms)
    np.savez_compressed(os.path.join(HERE, "optim.npz"), **save)
    print("optim golden: lr", lrs[:4], "mom", moms[:4], "norms", norms[:4])


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "assign":
        make_assign_goldens()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "pcr":
        make_pcr_loss_goldens()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "mg":
        make_mg_head_goldens()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "optim":
        make_optim_goldens()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "loss":
        make_loss_goldens()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "pp":
        make_pillar_goldens()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "second":
        make_second_stage_goldens()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "predict":
        make_predict_goldens()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "dense":
        make_neck_head_goldens()
        sys.exit(0)
    make_voxel_goldens()
    make_spconv_goldens()
    make_neck_head_goldens()
    make_predict_goldens()
    make_second_stage_goldens()
    make_pillar_goldens()
    make_loss_goldens()
    make_optim_goldens()
    make_assign_goldens()
    make_pcr_loss_goldens()
    make_mg_head_goldens()
