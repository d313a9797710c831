"""Seeded synthetic spinning-LiDAR scenes (Waymo-shaped).

The reference trains/evaluates on Waymo sweeps loaded as ``[N,5]`` float32 rows
``x, y, z, tanh(intensity), elongation`` (det3d/datasets/pipelines/loading.py:61-70).
There is no dataset on the build or GPU boxes, so every test and the benchmark use
this generator instead.  It is host-side numpy only: it
produces the *input* of the hot path and is never timed.

Scene model: a 64-beam sensor 2 m above a flat ground plane, a ring of vertical
cylinders that occlude rays, a far "wall" for rays that hit nothing, two returns
for a fraction of the rays.  The ray budget is tuned so that about 180 k points
fall inside the Waymo detection range ``[-75.2,-75.2,-2, 75.2,75.2,4]``.
"""
import numpy as np

WAYMO_RANGE = (-75.2, -75.2, -2.0, 75.2, 75.2, 4.0)
WAYMO_VOXEL = (0.1, 0.1, 0.15)
WAYMO_MAX_POINTS = 5
WAYMO_MAX_VOXELS = 150000


def lidar_scene(seed, n_beams=64, n_azimuth=2680, n_cylinders=120, second_return=0.10,
                sensor_height=2.0, shuffle=False):
    """Return one cloud ``float32 [N,5]``; deterministic in ``seed``."""
    rng = np.random.default_rng(seed)
    elev = np.deg2rad(np.linspace(-17.6, 2.4, n_beams))
    az = np.linspace(0.0, 2.0 * np.pi, n_azimuth, endpoint=False)
    az = az[None, :] + rng.uniform(0.0, 1e-3, size=(n_beams, 1))
    el = np.broadcast_to(elev[:, None], az.shape)
    dx, dy, dz = np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)

    # ground hit (z = 0 plane, sensor at z = sensor_height)
    with np.errstate(divide="ignore", invalid="ignore"):
        t_ground = np.where(dz < -1e-6, -sensor_height / dz, np.inf)
    # rays that hit nothing return from a far surface
    t_far = rng.uniform(30.0, 80.0, size=az.shape)
    t = np.minimum(t_ground, np.where(np.isfinite(t_ground), np.inf, t_far))
    t = np.where(np.isfinite(t), t, t_far)

    # vertical cylinders (cars / poles / trunks)
    cx = rng.uniform(-70.0, 70.0, n_cylinders)
    cy = rng.uniform(-70.0, 70.0, n_cylinders)
    cr = rng.uniform(0.5, 3.0, n_cylinders)
    ch = rng.uniform(1.0, 3.5, n_cylinders)
    hxy = np.sqrt(dx * dx + dy * dy)
    ux, uy = dx / hxy, dy / hxy
    for k in range(n_cylinders):
        if cx[k] * cx[k] + cy[k] * cy[k] < (cr[k] + 3.0) ** 2:
            continue                                    # keep the ego footprint free
        b = ux * cx[k] + uy * cy[k]
        disc = b * b - (cx[k] * cx[k] + cy[k] * cy[k] - cr[k] * cr[k])
        s = b - np.sqrt(np.maximum(disc, 0.0))          # horizontal distance to entry
        tt = s / hxy
        zhit = sensor_height + tt * dz
        hit = (disc > 0) & (s > 0) & (zhit >= 0.0) & (zhit <= ch[k]) & (tt < t)
        t = np.where(hit, tt, t)

    t = t * (1.0 + rng.normal(0.0, 0.002, size=t.shape))
    pts = np.stack([t * dx, t * dy, sensor_height + t * dz], axis=-1).reshape(-1, 3)

    # second returns: a little behind the first one along the same ray
    m2 = rng.uniform(size=t.shape) < second_return
    t2 = t[m2] + rng.uniform(0.3, 2.5, size=int(m2.sum()))
    p2 = np.stack([t2 * dx[m2], t2 * dy[m2], sensor_height + t2 * dz[m2]], axis=-1)
    pts = np.concatenate([pts, p2], axis=0)

    feats = rng.uniform(0.0, 1.0, size=(pts.shape[0], 2))
    out = np.concatenate([pts, feats], axis=1).astype(np.float32)
    if shuffle:
        rng.shuffle(out, axis=0)
    return np.ascontiguousarray(out)


def lidar_batch(cfg, batch, **kw):
    """Scenes ``seed = 1000*cfg + scene_idx``."""
    return [lidar_scene(1000 * cfg + i, **kw) for i in range(batch)]


def small_scene(seed, n_beams=16, n_azimuth=1250):
    """cfg-1 sized cloud (20 k rays)."""
    return lidar_scene(seed, n_beams=n_beams, n_azimuth=n_azimuth, n_cylinders=30, second_return=0.0)


def in_range_mask(points, pc_range=WAYMO_RANGE):
    lo = np.asarray(pc_range[:3], np.float32)
    hi = np.asarray(pc_range[3:], np.float32)
    return np.all((points[:, :3] >= lo) & (points[:, :3] < hi), axis=1)


def backbone_state(seed=0, num_input_features=5):
    """Seeded random weights for SpMiddleResNetFHD in the reference's state-dict format
    (det3d/models/backbones/scn.py:104-152; spconv weight layout [kD,kH,kW,Cin,Cout]) with
    non-trivial BatchNorm running statistics so that BN folding is exercised.
    There is no checkpoint on the build or GPU boxes."""
    rng = np.random.default_rng(seed)
    state = {}

    def conv(name, ks, cin, cout, bias):
        fan = cin * int(np.prod(ks))
        state[name + ".weight"] = (rng.normal(size=(*ks, cin, cout)) * np.sqrt(2.0 / fan)).astype(np.float32)
        if bias:
            state[name + ".bias"] = rng.uniform(-0.1, 0.1, cout).astype(np.float32)

    def bn(name, c):
        state[name + ".weight"] = rng.uniform(0.5, 1.5, c).astype(np.float32)
        state[name + ".bias"] = rng.normal(0, 0.1, c).astype(np.float32)
        state[name + ".running_mean"] = rng.normal(0, 0.1, c).astype(np.float32)
        state[name + ".running_var"] = rng.uniform(0.5, 1.5, c).astype(np.float32)

    def block(name, c):
        conv(name + ".conv1", (3, 3, 3), c, c, True); bn(name + ".bn1", c)
        conv(name + ".conv2", (3, 3, 3), c, c, True); bn(name + ".bn2", c)

    conv("conv_input.0", (3, 3, 3), num_input_features, 16, False); bn("conv_input.1", 16)
    block("conv1.0", 16); block("conv1.1", 16)
    for name, cin, cout in (("conv2", 16, 32), ("conv3", 32, 64), ("conv4", 64, 128)):
        conv(name + ".0", (3, 3, 3), cin, cout, False); bn(name + ".1", cout)
        block(name + ".3", cout); block(name + ".4", cout)
    conv("extra_conv.0", (3, 1, 1), 128, 128, False); bn("extra_conv.1", 128)
    return state


def random_module_state(module, seed):
    """Seeded, torch-RNG-independent random state dict for any nn.Module of this package (reference key names):
    conv / linear weights ~ N(0, 2/fan_in), biases small, BatchNorm / LayerNorm affine around 1 with non-trivial
    running statistics.  Used for the dense neck / head where no checkpoint is available."""
    rng = np.random.default_rng(seed)
    out = {}
    for key, t in sorted(module.state_dict().items()):
        shape = tuple(t.shape)
        name = key.rsplit(".", 1)[-1]
        if name == "num_batches_tracked":
            continue
        if name == "running_mean":
            v = rng.normal(0, 0.1, shape)
        elif name == "running_var":
            v = rng.uniform(0.5, 1.5, shape)
        elif name == "weight" and len(shape) >= 2 and not (len(shape) == 3 and "convnext" in key and ".1." in key):
            fan_in = int(np.prod(shape[1:])) if len(shape) > 2 else shape[1]
            if len(shape) == 4 and "decoder" in key and shape[2] == 4:
                fan_in = shape[0] * 4                       # ConvTranspose2d [Cin,Cout,4,4]: 4 taps reach an output
            v = rng.normal(0, np.sqrt(2.0 / max(fan_in, 1)), shape)
        elif name == "weight":
            v = rng.uniform(0.5, 1.5, shape)                # norm scale
        else:
            v = rng.normal(0, 0.05, shape)                  # biases
        out[key] = v.astype(np.float32)
    return out


# ------------------------------------------------------------------------------------------
# distillation training step (BASELINE configs[4]; configs/waymo/voxelnet/waymo_centerpoint_voxelnet_1x_distill.py)
# ------------------------------------------------------------------------------------------
WAYMO_TASKS = [dict(num_class=3, class_names=["VEHICLE", "PEDESTRIAN", "CYCLIST"])]
WAYMO_VOXEL_CFG = dict(range=list(WAYMO_RANGE), voxel_size=[0.1, 0.1, 0.15], max_points_in_voxel=5,
                       max_voxel_num=[150000, 200000], distillation=True)
WAYMO_TEST_CFG = dict(post_center_limit_range=[-80, -80, -10.0, 80, 80, 10.0],
                      nms=dict(use_rotate_nms=True, use_multi_class_nms=False, nms_pre_max_size=4096,
                               nms_post_max_size=500, nms_iou_threshold=0.7),
                      score_threshold=0.1, pc_range=[-75.2, -75.2], out_size_factor=8, voxel_size=[0.1, 0.1])


def distill_model_cfgs():
    """(teacher VoxelNet, student KD_VoxelNet) config dicts of the reference's Waymo distillation config (lines 17-75)."""
    import logging

    def det(kind, neck):
        return dict(type=kind, pretrained=None, reader=dict(type="VoxelFeatureExtractorV3", num_input_features=5),
                    backbone=dict(type="SpMiddleResNetFHD", num_input_features=5, ds_factor=8),
                    neck=dict(type=neck, layer_nums=[5, 5], ds_layer_strides=[1, 2], ds_num_filters=[128, 256],
                              us_layer_strides=[1, 2], us_num_filters=[256, 256], num_input_features=256,
                              logger=logging.getLogger(neck)),
                    bbox_head=dict(type="CenterHead", in_channels=512, tasks=WAYMO_TASKS, dataset="waymo", weight=2,
                                   code_weights=[1.0] * 8,
                                   common_heads={"reg": (2, 2), "height": (1, 2), "dim": (3, 2), "rot": (2, 2)}))
    return det("VoxelNet", "RPN"), det("KD_VoxelNet", "S2D_RPN")


def build_distill_models(device="cuda", precision=None, seed=0):
    """Teacher + student with seeded random weights (no checkpoints exist on the build / GPU boxes)."""
    import torch
    from . import registry
    t_cfg, s_cfg = distill_model_cfgs()
    models = []
    for i, cfg in enumerate((t_cfg, s_cfg)):
        m = registry.build_detector(cfg, train_cfg=None, test_cfg=WAYMO_TEST_CFG)
        m.backbone.load_state_dict({k: torch.as_tensor(v) for k, v in backbone_state(seed + i).items()})
        for j, sub in enumerate((m.neck, m.bbox_head)):
            sub.load_state_dict({k: torch.as_tensor(v) for k, v in random_module_state(sub, 100 * (seed + i) + 11 + j).items()},
                                strict=False)
        with torch.no_grad():
            m.bbox_head.tasks[0].hm[-1].bias.fill_(-2.19)
        m.to(device)
        if precision is not None:
            m.set_precision(precision)
        models.append(m)
    return models[0], models[1]


def random_targets(batch, H=188, W=188, max_objs=500, n_obj=60, num_cls=3, seed=0):
    """Synthetic CenterPoint targets in the layout ``AssignLabel`` produces (preprocess.py:489-653; lists with one entry per
    task): hm [B,C,H,W] with gaussian peaks, anno_box [B,max_objs,10], ind / cat i64 [B,max_objs], mask u8."""
    import torch
    rng = np.random.default_rng(seed)
    hm = np.zeros((batch, num_cls, H, W), np.float32)
    anno = np.zeros((batch, max_objs, 10), np.float32)
    ind = np.zeros((batch, max_objs), np.int64)
    mask = np.zeros((batch, max_objs), np.uint8)
    cat = np.zeros((batch, max_objs), np.int64)
    ys, xs = np.mgrid[0:H, 0:W]
    for b in range(batch):
        n = int(rng.integers(n_obj // 2, n_obj + 1))
        cells = rng.choice(H * W, n, replace=False)
        for k, cell in enumerate(cells):
            cy, cx, c = cell // W, cell % W, int(rng.integers(0, num_cls))
            r = float(rng.uniform(1.0, 3.0))
            g = np.exp(-((xs - cx) ** 2 + (ys - cy) ** 2) / (2 * r * r)).astype(np.float32)
            g[g < 1e-3] = 0
            hm[b, c] = np.maximum(hm[b, c], g)
            ind[b, k], mask[b, k], cat[b, k] = cell, 1, c
            rot = rng.uniform(-np.pi, np.pi)
            anno[b, k] = np.concatenate([rng.uniform(0, 1, 2), rng.uniform(-1, 2, 1), np.log(rng.uniform(0.5, 6, 3)),
                                         rng.normal(0, 1, 2), [np.sin(rot), np.cos(rot)]])
    t = torch.as_tensor
    return dict(hm=[t(hm)], anno_box=[t(anno)], ind=[t(ind)], mask=[t(mask)], cat=[t(cat)])


def distill_example(batch, cfg=1, device="cuda", small=False):
    """One synthetic distillation batch: the student's input is a thinned sweep of each scene, the teacher's dense /
    reconstruction input the full cloud (the reference's multi-sweep object-completed clouds are dataset products)."""
    import torch
    from .pipeline import Voxelization
    scenes = [small_scene(1000 * cfg + i) for i in range(batch)] if small else lidar_batch(cfg, batch)
    scenes = [s[in_range_mask(s)] for s in scenes]
    sparse = [s[::2] for s in scenes]
    vox = Voxelization(cfg=WAYMO_VOXEL_CFG)
    ex = vox(sparse, dense_points=scenes, reconstruction_points=scenes, mode="train", device=device)
    tg = random_targets(batch, seed=cfg)
    ex.update({k: [v[0].to(device)] for k, v in tg.items()})
    ex["metadata"] = [dict(token=f"synthetic_{cfg}_{i}") for i in range(batch)]
    return ex
