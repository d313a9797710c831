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


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "dense":
        make_neck_head_goldens()
        sys.exit(0)
    make_voxel_goldens()
    make_spconv_goldens()
    make_neck_head_goldens()
