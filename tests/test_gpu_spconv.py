"""GPU parity: rulebooks (bit-exact) and sparse convolution (fp32, 1e-3 relative per north_star;
observed ~1e-6) through the C ABI vs the oracle and the committed dense-conv3d goldens."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import ref_ops as R
from sparse2dense_b200 import ops, synth

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
SPC = sorted(glob.glob(os.path.join(GOLDEN, "spconv_*.npz")))
RTOL = 1e-3          # north_star tolerance (relative to the output scale)


def rel_err(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def gpu_rulebook(coors, batch, shape, kind, ks, st=1, pd=0):
    c = torch.from_numpy(np.ascontiguousarray(coors, np.int32)).cuda()
    idx = ops.build_grid_index(c, batch, tuple(int(v) for v in shape))
    if kind == "subm":
        tbl, pairs = ops.rulebook_subm(c, idx, ks, count_pairs=True)
        return c, tbl, int(pairs.item()), idx, tuple(shape)
    sc = ops.sparse_out_coords(c, c.shape[0], batch, tuple(int(v) for v in shape), ks, st, pd)
    oc = sc.coors
    tbl, pairs = ops.rulebook_sparse(oc, idx, ks, st, pd, count_pairs=True)
    return oc, tbl, int(pairs.item()), sc.index, sc.shape


@pytest.mark.parametrize("path", SPC, ids=[os.path.basename(p)[7:-4] for p in SPC])
def test_rulebook_and_conv_vs_goldens(path):
    g = np.load(path)
    kind = str(g["kind"])
    oc, tbl, pairs, _, oshape = gpu_rulebook(g["coors"], int(g["batch"]), g["shape"], kind, g["ksize"], g["stride"],
                                             g["pad"])
    np.testing.assert_array_equal(oc.cpu().numpy(), g["out_coors"])
    assert tuple(oshape) == tuple(g["out_shape"])
    if kind == "subm":
        rt, rp = R.rulebook_subm(g["coors"], g["shape"], g["ksize"])
    else:
        _, rt, _, rp = R.rulebook_sparse(g["coors"], g["shape"], g["ksize"], g["stride"], g["pad"])
    np.testing.assert_array_equal(tbl.cpu().numpy()[:, : rt.shape[1]], rt)          # rulebook bit-exact
    assert pairs == rp
    out = ops.spconv_fwd(torch.from_numpy(g["feats"]).cuda(), torch.from_numpy(g["weight"]).cuda(), tbl,
                         oc.shape[0])
    assert rel_err(out.cpu().numpy(), g["out"]) < 1e-5


@pytest.mark.parametrize("cin,cout", [(5, 16), (16, 16), (16, 32), (32, 32), (32, 64), (64, 64), (64, 128),
                                      (128, 128), (7, 12)])
@pytest.mark.parametrize("fused", [False, True])
def test_spconv_all_backbone_shapes(cin, cout, fused):
    """Every (Cin,Cout) of SpMiddleResNetFHD + a generic shape; ragged tile (N not a tile multiple)."""
    rng = np.random.default_rng(cin * 131 + cout)
    shape, batch, n = (7, 30, 30), 2, 1777
    lin = rng.permutation(rng.choice(batch * 7 * 900, n, replace=False))
    coors = np.stack([lin // 6300, (lin // 900) % 7, (lin // 30) % 30, lin % 30], 1).astype(np.int32)
    feats = rng.normal(size=(n, cin)).astype(np.float32)
    w = (rng.normal(size=(3, 3, 3, cin, cout)) / np.sqrt(27 * cin)).astype(np.float32)
    rt, _ = R.rulebook_subm(coors, shape, 3)
    c, tbl, _, _, _ = gpu_rulebook(coors, batch, shape, "subm", 3)
    np.testing.assert_array_equal(tbl.cpu().numpy(), rt)
    ref = R.spconv_fwd(feats, w, rt, wide=True)
    kw = {}
    if fused:
        scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
        shift = rng.normal(0, 0.2, cout).astype(np.float32)
        res = rng.normal(size=(n, cout)).astype(np.float32)
        ref = R.bn_act(ref, scale, shift, res, True)
        kw = dict(scale=torch.from_numpy(scale).cuda(), shift=torch.from_numpy(shift).cuda(),
                  residual=torch.from_numpy(res).cuda(), relu=True)
    out = ops.spconv_fwd(torch.from_numpy(feats).cuda(), torch.from_numpy(w).cuda(), tbl, n, **kw)
    assert rel_err(out.cpu().numpy(), ref) < 1e-5


def test_empty_and_single_row():
    dev = "cuda"
    c = torch.zeros((0, 4), dtype=torch.int32, device=dev)
    idx = ops.build_grid_index(c, 1, (4, 8, 8))
    tbl = ops.rulebook_subm(c, idx, 3)
    out = ops.spconv_fwd(torch.zeros((0, 16), device=dev), torch.zeros((3, 3, 3, 16, 16), device=dev), tbl, 0)
    assert out.shape == (0, 16)
    sc = ops.sparse_out_coords(c, 0, 1, (4, 8, 8), 3, 2, 1)
    assert sc.n == 0
    c1 = torch.tensor([[0, 3, 7, 7]], dtype=torch.int32, device=dev)               # corner voxel
    idx1 = ops.build_grid_index(c1, 1, (4, 8, 8))
    t1 = ops.rulebook_subm(c1, idx1, 3).cpu().numpy()
    assert t1[13, 0] == 0 and (np.delete(t1[:, 0], 13) == -1).all()
    sc1 = ops.sparse_out_coords(c1, 1, 1, (4, 8, 8), 3, 2, 1)
    oc, tb, _, _ = R.rulebook_sparse(c1.cpu().numpy(), (4, 8, 8), 3, 2, 1)
    np.testing.assert_array_equal(sc1.coors.cpu().numpy(), oc)


def test_full_size_rulebooks_bit_exact_and_properties():
    """One Waymo-sized scene: every rulebook of the backbone equals the oracle's bit for bit."""
    cloud = synth.lidar_scene(1000)
    v, c3, n = R.points_to_voxel(cloud, synth.WAYMO_VOXEL, synth.WAYMO_RANGE, 5, True, 150000)
    coors = np.concatenate([np.zeros((len(c3), 1), np.int32), c3], 1)
    shape = (41, 1504, 1504)
    c, tbl, pairs, idx, _ = gpu_rulebook(coors, 1, shape, "subm", 3)
    rt, rp = R.rulebook_subm(coors, shape, 3)
    t = tbl.cpu().numpy()
    np.testing.assert_array_equal(t, rt)
    assert pairs == rp
    assert np.array_equal(t[13], np.arange(len(coors)))                              # centre = identity
    for ks, st, pd in ((3, 2, 1), (3, 2, 1), (3, 2, (0, 1, 1)), ((3, 1, 1), (2, 1, 1), 0)):
        oc_ref, tbl_ref, shape_o, p_ref = R.rulebook_sparse(coors, shape, ks, st, pd)
        oc, tb, p, _, so = gpu_rulebook(coors, 1, shape, "sparse", ks, st, pd)
        np.testing.assert_array_equal(oc.cpu().numpy(), oc_ref)
        np.testing.assert_array_equal(tb.cpu().numpy(), tbl_ref)
        assert p == p_ref and tuple(so) == tuple(shape_o)
        lin = ((oc_ref[:, 1].astype(np.int64) * so[1]) + oc_ref[:, 2]) * so[2] + oc_ref[:, 3]
        assert (np.diff(lin) > 0).all()                                              # canonical ascending order
        coors, shape = oc_ref, tuple(int(x) for x in shape_o)


def test_dense_bev_matches_oracle():
    rng = np.random.default_rng(2)
    n, C, B, D, H, W = 500, 128, 2, 2, 20, 24
    lin = rng.choice(B * D * H * W, n, replace=False)
    coors = np.stack([lin // (D * H * W), (lin // (H * W)) % D, (lin // W) % H, lin % W], 1).astype(np.int32)
    feats = rng.normal(size=(n, C)).astype(np.float32)
    bev = ops.dense_bev(torch.from_numpy(feats).cuda(), torch.from_numpy(coors).cuda(), B, (D, H, W))
    np.testing.assert_array_equal(bev.cpu().numpy(), R.dense_bev(feats, coors, B, (D, H, W)))
    bev2 = ops.dense_bev_rowwise(torch.from_numpy(feats).cuda(), torch.from_numpy(coors).cuda(), B, (D, H, W))
    assert torch.equal(bev, bev2)                                # tiled (output-stationary) == row-stationary kernel
