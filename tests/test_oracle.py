"""CPU: the oracle against the committed golden vectors (reference-generated for the voxelizer)."""
import glob
import os

import numpy as np
import pytest

from oracle import backbone as OB
from oracle import ref_ops as R
from sparse2dense_b200 import synth

from conftest import GOLDEN

VOX = sorted(glob.glob(os.path.join(GOLDEN, "voxelize_*.npz")))
SPC = sorted(glob.glob(os.path.join(GOLDEN, "spconv_*.npz")))


def test_goldens_present():
    assert len(VOX) >= 7 and len(SPC) >= 5


@pytest.mark.parametrize("path", VOX, ids=[os.path.basename(p)[9:-4] for p in VOX])
def test_voxelizer_matches_reference_numba(path):
    """Bit-exact vs the outputs of det3d/ops/point_cloud/point_cloud_ops.py:112-184 (pins the oracle)."""
    g = np.load(path)
    v, c, n = R.points_to_voxel(g["points"], g["voxel_size"], g["coors_range"], int(g["max_points"]), True,
                                int(g["max_voxels"]))
    assert c.dtype == np.int32 and n.dtype == np.int32
    np.testing.assert_array_equal(c, g["coors"])
    np.testing.assert_array_equal(n, g["num_points"])
    np.testing.assert_array_equal(v, g["voxels"])


def test_voxelizer_cap_semantics():
    """max_voxels: later voxels dropped, existing voxels keep filling (point_cloud_ops.py:44-54)."""
    pts = np.array([[0.05, 0.05, 0.05, 1, 1], [0.35, 0.05, 0.05, 2, 2], [0.06, 0.05, 0.05, 3, 3]], np.float32)
    v, c, n = R.points_to_voxel(pts, (0.1, 0.1, 0.1), (0, 0, 0, 1, 1, 1), 5, True, 1)
    assert len(c) == 1 and n[0] == 2
    np.testing.assert_array_equal(c[0], [0, 0, 0])
    np.testing.assert_array_equal(v[0, :2, 3], [1, 3])


@pytest.mark.parametrize("path", SPC, ids=[os.path.basename(p)[7:-4] for p in SPC])
def test_sparse_conv_matches_dense_conv3d(path):
    """R2 (rulebook + gather-GEMM) vs the committed R1 outputs (dense F.conv3d formulation)."""
    g = np.load(path)
    if str(g["kind"]) == "subm":
        tbl, pairs = R.rulebook_subm(g["coors"], g["shape"], g["ksize"])
        oc = g["coors"]
    else:
        oc, tbl, oshape, pairs = R.rulebook_sparse(g["coors"], g["shape"], g["ksize"], g["stride"], g["pad"])
        np.testing.assert_array_equal(oshape, g["out_shape"])
    np.testing.assert_array_equal(oc, g["out_coors"])
    assert pairs == int((tbl >= 0).sum())
    out = R.spconv_fwd(g["feats"], g["weight"], tbl)
    np.testing.assert_allclose(out, g["out"], rtol=1e-5, atol=1e-5)
    out64 = R.spconv_fwd(g["feats"], g["weight"], tbl, wide=True)
    np.testing.assert_allclose(out64, g["out"], rtol=1e-6, atol=1e-6)


def test_subm_rulebook_is_symmetric():
    rng = np.random.default_rng(0)
    lin = rng.choice(2 * 6 * 10 * 10, 200, replace=False)
    coors = np.stack([lin // 600, (lin // 100) % 6, (lin // 10) % 10, lin % 10], 1).astype(np.int32)
    tbl, _ = R.rulebook_subm(coors, (6, 10, 10), 3)
    assert np.array_equal(tbl[13], np.arange(200))          # centre offset maps every row to itself
    for k in range(27):                                    # j = nbr_k(i)  <=>  i = nbr_{26-k}(j)
        i = np.nonzero(tbl[k] >= 0)[0]
        assert np.array_equal(tbl[26 - k][tbl[k][i]], i)


def test_sparse_conv_linearity_and_empty():
    rng = np.random.default_rng(1)
    coors = np.array([[0, 1, 2, 3], [0, 1, 2, 4], [1, 0, 0, 0]], np.int32)
    tbl, _ = R.rulebook_subm(coors, (4, 6, 6), 3)
    w = rng.normal(size=(3, 3, 3, 4, 8)).astype(np.float32)
    a, b = rng.normal(size=(2, 3, 4)).astype(np.float32)
    np.testing.assert_allclose(R.spconv_fwd(a + b, w, tbl), R.spconv_fwd(a, w, tbl) + R.spconv_fwd(b, w, tbl),
                               rtol=1e-5, atol=1e-5)
    oc, tbl0, _, pairs = R.rulebook_sparse(np.zeros((0, 4), np.int32), (4, 6, 6), 3, 2, 1)
    assert len(oc) == 0 and pairs == 0 and tbl0.shape == (27, 0)


def test_backbone_oracle_small_grid_against_torch():
    """Whole SpMiddleResNetFHD restatement vs a dense torch conv3d pipeline on a tiny grid."""
    import torch
    import torch.nn.functional as F
    rng = np.random.default_rng(5)
    state = OB.random_state(3)
    grid = (16, 16, 24)                                  # x, y, z  -> sparse shape (25, 16, 16)
    shape = (25, 16, 16)
    B, n = 2, 600
    lin = rng.choice(B * 24 * 256, n, replace=False)
    coors = np.stack([lin // (24 * 256), (lin // 256) % 24, (lin // 16) % 16, lin % 16], 1).astype(np.int32)
    feats = rng.normal(size=(n, 5)).astype(np.float32)
    bev, multi = OB.backbone_forward(state, feats, coors, B, grid)

    def tw(name):
        return torch.from_numpy(state[name]).double().permute(4, 3, 0, 1, 2).contiguous()

    def bn(x, prefix, mask, bias=None):
        s, h = OB._bn(state, prefix, bias)
        y = x * torch.from_numpy(s).double().view(1, -1, 1, 1, 1) + torch.from_numpy(h).double().view(1, -1, 1, 1, 1)
        return y, mask

    x = torch.zeros(B, 5, *shape, dtype=torch.float64)
    m = torch.zeros(B, 1, *shape, dtype=torch.float64)
    c = torch.from_numpy(coors).long()
    x[c[:, 0], :, c[:, 1], c[:, 2], c[:, 3]] = torch.from_numpy(feats).double()
    m[c[:, 0], 0, c[:, 1], c[:, 2], c[:, 3]] = 1

    def subm(x, m, wname, bnname, bias=None, res=None, relu=True):
        y = F.conv3d(x, tw(wname), padding=1)
        y, _ = bn(y, bnname, m, bias)
        if res is not None:
            y = y + res
        if relu:
            y = torch.relu(y)
        return y * m

    def block(x, m, p):
        o = subm(x, m, p + ".conv1.weight", p + ".bn1", state[p + ".conv1.bias"])
        return subm(o, m, p + ".conv2.weight", p + ".bn2", state[p + ".conv2.bias"], res=x)

    def down(x, m, p, ks, st, pd):
        y = F.conv3d(x, tw(p + ".0.weight"), stride=st, padding=pd)
        m2 = (F.conv3d(m, torch.ones(1, 1, *ks, dtype=torch.float64), stride=st, padding=pd) > 0).double()
        y, _ = bn(y, p + ".1", m2)
        return torch.relu(y) * m2, m2

    x = subm(x, m, "conv_input.0.weight", "conv_input.1")
    x = block(block(x, m, "conv1.0"), m, "conv1.1")
    for p, pd in (("conv2", 1), ("conv3", 1), ("conv4", (0, 1, 1))):
        x, m = down(x, m, p, (3, 3, 3), 2, pd)
        x = block(block(x, m, p + ".3"), m, p + ".4")
    x, m = down(x, m, "extra_conv", (3, 1, 1), (2, 1, 1), 0)
    ref = x.reshape(B, -1, x.shape[3], x.shape[4]).float().numpy()
    assert ref.shape == bev.shape
    scale = np.abs(ref).max()
    assert np.abs(bev - ref).max() <= 1e-4 * scale


def test_synth_is_deterministic_and_waymo_sized():
    a, b = synth.lidar_scene(1000), synth.lidar_scene(1000)
    assert np.array_equal(a, b) and a.dtype == np.float32 and a.shape[1] == 5
    n_in = int(synth.in_range_mask(a).sum())
    assert 0.95 * 180000 <= n_in <= 1.05 * 180000


def test_bench_cpu_arm_states_match_the_oracle_full_forward_keys():
    """bench.py --impl reference builds the same seeded weights as the GPU arm WITHOUT a CUDA device; one small scene goes
    through the CPU restatement of the whole two-stage forward (configs[2]) and of the pillar student (configs[3])."""
    import bench
    from oracle import full_forward as FF
    from sparse2dense_b200 import synth
    cloud = synth.small_scene(5)
    for name, fn in (("full", FF.scene_forward), ("pillar", FF.pillar_scene_forward)):
        states = bench.cpu_states(name)
        t = {}
        out = fn(states, cloud, timings=t)
        assert out["boxes"].shape[1] == 7 and len(out["scores"]) == len(out["labels"]) == len(out["boxes"]) <= 500
        assert set(t) >= {"neck_head", "decode_nms", "second_stage"}
