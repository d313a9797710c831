"""GPU parity: s2d_voxelize (through the C ABI) vs the reference-generated goldens and the oracle."""
import glob
import os

import numpy as np
import pytest
import torch

from oracle import ref_ops as R
from sparse2dense_b200 import ops, synth

from conftest import GOLDEN

pytestmark = pytest.mark.gpu
VOX = sorted(glob.glob(os.path.join(GOLDEN, "voxelize_*.npz")))


def run_gpu(clouds, vs, rg, max_points, max_voxels, mean_channels=None):
    offs = np.concatenate([[0], np.cumsum([len(c) for c in clouds])]).astype(int).tolist()
    f = clouds[0].shape[1]
    cat = np.concatenate(clouds, 0) if offs[-1] else np.zeros((0, f), np.float32)
    pts = torch.from_numpy(cat).cuda()
    vb = ops.voxelize(pts, offs, vs, rg, max_points, max_voxels, want_voxels=True, mean_channels=mean_channels)
    torch.cuda.synchronize()
    return vb


def compare_scene(vb, b, ref):
    v, c, n = ref
    o = vb.offsets_host()
    lo, hi = o[b], o[b + 1]
    assert hi - lo == len(c)
    got_c = vb.coors[lo:hi].cpu().numpy()
    assert (got_c[:, 0] == b).all()
    np.testing.assert_array_equal(got_c[:, 1:], c)
    np.testing.assert_array_equal(vb.num_points[lo:hi].cpu().numpy(), n)
    np.testing.assert_array_equal(vb.voxels[lo:hi].cpu().numpy(), v)


@pytest.mark.parametrize("path", VOX, ids=[os.path.basename(p)[9:-4] for p in VOX])
def test_voxelize_bit_exact_vs_reference_goldens(path):
    g = np.load(path)
    vb = run_gpu([g["points"]], g["voxel_size"], g["coors_range"], int(g["max_points"]), int(g["max_voxels"]))
    compare_scene(vb, 0, (g["voxels"], g["coors"], g["num_points"]))


def test_voxelize_batch_ragged_with_empty_scene_and_cap():
    a = synth.small_scene(21)
    b = np.zeros((0, 5), np.float32)
    c = synth.small_scene(22)[:7000]
    d = synth.small_scene(23)
    mv = 6000                                                  # scene a and d overflow the cap
    vb = run_gpu([a, b, c, d], synth.WAYMO_VOXEL, synth.WAYMO_RANGE, 5, mv, mean_channels=5)
    for i, cloud in enumerate([a, b, c, d]):
        compare_scene(vb, i, R.points_to_voxel(cloud, synth.WAYMO_VOXEL, synth.WAYMO_RANGE, 5, True, mv))
    ref_mean = R.voxel_mean(vb.voxels.cpu().numpy(), vb.num_points.cpu().numpy())
    np.testing.assert_allclose(vb.mean.cpu().numpy(), ref_mean, rtol=1e-6, atol=1e-6)
    np.testing.assert_allclose(ops.voxel_mean(vb.voxels, vb.num_points).cpu().numpy(), ref_mean, rtol=1e-6, atol=1e-6)


def test_voxelize_full_size_batch4_bit_exact():
    """BASELINE configs[1] size: 4 clouds of ~180 k in-range points on the 1504x1504x40 grid."""
    clouds = synth.lidar_batch(1, 4)
    vb = run_gpu(clouds, synth.WAYMO_VOXEL, synth.WAYMO_RANGE, synth.WAYMO_MAX_POINTS, synth.WAYMO_MAX_VOXELS)
    for i, cloud in enumerate(clouds):
        compare_scene(vb, i, R.points_to_voxel(cloud, synth.WAYMO_VOXEL, synth.WAYMO_RANGE, 5, True, 150000))


def test_voxelize_properties_shuffled_points():
    """Size-independent properties: unique coordinates, counts conserved, same voxel SET under shuffling."""
    cloud = synth.lidar_scene(77)
    rng = np.random.default_rng(0)
    shuf = cloud[rng.permutation(len(cloud))]
    va = run_gpu([cloud], synth.WAYMO_VOXEL, synth.WAYMO_RANGE, 5, 150000)
    vs = run_gpu([shuf], synth.WAYMO_VOXEL, synth.WAYMO_RANGE, 5, 150000)
    ca, cs = va.coors.cpu().numpy(), vs.coors.cpu().numpy()
    key = lambda c: (c[:, 1].astype(np.int64) * 1504 + c[:, 2]) * 1504 + c[:, 3]
    ka, ks = key(ca), key(cs)
    assert len(np.unique(ka)) == len(ka)
    assert np.array_equal(np.sort(ka), np.sort(ks))
    n_in = int(synth.in_range_mask(cloud).sum())
    na = va.num_points.cpu().numpy()
    assert na.min() >= 1 and na.max() <= 5 and na.sum() <= n_in


def test_voxelize_rejects_bad_arguments():
    pts = torch.zeros((10, 5), device="cuda")
    with pytest.raises(RuntimeError):
        ops.voxelize(pts, [0, 5], synth.WAYMO_VOXEL, synth.WAYMO_RANGE, 5, 100)      # offsets do not end at N


def test_device_side_voxelization_step_all_five_variants_bit_exact():
    """pipeline.Voxelization (the distillation pipeline's five voxelizations, batched on the device) vs the oracle."""
    from sparse2dense_b200.pipeline import Voxelization
    cfg = dict(range=list(synth.WAYMO_RANGE), voxel_size=list(synth.WAYMO_VOXEL), max_points_in_voxel=5,
               max_voxel_num=150000, distillation=True)
    step = Voxelization(cfg=cfg)
    rng = np.random.default_rng(0)
    pts = [synth.small_scene(11), synth.small_scene(12)]
    dense = [np.concatenate([p, p[rng.integers(0, len(p), 3000)] + rng.normal(0, 0.05, (3000, 5)).astype(np.float32)]) for p in pts]
    recon = [d[rng.permutation(len(d))[: len(d) // 2]] for d in dense]
    ex = step(pts, dense_points=dense, reconstruction_points=recon)
    vs = np.asarray(synth.WAYMO_VOXEL, np.float32)
    for prefix, suffix, clouds, scale in [("", "", pts, 1), ("dense_", "", dense, 1), ("reconstruction_", "", recon, 1),
                                          ("reconstruction_", "_2", recon, 2), ("reconstruction_", "_4", recon, 4)]:
        v = ex[f"{prefix}voxels{suffix}"].cpu().numpy()
        c = ex[f"{prefix}coordinates{suffix}"].cpu().numpy()
        n = ex[f"{prefix}num_points{suffix}"].cpu().numpy()
        counts = ex[f"{prefix}num_voxels{suffix}"].numpy()
        off = 0
        for b, cloud in enumerate(clouds):
            rv, rc, rn = R.points_to_voxel(cloud, [x * scale for x in vs], synth.WAYMO_RANGE, 5, True, 150000)
            k = len(rv)
            assert counts[b] == k
            assert np.array_equal(v[off:off + k], rv) and np.array_equal(n[off:off + k], rn)
            assert np.array_equal(c[off:off + k, 1:], rc) and (c[off:off + k, 0] == b).all()
            off += k
        assert off == len(v)
    assert tuple(ex["shape"][0]) == (1504, 1504, 40)
