"""GPU parity: reader + SpMiddleResNetFHD through the reference-shaped module API vs the oracle."""
import numpy as np
import pytest
import torch

from oracle import backbone as OB
from oracle import ref_ops as R
from sparse2dense_b200 import ops, registry, synth

pytestmark = pytest.mark.gpu
RTOL = 1e-3          # north_star: 1e-3 relative fp32


def make_backbone(state):
    bb = registry.build_backbone(dict(type="SpMiddleResNetFHD", num_input_features=5, ds_factor=8))
    missing = bb.load_state_dict({k: torch.from_numpy(v) for k, v in state.items()}, strict=False)
    assert not missing.unexpected_keys and all(k.endswith("num_batches_tracked") for k in missing.missing_keys)
    return bb.cuda().eval()


def rel_err(a, b):
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


@pytest.mark.parametrize("batch", [1, 2])
def test_backbone_small_clouds_vs_oracle(batch):
    state = OB.random_state(0)
    bb = make_backbone(state)
    clouds = [synth.small_scene(31 + i) for i in range(batch)]
    offs = np.concatenate([[0], np.cumsum([len(c) for c in clouds])]).tolist()
    vb = ops.voxelize(torch.from_numpy(np.concatenate(clouds)).cuda(), offs, synth.WAYMO_VOXEL, synth.WAYMO_RANGE, 5,
                      150000, want_voxels=True)
    reader = registry.build_reader(dict(type="VoxelFeatureExtractorV3", num_input_features=5))
    feats = reader(vb.voxels, vb.num_points)
    with torch.no_grad():
        bev, multi = bb(feats, vb.coors, batch, [1504, 1504, 40])
    torch.cuda.synchronize()
    ref_bev, ref_multi = OB.backbone_forward(state, feats.cpu().numpy(), vb.coors.cpu().numpy(), batch,
                                             (1504, 1504, 40), wide=True)
    assert bev.shape == (batch, 256, 188, 188)
    for name in ("conv1", "conv2", "conv3", "conv4"):
        x, c, s = ref_multi[name]
        np.testing.assert_array_equal(multi[name].indices.cpu().numpy(), c)            # indices bit-exact
        assert tuple(multi[name].spatial_shape) == tuple(s)
        assert rel_err(multi[name].features.cpu().numpy(), x) < RTOL, name
    assert rel_err(bev.cpu().numpy(), ref_bev) < RTOL
    assert np.array_equal(bev.cpu().numpy() != 0, ref_bev != 0) or rel_err(bev.cpu().numpy(), ref_bev) < 1e-5


def test_backbone_full_size_scene_vs_oracle():
    """One 180 k-point scene on the 1504x1504x40 grid (oracle takes ~20 s)."""
    state = OB.random_state(1)
    bb = make_backbone(state)
    cloud = synth.lidar_scene(1000)
    vb = ops.voxelize(torch.from_numpy(cloud).cuda(), [0, len(cloud)], synth.WAYMO_VOXEL, synth.WAYMO_RANGE, 5, 150000,
                      want_voxels=False, mean_channels=5)
    with torch.no_grad():
        bev, multi = bb(vb.mean, vb.coors, 1, [1504, 1504, 40])
    ref_bev, ref_multi = OB.backbone_forward(state, vb.mean.cpu().numpy(), vb.coors.cpu().numpy(), 1,
                                             (1504, 1504, 40), wide=False)
    for name in ("conv1", "conv2", "conv3", "conv4"):
        np.testing.assert_array_equal(multi[name].indices.cpu().numpy(), ref_multi[name][1])
        assert rel_err(multi[name].features.cpu().numpy(), ref_multi[name][0]) < RTOL, name
    assert rel_err(bev.cpu().numpy(), ref_bev) < RTOL


def test_backbone_batch4_properties():
    """configs[1] size (batch 4): scenes are independent -> batched result == per-scene results."""
    state = OB.random_state(2)
    bb = make_backbone(state)
    clouds = synth.lidar_batch(1, 4)
    offs = np.concatenate([[0], np.cumsum([len(c) for c in clouds])]).tolist()
    vb = ops.voxelize(torch.from_numpy(np.concatenate(clouds)).cuda(), offs, synth.WAYMO_VOXEL, synth.WAYMO_RANGE, 5,
                      150000, want_voxels=False, mean_channels=5)
    with torch.no_grad():
        bev4, _ = bb(vb.mean, vb.coors, 4, [1504, 1504, 40])
        v1 = ops.voxelize(torch.from_numpy(clouds[2]).cuda(), [0, len(clouds[2])], synth.WAYMO_VOXEL,
                          synth.WAYMO_RANGE, 5, 150000, want_voxels=False, mean_channels=5)
        bev1, _ = bb(v1.mean, v1.coors, 1, [1504, 1504, 40])
    assert bev4.shape == (4, 256, 188, 188)
    assert torch.equal(bev4[2], bev1[0])                       # deterministic, order-independent kernels
    assert torch.isfinite(bev4).all()
