"""GPU parity of the SECOND backbone SpMiddleFHD (backbone part; scn.py:187-289) through the
reference-shaped module API against the CPU oracle restatement (oracle/backbone.py::fhd_forward).  Bar: 1e-3 relative."""
import numpy as np
import pytest
import torch

from oracle import backbone as OB
from oracle import ref_ops as R
from sparse2dense_b200 import ops, registry, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("precision", [ops.PRECISION_FP32, ops.PRECISION_AUTO])
def test_spmiddlefhd_small_clouds_vs_oracle(precision):
    bb = registry.build_backbone(dict(type="SpMiddleFHD", num_input_features=5, ds_factor=8))
    state = synth.random_module_state(bb, 21)
    bb.load_state_dict({k: torch.as_tensor(v) for k, v in state.items()}, strict=False)
    bb = bb.cuda().eval()
    bb.set_precision(precision)
    clouds = [synth.small_scene(90), synth.small_scene(91)]
    grid = R.grid_size(synth.WAYMO_VOXEL, synth.WAYMO_RANGE)
    feats, coors = [], []
    for b, c in enumerate(clouds):
        v, co, n = R.points_to_voxel(c, synth.WAYMO_VOXEL, synth.WAYMO_RANGE, 5, True, 150000)
        feats.append(R.voxel_mean(v, n))
        coors.append(np.concatenate([np.full((len(co), 1), b, np.int32), co], 1))
    feats, coors = np.concatenate(feats), np.concatenate(coors)
    want, (f4, c4, s4) = OB.fhd_forward(state, feats, coors, 2, grid)
    with torch.no_grad():
        got, conv_4 = bb(torch.from_numpy(feats).cuda(), torch.from_numpy(coors).cuda(), 2, grid)
    assert tuple(got.shape) == tuple(want.shape) == (2, 128, 188, 188)
    assert np.array_equal(conv_4.indices.cpu().numpy(), c4) and tuple(conv_4.spatial_shape) == tuple(s4)
    err = lambda a, b: float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
    assert err(conv_4.features.cpu().numpy(), f4) < 1e-3
    assert err(got.cpu().numpy(), want) < 1e-3
    assert bb.bev_hw(grid) == (188, 188)
