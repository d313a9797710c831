"""CPU: host-side mirror of the reference interfaces (registry, module tree, state-dict names)."""
import numpy as np
import pytest
import torch

from oracle import backbone as OB
from sparse2dense_b200 import registry, spconv
from sparse2dense_b200.backbones import SpMiddleResNetFHD, SparseBasicBlock
from sparse2dense_b200.readers import VoxelFeatureExtractorV3


def test_registry_behaviour_matches_reference():
    r = registry.Registry("thing")

    @r.register_module
    class A:
        def __init__(self, x, y=2):
            self.x, self.y = x, y

    with pytest.raises(KeyError):
        r.register_module(A)                                  # duplicate (registry.py:37-40)
    with pytest.raises(TypeError):
        r.register_module(lambda: 0)                          # not a class
    obj = registry.build_from_cfg(dict(type="A", x=1), r, dict(y=5))
    assert (obj.x, obj.y) == (1, 5)
    with pytest.raises(KeyError):
        registry.build_from_cfg(dict(type="B"), r)
    with pytest.raises(AssertionError):
        registry.build_from_cfg(dict(x=1), r)
    assert registry.build_from_cfg(dict(type=A, x=3), r).x == 3


def test_backbone_builds_from_reference_config_dict():
    # configs/waymo/voxelnet/two_stage/waymo_centerpoint_voxelnet_two_stage_distill.py backbone/reader dicts
    bb = registry.build_backbone(dict(type="SpMiddleResNetFHD", num_input_features=5, ds_factor=8))
    rd = registry.build_reader(dict(type="VoxelFeatureExtractorV3", num_input_features=5))
    assert isinstance(bb, SpMiddleResNetFHD) and isinstance(rd, VoxelFeatureExtractorV3)
    assert sum(p.numel() for p in bb.parameters()) == 2695312


def test_state_dict_names_and_shapes_match_reference_layout():
    bb = SpMiddleResNetFHD(num_input_features=5)
    sd = bb.state_dict()
    ref = OB.random_state(0)                                  # hand-listed reference key set
    for k, v in ref.items():
        assert k in sd, k
        assert tuple(sd[k].shape) == v.shape, k
    extra = [k for k in sd if k not in ref and not k.endswith("num_batches_tracked")]
    assert extra == []
    assert sd["conv_input.0.weight"].shape == (3, 3, 3, 5, 16)       # spconv layout [kD,kH,kW,Cin,Cout]
    assert sd["extra_conv.0.weight"].shape == (3, 1, 1, 128, 128)
    assert "conv1.0.conv1.bias" in sd and "conv2.0.bias" not in sd   # scn.py:58 quirk vs bias=False
    bb.load_state_dict({k: torch.from_numpy(v) for k, v in ref.items()}, strict=False)


def test_module_tree_mirrors_scn_py():
    bb = SpMiddleResNetFHD(num_input_features=5)
    assert isinstance(bb.conv_input, spconv.SparseSequential) and isinstance(bb.conv_input[0], spconv.SubMConv3d)
    assert bb.conv_input[0].indice_key == "res0" and bb.conv1[0].conv1.indice_key == "res0"
    assert isinstance(bb.conv2[0], spconv.SparseConv3d) and bb.conv2[0].stride == (2, 2, 2)
    assert bb.conv4[0].padding == (0, 1, 1)
    assert bb.extra_conv[0].kernel_size == (3, 1, 1) and bb.extra_conv[0].stride == (2, 1, 1)
    assert isinstance(bb.conv3[3], SparseBasicBlock) and isinstance(bb.conv3[3], spconv.SparseModule)
    assert bb.conv1[0].bn1.eps == 1e-3 and bb.conv1[0].bn1.momentum == 0.01


def test_out_capacity_bound_is_an_upper_bound():
    from oracle import ref_ops as R
    from sparse2dense_b200 import ops
    rng = np.random.default_rng(0)
    lin = rng.choice(2 * 9 * 20 * 20, 900, replace=False)
    coors = np.stack([lin // 3600, (lin // 400) % 9, (lin // 20) % 20, lin % 20], 1).astype(np.int32)
    for ks, st, pd in ((3, 2, 1), (3, 2, (0, 1, 1)), ((3, 1, 1), (2, 1, 1), 0)):
        oc, _, oshape, _ = R.rulebook_sparse(coors, (9, 20, 20), ks, st, pd)
        assert len(oc) <= ops.out_capacity_bound(len(coors), 2, tuple(oshape), ks, st)
