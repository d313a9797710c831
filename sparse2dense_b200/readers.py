"""Readers (det3d/models/readers/voxel_encoder.py)."""
from torch import nn

from . import ops
from .registry import READERS


@READERS.register_module
class VoxelFeatureExtractorV3(nn.Module):
    """Mean of the points of each voxel (voxel_encoder.py:8-24), one CUDA kernel."""

    def __init__(self, num_input_features=4, norm_cfg=None, name="VoxelFeatureExtractorV3"):
        super(VoxelFeatureExtractorV3, self).__init__()
        self.name = name
        self.num_input_features = num_input_features

    def forward(self, features, num_voxels, coors=None):
        assert self.num_input_features == features.shape[-1]
        return ops.voxel_mean(features, num_voxels, self.num_input_features)
