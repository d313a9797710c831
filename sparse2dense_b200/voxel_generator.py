"""Drop-in for the reference's voxelizer front-ends, running on the GPU.

``points_to_voxel`` keeps the signature and return value of
det3d/ops/point_cloud/point_cloud_ops.py:112-184 (numpy in / numpy out), ``VoxelGenerator`` those
of det3d/core/input/voxel_generator.py:5-46.  ``voxelize_batch`` is the device-resident form the
hot path uses (no host round trip): a list of clouds -> collate_kitti-style batch tensors
(det3d/torchie/parallel/collate.py:106-108,137-144).
"""
import numpy as np
import torch

from . import ops


def points_to_voxel(points, voxel_size, coors_range, max_points=35, reverse_index=True, max_voxels=20000):
    if not reverse_index:
        raise NotImplementedError("the reference hot path only uses reverse_index=True (voxel_generator.py:23-30)")
    pts = torch.as_tensor(np.ascontiguousarray(points, dtype=np.float32)).cuda()
    vb = ops.voxelize(pts, [0, pts.shape[0]], voxel_size, coors_range, max_points, max_voxels, want_voxels=True)
    return (vb.voxels.cpu().numpy(), vb.coors[:, 1:].contiguous().cpu().numpy(), vb.num_points.cpu().numpy())


class VoxelGenerator:
    def __init__(self, voxel_size, point_cloud_range, max_num_points, max_voxels=20000):
        point_cloud_range = np.array(point_cloud_range, dtype=np.float32)
        voxel_size = np.array(voxel_size, dtype=np.float32)
        grid_size = (point_cloud_range[3:] - point_cloud_range[:3]) / voxel_size
        grid_size = np.round(grid_size).astype(np.int64)
        self._voxel_size = voxel_size
        self._point_cloud_range = point_cloud_range
        self._max_num_points = max_num_points
        self._max_voxels = max_voxels
        self._grid_size = grid_size

    def generate(self, points, max_voxels=-1):
        if max_voxels == -1:
            max_voxels = self._max_voxels
        return points_to_voxel(points, self._voxel_size, self._point_cloud_range, self._max_num_points, True,
                               max_voxels)

    def generate_batch(self, points, scene_offsets, want_voxels=True, mean_channels=None):
        """Device-resident batch form: points cuda f32 [N,F] + host offsets -> ops.VoxelBatch."""
        return ops.voxelize(points, scene_offsets, self._voxel_size, self._point_cloud_range, self._max_num_points,
                            self._max_voxels, want_voxels=want_voxels, mean_channels=mean_channels)

    @property
    def voxel_size(self):
        return self._voxel_size

    @property
    def max_num_points_per_voxel(self):
        return self._max_num_points

    @property
    def point_cloud_range(self):
        return self._point_cloud_range

    @property
    def grid_size(self):
        return self._grid_size
