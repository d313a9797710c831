"""The hot path as one callable: points -> voxelize (+ fused reader mean) -> SpMiddleResNetFHD -> BEV.

This is what ``bench.py`` and ``__graft_entry__.smoke()`` drive.  It is assembled from the same
registry entries a reference config names (``VoxelFeatureExtractorV3``, ``SpMiddleResNetFHD``;
configs/waymo/voxelnet/two_stage/waymo_centerpoint_voxelnet_two_stage_distill.py:52-56,170-176).
"""
import numpy as np
import torch

from . import ops, registry, synth
from .voxel_generator import VoxelGenerator


class VoxelBackbonePath:
    def __init__(self, state=None, precision=ops.PRECISION_FP32, device="cuda", num_input_features=5):
        if not torch.cuda.is_available():
            raise RuntimeError("VoxelBackbonePath needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device(device)
        self.generator = VoxelGenerator(synth.WAYMO_VOXEL, synth.WAYMO_RANGE, synth.WAYMO_MAX_POINTS,
                                        synth.WAYMO_MAX_VOXELS)
        self.reader = registry.build_reader(dict(type="VoxelFeatureExtractorV3",
                                                 num_input_features=num_input_features))
        self.backbone = registry.build_backbone(dict(type="SpMiddleResNetFHD",
                                                     num_input_features=num_input_features, ds_factor=8))
        if state is not None:
            self.backbone.load_state_dict({k: torch.as_tensor(v) for k, v in state.items()}, strict=False)
        self.backbone.to(self.device).eval()
        self.backbone.set_precision(precision)
        self.num_input_features = num_input_features
        self.grid = [int(v) for v in self.generator.grid_size]          # (x, y, z) = example["shape"][0]

    @torch.no_grad()
    def forward_points(self, points, scene_offsets):
        """points: cuda f32 [N,5] (scenes concatenated), scene_offsets: host ints [B+1] -> BEV [B,256,188,188]."""
        batch = len(scene_offsets) - 1
        vb = self.generator.generate_batch(points, scene_offsets, want_voxels=False,
                                           mean_channels=self.num_input_features)
        n = vb.n                                                          # host sync #1 (voxel count)
        bev, _ = self.backbone(vb.mean_buffer[:n], vb.coors_buffer[:n], batch, self.grid)
        return bev

    @torch.no_grad()
    def forward_host(self, points_host, scene_offsets, out_host=None):
        """End-to-end form: pinned host points -> device -> hot path -> pinned host BEV."""
        pts = points_host.to(self.device, non_blocking=True)
        bev = self.forward_points(pts, scene_offsets)
        if out_host is None:
            out_host = torch.empty(bev.shape, dtype=bev.dtype, pin_memory=True)
        out_host.copy_(bev, non_blocking=True)
        return out_host


def concat_clouds(clouds, pin=True):
    offs = np.concatenate([[0], np.cumsum([len(c) for c in clouds])]).astype(np.int64).tolist()
    cat = torch.from_numpy(np.ascontiguousarray(np.concatenate(clouds, 0), dtype=np.float32))
    if pin and torch.cuda.is_available():
        cat = cat.pin_memory()
    return cat, offs
