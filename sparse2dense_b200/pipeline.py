"""GPU-side data boundary: the ``Voxelization`` pipeline step
(det3d/datasets/pipelines/preprocess.py:276-463) and the voxel part of ``collate_kitti``
(det3d/torchie/parallel/collate.py:106-108,137-144) for a whole batch on the device.

The reference voxelizes every scene on the CPU inside DataLoader workers (0.4 s per cloud and per voxel size; the
distillation pipeline does it five times per scene) and ships ~15 MB of padded voxels per scene over PCIe.  Here the
raw points go to the device once and each of the (up to five) voxelizations is one ``s2d_voxelize`` call for the
batch, bit-exact with the reference's serial loop."""
import numpy as np
import torch

from .voxel_generator import VoxelGenerator


def _cfg(cfg, name, default=None):
    if isinstance(cfg, dict):
        return cfg.get(name, default)
    return getattr(cfg, name, default)


class Voxelization(object):
    """Same constructor contract as the reference step (``cfg`` with range / voxel_size / max_points_in_voxel /
    max_voxel_num, optional ``distillation``); ``__call__`` works on batches of device point clouds."""

    def __init__(self, **kwargs):
        cfg = kwargs.get("cfg", None)
        distillation = kwargs.get("distillation", None)
        self.range = _cfg(cfg, "range")
        self.voxel_size = _cfg(cfg, "voxel_size")
        self.max_points_in_voxel = _cfg(cfg, "max_points_in_voxel")
        mv = _cfg(cfg, "max_voxel_num")
        self.max_voxel_num = [mv, mv] if isinstance(mv, int) else list(mv)
        self.distillation = bool(_cfg(cfg, "distillation", False) if distillation is None else distillation)
        if _cfg(cfg, "double_flip", False):
            raise NotImplementedError("double-flip test-time augmentation is not used by the Waymo configs")

        def gen(scale):
            return VoxelGenerator([x * scale for x in self.voxel_size], self.range, self.max_points_in_voxel,
                                  self.max_voxel_num[0])
        self.voxel_generator = gen(1)
        if self.distillation:                                    # preprocess.py:294-313
            self.voxel_generator_, self.voxel_generator_2, self.voxel_generator_4 = gen(1), gen(2), gen(4)

    @staticmethod
    def _cat(clouds, device):
        pts = [torch.as_tensor(np.ascontiguousarray(c, np.float32)) if not torch.is_tensor(c) else c.float() for c in clouds]
        offs = np.concatenate([[0], np.cumsum([int(p.shape[0]) for p in pts])]).astype(np.int64).tolist()
        return torch.cat([p.to(device, non_blocking=True) for p in pts]).contiguous(), offs

    def _voxels(self, generator, clouds, device, prefix, suffix, out, max_voxels):
        pts, offs = self._cat(clouds, device)
        saved = generator._max_voxels
        generator._max_voxels = max_voxels
        try:
            vb = generator.generate_batch(pts, offs, want_voxels=True)
        finally:
            generator._max_voxels = saved
        counts = np.diff(vb.offsets_host())
        out[f"{prefix}voxels{suffix}"] = vb.voxels
        out[f"{prefix}coordinates{suffix}"] = vb.coors                     # (b, z, y, x): collate.py:137-144
        out[f"{prefix}num_points{suffix}"] = vb.num_points
        out[f"{prefix}num_voxels{suffix}"] = torch.as_tensor(counts.astype(np.int64))

    def __call__(self, points, dense_points=None, reconstruction_points=None, mode="val", device="cuda"):
        """points (and, for distillation, dense_points / reconstruction_points): lists of per-scene ``[N,F]`` arrays or
        tensors -> the model's example dict (voxel keys only; targets come from ``AssignLabel``)."""
        max_voxels = self.max_voxel_num[0] if mode == "train" else self.max_voxel_num[1]
        ex = {}
        self._voxels(self.voxel_generator, points, device, "", "", ex, max_voxels)
        ex["shape"] = [np.asarray(self.voxel_generator.grid_size)] * len(points)   # the base grid for every variant (quirk)
        if self.distillation:
            assert dense_points is not None and reconstruction_points is not None
            self._voxels(self.voxel_generator, dense_points, device, "dense_", "", ex, max_voxels)
            self._voxels(self.voxel_generator_, reconstruction_points, device, "reconstruction_", "", ex, max_voxels)
            self._voxels(self.voxel_generator_2, reconstruction_points, device, "reconstruction_", "_2", ex, max_voxels)
            self._voxels(self.voxel_generator_4, reconstruction_points, device, "reconstruction_", "_4", ex, max_voxels)
        return ex
