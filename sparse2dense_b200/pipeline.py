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


class AssignLabel(object):
    """``AssignLabel`` pipeline step (det3d/datasets/pipelines/preprocess.py:479-653) for a whole batch on the device:
    same constructor contract (``cfg`` with out_size_factor / target_assigner.tasks / gaussian_overlap / max_objs /
    min_radius); ``__call__`` takes the per-scene annotations and returns the training keys of the example dict
    (``hm``, ``anno_box``, ``ind``, ``mask``, ``cat``: lists with one ``[B, ...]`` device tensor per task, and
    ``gt_boxes_and_cls``).  The host only regroups the (<= 500) objects of a scene class-major per task as the reference
    does (:505-537) and uploads 18 KB per scene; heat maps, indices and regression targets are one kernel launch per task
    (``s2d_assign_label``)."""

    def __init__(self, **kwargs):
        cfg = kwargs["cfg"]
        self.out_size_factor = _cfg(cfg, "out_size_factor")
        tasks = _cfg(_cfg(cfg, "target_assigner"), "tasks")
        self.num_classes = [int(_cfg(t, "num_class")) for t in tasks]
        self.gaussian_overlap = _cfg(cfg, "gaussian_overlap")
        self._max_objs = int(_cfg(cfg, "max_objs"))
        self._min_radius = int(_cfg(cfg, "min_radius"))

    def __call__(self, gt_boxes, gt_classes, grid_size, pc_range, voxel_size, device="cuda"):
        """gt_boxes: list (per scene) of ``[n,9]`` float32 (x,y,z,w,l,h,vx,vy,rot); gt_classes: list of ``[n]`` ints,
        1-based over all tasks; grid_size (x, y[, z]), pc_range, voxel_size as ``res["lidar"]["voxels"]`` holds them."""
        import ctypes
        from . import _lib
        B, M = len(gt_boxes), self._max_objs
        W, H = int(grid_size[0]) // self.out_size_factor, int(grid_size[1]) // self.out_size_factor
        lib = _lib.load()
        out = dict(hm=[], anno_box=[], ind=[], mask=[], cat=[])
        flag, per_task_bc = 0, []
        for ncls in self.num_classes:
            boxes = np.zeros((B, M, 9), np.float32)
            classes = np.zeros((B, M), np.int32)
            counts = np.zeros((B,), np.int32)
            for b in range(B):
                gb, gc = np.asarray(gt_boxes[b], np.float32).reshape(-1, 9), np.asarray(gt_classes[b]).reshape(-1)
                order = np.concatenate([np.where(gc == c + 1 + flag)[0] for c in range(ncls)]) if gc.size else np.zeros(0, int)
                n = min(order.size, M)
                assert order.size <= M, "more objects than max_objs"
                boxes[b, :n], classes[b, :n], counts[b] = gb[order[:n]], gc[order[:n]] - flag, n
            t = lambda a: torch.from_numpy(a).to(device)
            d_boxes, d_cls, d_cnt = t(boxes), t(classes), t(counts)
            hm = torch.empty((B, ncls, H, W), dtype=torch.float32, device=device)
            anno = torch.empty((B, M, 10), dtype=torch.float32, device=device)
            ind = torch.empty((B, M), dtype=torch.int64, device=device)
            mask = torch.empty((B, M), dtype=torch.uint8, device=device)
            cat = torch.empty((B, M), dtype=torch.int64, device=device)
            bc = torch.empty((B, M, 10), dtype=torch.float32, device=device)
            _lib.check(lib.s2d_assign_label(d_boxes.data_ptr(), d_cls.data_ptr(), d_cnt.data_ptr(), B, M, ncls, H, W,
                                            float(pc_range[0]), float(pc_range[1]), float(voxel_size[0]), float(voxel_size[1]),
                                            int(self.out_size_factor), float(self.gaussian_overlap), self._min_radius, flag,
                                            hm.data_ptr(), anno.data_ptr(), ind.data_ptr(), mask.data_ptr(), cat.data_ptr(),
                                            bc.data_ptr(), torch.cuda.current_stream().cuda_stream), "s2d_assign_label")
            for key, v in zip(("hm", "anno_box", "ind", "mask", "cat"), (hm, anno, ind, mask, cat)):
                out[key].append(v)
            per_task_bc.append((bc, counts))
            flag += ncls
        if len(per_task_bc) == 1:
            out["gt_boxes_and_cls"] = per_task_bc[0][0]
        else:                                   # flatten(gt_boxes) over tasks (preprocess.py:625-640): task-major rows
            bc = torch.zeros((B, M, 10), dtype=torch.float32, device=device)
            for b in range(B):
                off = 0
                for t_bc, counts in per_task_bc:
                    n = int(counts[b])
                    assert off + n <= M
                    bc[b, off:off + n] = t_bc[b, :n]
                    off += n
            out["gt_boxes_and_cls"] = bc
        return out
