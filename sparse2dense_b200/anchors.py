"""Anchors and box coder of the SECOND-style anchor head.

* ``create_anchors_3d_range`` / ``AnchorGeneratorRange``: det3d/core/bbox/box_np_ops.py:857-929,
  det3d/core/anchor/anchor_generator.py:64-123 -- host-side numpy, once per feature-map size;
* ``task_anchors``: the per-task concatenation of ``TargetAssigner.generate_anchors``
  (det3d/core/anchor/target_assigner.py:139-158) as ``AssignTarget`` puts it into ``example["anchors"]``
  (det3d/datasets/pipelines/preprocess.py:697-777);
* ``GroundBox3dCoderTorch``: det3d/core/bbox/box_coders.py:31-60,106-115 over ``second_box_encode / second_box_decode``
  (det3d/core/bbox/box_torch_ops.py:30-160), 7- and 9-dimensional boxes;
* ``build_box_coder`` / ``build_anchor_generator``: det3d/builder.py:65-100,296-338 (re-exported as ``det3d.builder``).
"""
import numpy as np
import torch


def create_anchors_3d_range(feature_size, anchor_range, sizes=(1.6, 3.9, 1.56), rotations=(0, np.pi / 2), velocities=None,
                            dtype=np.float32):
    """feature_size [D, H, W] (z, y, x) -> anchors [D, H, W, num_sizes, num_rots, 7 (9 with velocities)]."""
    anchor_range = np.array(anchor_range, dtype)
    stride = (anchor_range[3] - anchor_range[0]) / feature_size[2]
    z_centers = np.linspace(anchor_range[2], anchor_range[5], feature_size[0], dtype=dtype)
    y_centers = np.linspace(anchor_range[1], anchor_range[4], feature_size[1], endpoint=False, dtype=dtype) + stride / 2
    x_centers = np.linspace(anchor_range[0], anchor_range[3], feature_size[2], endpoint=False, dtype=dtype) + stride / 2
    rotations = np.array(rotations, dtype=dtype)
    sizes = np.reshape(np.array(sizes, dtype=dtype), [-1, 3])
    combines = sizes
    if velocities is not None:
        velocities = np.array(velocities, dtype=dtype).reshape([-1, 2])
        combines = np.hstack([sizes, velocities]).reshape([-1, 5])
    rets = list(np.meshgrid(x_centers, y_centers, z_centers, rotations, indexing="ij"))     # numpy 2 returns a tuple
    tile_shape = [1] * 5
    tile_shape[-2] = int(sizes.shape[0])
    for i in range(len(rets)):
        rets[i] = np.tile(rets[i][..., np.newaxis, :], tile_shape)[..., np.newaxis]
    combines = np.reshape(combines, [1, 1, 1, -1, 1, combines.shape[-1]])
    tile_size_shape = list(rets[0].shape)
    tile_size_shape[3] = 1
    rets.insert(3, np.tile(combines, tile_size_shape))
    return np.transpose(np.concatenate(rets, axis=-1), [2, 1, 0, 3, 4, 5])


class AnchorGeneratorRange:
    def __init__(self, anchor_ranges, sizes=(1.6, 3.9, 1.56), rotations=(0, np.pi / 2), velocities=None, class_name=None,
                 match_threshold=-1, unmatch_threshold=-1, dtype=np.float32):
        self._sizes, self._anchor_ranges, self._rotations, self._velocities = sizes, anchor_ranges, rotations, velocities
        self._dtype, self._class_name = dtype, class_name
        self._match_threshold, self._unmatch_threshold = match_threshold, unmatch_threshold

    class_name = property(lambda self: self._class_name)
    match_threshold = property(lambda self: self._match_threshold)
    unmatch_threshold = property(lambda self: self._unmatch_threshold)

    @property
    def num_anchors_per_localization(self):
        return len(self._rotations) * np.array(self._sizes).reshape([-1, 3]).shape[0]

    def generate(self, feature_map_size):
        self._anchors = create_anchors_3d_range(feature_map_size, self._anchor_ranges, self._sizes, self._rotations,
                                                self._velocities, self._dtype)
        return self._anchors


def _get(cfg, name, default=None):
    return cfg.get(name, default) if hasattr(cfg, "get") else getattr(cfg, name, default)


def build_anchor_generator(anchor_config):
    """det3d/builder.py:296-338 (``anchor_generator_range`` is the only kind the Waymo configs use)."""
    if _get(anchor_config, "type") != "anchor_generator_range":
        raise ValueError(f"unsupported anchor generator type {_get(anchor_config, 'type')!r}")
    return AnchorGeneratorRange(sizes=_get(anchor_config, "sizes"), anchor_ranges=_get(anchor_config, "anchor_ranges"),
                                rotations=_get(anchor_config, "rotations"), velocities=_get(anchor_config, "velocities"),
                                match_threshold=_get(anchor_config, "matched_threshold"),
                                unmatch_threshold=_get(anchor_config, "unmatched_threshold"),
                                class_name=_get(anchor_config, "class_name"))


def task_anchors(target_assigner_cfg, feature_map_size):
    """-> list (per task) of float32 [H*W*A_task, ndim]: the generators of a task's classes concatenated on the
    anchors-per-location axis (target_assigner.py:139-158) and flattened (preprocess.py:707-709).
    feature_map_size = [W, H, 1][::-1] = [1, H, W] as ``AssignTarget`` builds it (preprocess.py:698-699)."""
    gens = [build_anchor_generator(a) for a in _get(target_assigner_cfg, "anchor_generators")]
    out = []
    for task in _get(target_assigner_cfg, "tasks"):
        names = list(_get(task, "class_names"))
        parts = []
        for g in gens:
            if g.class_name in names:
                a = g.generate(feature_map_size)
                parts.append(a.reshape([*a.shape[:3], -1, a.shape[-1]]))
        anchors = np.concatenate(parts, axis=-2)
        out.append(np.ascontiguousarray(anchors.reshape([-1, anchors.shape[-1]]), np.float32))
    return out


class GroundBox3dCoderTorch:
    """box_coders.py:31-60,106-115: residual coding against the anchor diagonal / height, log sizes, additive yaw."""

    def __init__(self, linear_dim=False, vec_encode=False, n_dim=7, norm_velo=False):
        self.linear_dim, self.vec_encode, self.norm_velo, self.n_dim = linear_dim, vec_encode, norm_velo, n_dim

    @property
    def code_size(self):
        return self.n_dim + 1 if self.vec_encode else self.n_dim

    def decode_torch(self, box_encodings, anchors):
        """box_torch_ops.py:87-160 (``second_box_decode``); tensors [..., 7 | 9] (+1 code with the angle vector)."""
        nd = anchors.shape[-1]
        a = torch.split(anchors, 1, dim=-1)
        t = torch.split(box_encodings, 1, dim=-1)
        xa, ya, za, wa, la, ha = a[:6]
        ra = a[-1]
        xt, yt, zt, wt, lt, ht = t[:6]
        diagonal = torch.sqrt(la ** 2 + wa ** 2)
        xg, yg, zg = xt * diagonal + xa, yt * diagonal + ya, zt * ha + za
        if self.linear_dim:
            lg, wg, hg = (lt + 1) * la, (wt + 1) * wa, (ht + 1) * ha
        else:
            lg, wg, hg = torch.exp(lt) * la, torch.exp(wt) * wa, torch.exp(ht) * ha
        ret = [xg, yg, zg, wg, lg, hg]
        if nd > 7:
            vxa, vya, vxt, vyt = a[6], a[7], t[6], t[7]
            if self.norm_velo:
                ret.extend([vxt * diagonal + vxa, vyt * diagonal + vya])
            else:
                ret.extend([vxt + vxa, vyt + vya])
        if self.vec_encode:
            rtx, rty = t[-2], t[-1]
            ret.append(torch.atan2(rty + torch.sin(ra), rtx + torch.cos(ra)))
        else:
            ret.append(t[-1] + ra)
        return torch.cat(ret, dim=-1)

    def encode_torch(self, boxes, anchors):
        """box_torch_ops.py:30-84 (``second_box_encode``)."""
        nd = anchors.shape[-1]
        a = torch.split(anchors, 1, dim=-1)
        g = torch.split(boxes, 1, dim=-1)
        xa, ya, za, wa, la, ha = a[:6]
        xg, yg, zg, wg, lg, hg = g[:6]
        ra, rg = a[-1], g[-1]
        diagonal = torch.sqrt(la ** 2 + wa ** 2)
        ret = [(xg - xa) / diagonal, (yg - ya) / diagonal, (zg - za) / ha]
        if self.linear_dim:
            ret.extend([wg / wa - 1, lg / la - 1, hg / ha - 1])
        else:
            ret.extend([torch.log(wg / wa), torch.log(lg / la), torch.log(hg / ha)])
        if nd > 7:
            vxa, vya, vxg, vyg = a[6], a[7], g[6], g[7]
            ret.extend([vxg - vxa, vyg - vya])
        if self.vec_encode:
            ret.extend([torch.cos(rg) - torch.cos(ra), torch.sin(rg) - torch.sin(ra)])
        else:
            ret.append(rg - ra)
        return torch.cat(ret, dim=-1)


def build_box_coder(box_coder_config):
    """det3d/builder.py:65-100."""
    cfg = box_coder_config
    if cfg["type"] != "ground_box3d_coder":
        raise ValueError(f"unsupported box_coder type {cfg['type']!r} (the Waymo configs use ground_box3d_coder)")
    return GroundBox3dCoderTorch(cfg["linear_dim"], cfg["encode_angle_vector"], n_dim=cfg.get("n_dim", 9),
                                 norm_velo=cfg.get("norm_velo", False))
