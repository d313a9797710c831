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


class PFNLayer(nn.Module):
    """pillar_encoder.py:16-56 (module tree / state-dict keys: ``linear.weight``, ``norm.*``)."""

    def __init__(self, in_channels, out_channels, norm_cfg=None, last_layer=False):
        super().__init__()
        self.name = "PFNLayer"
        self.last_vfe = last_layer
        if not self.last_vfe:
            out_channels = out_channels // 2
        self.units = out_channels
        if norm_cfg is None:
            norm_cfg = dict(type="BN1d", eps=1e-3, momentum=0.01)
        self.norm_cfg = norm_cfg
        from .registry import build_norm_layer
        self.linear = nn.Linear(in_channels, self.units, bias=False)
        self.norm = build_norm_layer(self.norm_cfg, self.units)[1]


@READERS.register_module
class PillarFeatureNet(nn.Module):
    """pillar_encoder.py:59-154.  The two-layer net of the Waymo pillar configs (num_filters=[64, 64], 5 point features)
    runs as ONE fused kernel (``s2d_pfn_fwd``); eval mode only."""

    def __init__(self, num_input_features=4, num_filters=(64,), with_distance=False, voxel_size=(0.2, 0.2, 4),
                 pc_range=(0, -40, -3, 70.4, 40, 1), norm_cfg=None):
        super().__init__()
        self.name = "PillarFeatureNet"
        assert len(num_filters) > 0
        self.num_input = num_input_features
        num_input_features += 5
        if with_distance:
            num_input_features += 1
        self._with_distance = with_distance
        num_filters = [num_input_features] + list(num_filters)
        layers = []
        for i in range(len(num_filters) - 1):
            layers.append(PFNLayer(num_filters[i], num_filters[i + 1], norm_cfg=norm_cfg,
                                   last_layer=i >= len(num_filters) - 2))
        self.pfn_layers = nn.ModuleList(layers)
        self.vx = voxel_size[0]
        self.vy = voxel_size[1]
        self.x_offset = self.vx / 2 + pc_range[0]
        self.y_offset = self.vy / 2 + pc_range[1]
        self._fold = None

    def _folded(self):
        import torch
        srcs = [t for l in self.pfn_layers for t in (l.linear.weight, l.norm.weight, l.norm.bias, l.norm.running_mean,
                                                     l.norm.running_var)]
        key = tuple((t.data_ptr(), t._version) for t in srcs)
        if self._fold is None or self._fold[0] != key:
            out = []
            for l in self.pfn_layers:
                inv = torch.rsqrt(l.norm.running_var.float() + l.norm.eps)
                scale = (l.norm.weight.float() * inv).contiguous()
                shift = (l.norm.bias.float() - l.norm.running_mean.float() * scale).contiguous()
                out += [l.linear.weight.detach().float().contiguous(), scale.detach(), shift.detach()]
            self._fold = (key, out)
        return self._fold[1]

    def forward(self, features, num_voxels, coors):
        import torch
        from . import _lib
        if self.training:
            raise NotImplementedError("PillarFeatureNet is inference only in this build (call .eval())")
        ok = (len(self.pfn_layers) == 2 and not self._with_distance and self.num_input == 5 and
              self.pfn_layers[0].units == 32 and self.pfn_layers[1].units == 64 and features.shape[2] == 5)
        if not ok:
            raise NotImplementedError("only the Waymo pillar reader (5 features, num_filters=[64, 64]) is built")
        ops._need_cuda(features, num_voxels, coors)
        m, p, f = features.shape
        w0, s0, b0, w1, s1, b1 = self._folded()
        out = torch.empty((m, 64), dtype=torch.float32, device=features.device)
        _lib.check(_lib.load().s2d_pfn_fwd(features.contiguous().data_ptr(), num_voxels.int().contiguous().data_ptr(),
                                           coors.int().contiguous().data_ptr(), m, p, f, float(self.vx), float(self.vy),
                                           float(self.x_offset), float(self.y_offset), w0.data_ptr(), s0.data_ptr(),
                                           b0.data_ptr(), w1.data_ptr(), s1.data_ptr(), b1.data_ptr(), out.data_ptr(),
                                           ops._stream()), "s2d_pfn_fwd")
        return out
