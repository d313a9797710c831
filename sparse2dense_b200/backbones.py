"""Sparse 3-D conv backbone of CenterPoint-VoxelNet (det3d/models/backbones/scn.py:42-185).

Module tree, constructor arguments, forward signature, return value and state-dict keys are the
reference's; the arithmetic runs in libs2d_b200.so.  In eval mode every
``conv -> BatchNorm1d -> ReLU (-> += identity -> ReLU)`` group is one fused kernel launch and the
rulebooks of all four strided convs are resolved up front with a single host synchronisation.
"""
import numpy as np
import torch
from torch import nn

from . import ops, spconv
from .registry import BACKBONES, build_norm_layer
from .spconv import SparseConv3d, SubMConv3d


def conv3x3(in_planes, out_planes, stride=1, indice_key=None, bias=True):
    """3x3x3 submanifold convolution (scn.py:16-26)."""
    return spconv.SubMConv3d(in_planes, out_planes, kernel_size=3, stride=stride, padding=1, bias=bias,
                             indice_key=indice_key)


def conv1x1(in_planes, out_planes, stride=1, indice_key=None, bias=True):
    """1x1x1 submanifold convolution (scn.py:29-39)."""
    return spconv.SubMConv3d(in_planes, out_planes, kernel_size=1, stride=stride, padding=1, bias=bias,
                             indice_key=indice_key)


class SparseBasicBlock(spconv.SparseModule):
    """scn.py:42-85.  Note the reference quirk kept here: ``bias = norm_cfg is not None`` is always
    True because ``norm_cfg`` is defaulted first, so both convs carry a bias."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, norm_cfg=None, downsample=None, indice_key=None):
        super(SparseBasicBlock, self).__init__()
        if norm_cfg is None:
            norm_cfg = dict(type="BN1d", eps=1e-3, momentum=0.01)
        bias = norm_cfg is not None
        self.conv1 = conv3x3(inplanes, planes, stride, indice_key=indice_key, bias=bias)
        self.bn1 = build_norm_layer(norm_cfg, planes)[1]
        self.relu = nn.ReLU()
        self.conv2 = conv3x3(planes, planes, indice_key=indice_key, bias=bias)
        self.bn2 = build_norm_layer(norm_cfg, planes)[1]
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        identity = x
        if spconv._is_fusable_bn(self.bn1) and spconv._is_fusable_bn(self.bn2):
            out = self.conv1.fused_forward(x, bn=self.bn1, relu=True)
            if self.downsample is not None:
                identity = self.downsample(x)
            return self.conv2.fused_forward(out, bn=self.bn2, relu=True, residual=identity.features)
        out = self.conv1(x)
        out.features = self.bn1(out.features)
        out.features = self.relu(out.features)
        out = self.conv2(out)
        out.features = self.bn2(out.features)
        if self.downsample is not None:
            identity = self.downsample(x)
        out.features = out.features + identity.features
        out.features = self.relu(out.features)
        return out


@BACKBONES.register_module
class SpMiddleResNetFHD(nn.Module):
    def __init__(self, num_input_features=128, norm_cfg=None, name="SpMiddleResNetFH", is_student=False, **kwargs):
        super(SpMiddleResNetFHD, self).__init__()
        self.name = name
        self.dcn = None
        self.zero_init_residual = False
        self.is_student = is_student
        if norm_cfg is None:
            norm_cfg = dict(type="BN1d", eps=1e-3, momentum=0.01)

        self.conv_input = spconv.SparseSequential(
            SubMConv3d(num_input_features, 16, 3, bias=False, indice_key="res0"),
            build_norm_layer(norm_cfg, 16)[1],
            nn.ReLU(inplace=True),
        )
        self.conv1 = spconv.SparseSequential(
            SparseBasicBlock(16, 16, norm_cfg=norm_cfg, indice_key="res0"),
            SparseBasicBlock(16, 16, norm_cfg=norm_cfg, indice_key="res0"),
        )
        self.conv2 = spconv.SparseSequential(
            SparseConv3d(16, 32, 3, 2, padding=1, bias=False),
            build_norm_layer(norm_cfg, 32)[1],
            nn.ReLU(inplace=True),
            SparseBasicBlock(32, 32, norm_cfg=norm_cfg, indice_key="res1"),
            SparseBasicBlock(32, 32, norm_cfg=norm_cfg, indice_key="res1"),
        )
        self.conv3 = spconv.SparseSequential(
            SparseConv3d(32, 64, 3, 2, padding=1, bias=False),
            build_norm_layer(norm_cfg, 64)[1],
            nn.ReLU(inplace=True),
            SparseBasicBlock(64, 64, norm_cfg=norm_cfg, indice_key="res2"),
            SparseBasicBlock(64, 64, norm_cfg=norm_cfg, indice_key="res2"),
        )
        self.conv4 = spconv.SparseSequential(
            SparseConv3d(64, 128, 3, 2, padding=[0, 1, 1], bias=False),
            build_norm_layer(norm_cfg, 128)[1],
            nn.ReLU(inplace=True),
            SparseBasicBlock(128, 128, norm_cfg=norm_cfg, indice_key="res3"),
            SparseBasicBlock(128, 128, norm_cfg=norm_cfg, indice_key="res3"),
        )
        self.extra_conv = spconv.SparseSequential(
            SparseConv3d(128, 128, (3, 1, 1), (2, 1, 1), bias=False),
            build_norm_layer(norm_cfg, 128)[1],
            nn.ReLU(),
        )

    def set_precision(self, precision):
        """Choose the arithmetic of the sparse convs: ops.PRECISION_FP32 (CUDA cores), PRECISION_TF32X3
        (tcgen05, split TF32, ~fp32 accuracy) or PRECISION_TF32 (tcgen05 single pass).  Layers whose
        shape has no tensor-core kernel (Cin < 32) stay on the fp32 kernel."""
        for m in self.modules():
            if isinstance(m, spconv.SparseConvolution):
                ok = precision != ops.PRECISION_FP32 and ops.tf32_supported(m.in_channels, m.out_channels)
                m.precision = precision if ok else ops.PRECISION_FP32
                # AUTO / BF16X2, inference only: a narrow input (the 5-channel first layer) is zero-padded onto the
                # tensor-core kernel; its training path stays on the fp32 kernel (m.precision)
                m.pad_narrow_input = (precision in (ops.PRECISION_AUTO, ops.PRECISION_BF16X2) and m.in_channels < 16 and
                                      m.out_channels % 16 == 0)

    def bev_hw(self, input_shape):
        """(H, W) of the BEV map this backbone produces for a voxel grid ``input_shape`` = (x, y, z)."""
        shape = tuple(int(v) for v in (np.array(input_shape[::-1]) + [1, 0, 0]))
        for m in (self.conv2[0], self.conv3[0], self.conv4[0], self.extra_conv[0]):
            shape = ops.conv_out_shape(shape, m.kernel_size, m.stride, m.padding, m.dilation)
        return int(shape[1]), int(shape[2])

    def forward(self, voxel_features, coors, batch_size, input_shape, index=None, as_rows=False):
        """scn.py:156-185.  ``as_rows=True`` returns the BEV map as NHWC rows ``[B*H*W, C*D]`` (what the dense
        stage of this package consumes) instead of the reference's NCHW ``[B, C*D, H, W]``."""
        sparse_shape = np.array(input_shape[::-1]) + [1, 0, 0]          # scn.py:159 (depth 41, not 40)
        coors = coors.int().contiguous()
        ret = spconv.SparseConvTensor(voxel_features, coors, sparse_shape, batch_size)
        ret._index = index
        # coordinate phase of conv2.0, conv3.0, conv4.0, extra_conv.0 with one host sync
        spconv.plan_coords(ret, [self.conv2[0], self.conv3[0], self.conv4[0], self.extra_conv[0]])

        x = self.conv_input(ret)
        x_conv1 = self.conv1(x)
        x_conv2 = self.conv2(x_conv1)
        x_conv3 = self.conv3(x_conv2)
        x_conv4 = self.conv4(x_conv3)
        ret = self.extra_conv(x_conv4)

        if torch.is_grad_enabled() and ret.features.requires_grad:
            from .autograd import DenseBEV
            ret = DenseBEV.apply(ret.features, ret.indices, batch_size, ret.spatial_shape, bool(as_rows))
        elif as_rows:
            ret = ops.dense_bev_rows(ret.features, ret.indices, batch_size, ret.spatial_shape)
        else:
            ret = ret.dense()
            N, C, D, H, W = ret.shape
            ret = ret.view(N, C * D, H, W)

        multi_scale_voxel_features = {
            "conv1": x_conv1,
            "conv2": x_conv2,
            "conv3": x_conv3,
            "conv4": x_conv4,
        }
        return ret, multi_scale_voxel_features


@BACKBONES.register_module
class SpMiddleFHD(nn.Module):
    """The SECOND backbone (det3d/models/backbones/scn.py:187-289): the plain (non-residual) sparse stack
    16-16 | 32-32-32 | 64-64-64-64 | 64-64-64-64 + the (3,1,1) ``extra_conv``; same constructor arguments, module tree and
    state-dict keys (``middle_conv.0.weight`` ...), returns ``(dense [B, 128, H, W], conv_4 SparseConvTensor)``.
    Every conv -> BatchNorm1d -> ReLU group is one fused launch in eval mode and runs the training operators in train mode
    (``spconv.SparseSequential``)."""

    def __init__(self, num_input_features=128, norm_cfg=None, name="SpMiddleFHD", **kwargs):
        super(SpMiddleFHD, self).__init__()
        self.name = name
        self.dcn = None
        self.zero_init_residual = False
        if norm_cfg is None:
            norm_cfg = dict(type="BN1d", eps=1e-3, momentum=0.01)

        def group(conv):
            return [conv, build_norm_layer(norm_cfg, conv.out_channels)[1], nn.ReLU()]
        layers = []
        layers += group(SubMConv3d(num_input_features, 16, 3, bias=False, indice_key="subm0"))
        layers += group(SubMConv3d(16, 16, 3, bias=False, indice_key="subm0"))
        layers += group(SparseConv3d(16, 32, 3, 2, padding=1, bias=False))
        layers += group(SubMConv3d(32, 32, 3, indice_key="subm1", bias=False))
        layers += group(SubMConv3d(32, 32, 3, indice_key="subm1", bias=False))
        layers += group(SparseConv3d(32, 64, 3, 2, padding=1, bias=False))
        for _ in range(3):
            layers += group(SubMConv3d(64, 64, 3, indice_key="subm2", bias=False))
        layers += group(SparseConv3d(64, 64, 3, 2, padding=[0, 1, 1], bias=False))
        for _ in range(3):
            layers += group(SubMConv3d(64, 64, 3, indice_key="subm3", bias=False))
        self.middle_conv = spconv.SparseSequential(*layers)
        self.extra_conv = spconv.SparseSequential(*group(SparseConv3d(64, 64, (3, 1, 1), (2, 1, 1), bias=False)))

    set_precision = SpMiddleResNetFHD.set_precision

    def _strided(self):
        return [m for m in list(self.middle_conv) + list(self.extra_conv)
                if isinstance(m, spconv.SparseConvolution) and not m.subm]

    def bev_hw(self, input_shape):
        shape = tuple(int(v) for v in (np.array(input_shape[::-1]) + [1, 0, 0]))
        for m in self._strided():
            shape = ops.conv_out_shape(shape, m.kernel_size, m.stride, m.padding, m.dilation)
        return int(shape[1]), int(shape[2])

    def forward(self, voxel_features, coors, batch_size, input_shape):
        sparse_shape = np.array(input_shape[::-1]) + [1, 0, 0]          # scn.py:274
        coors = coors.int().contiguous()
        ret = spconv.SparseConvTensor(voxel_features, coors, sparse_shape, batch_size)
        spconv.plan_coords(ret, self._strided())                        # all output coordinate sets, one host sync
        conv_4 = self.middle_conv(ret)
        ret = self.extra_conv(conv_4)
        if torch.is_grad_enabled() and ret.features.requires_grad:
            from .autograd import DenseBEV
            ret = DenseBEV.apply(ret.features, ret.indices, batch_size, ret.spatial_shape, False)
        else:
            ret = ret.dense()
            N, C, D, H, W = ret.shape
            ret = ret.view(N, C * D, H, W)
        return ret, conv_4


@BACKBONES.register_module
class PointPillarsScatter_S2D(nn.Module):
    """Pillar scatter + the S2D module of the pillar student (det3d/models/readers/pillar_encoder.py:219-394, registered
    under BACKBONES there too).  Same module tree / state-dict keys; eval forward (the PCR generator is train-only).
    The canvas is built directly as NHWC rows (``s2d_dense_bev_nhwc`` with D = 1) and every layer runs on rows."""

    def __init__(self, num_input_features=64, norm_cfg=None, name="PointPillarsScatter", **kwargs):
        super().__init__()
        self.name = "PointPillarsScatter"
        self.nchannels = num_input_features
        S = nn.Sequential

        def convnext():
            return S(nn.Conv2d(256, 256, kernel_size=7, padding=3, groups=256), nn.LayerNorm([256, 59, 59], eps=1e-6),
                     nn.Conv2d(256, 1024, 1, 1, 0), nn.GELU(), nn.Conv2d(1024, 256, 1, 1, 0))
        self.encoder_1 = S(nn.MaxPool2d(2, 2), nn.Conv2d(64, 32, 1, 1, 0), nn.BatchNorm2d(32), nn.GELU(),
                           nn.Conv2d(32, 32, 2, 2), nn.BatchNorm2d(32), nn.GELU(),
                           nn.Conv2d(32, 128, 1, 1, 0), nn.BatchNorm2d(128), nn.GELU())
        self.encoder_2 = S(nn.Conv2d(128, 128, 3, 2, 1), nn.BatchNorm2d(128), nn.GELU(),
                           nn.Conv2d(128, 256, 3, 1, 1), nn.BatchNorm2d(256), nn.GELU())
        self.convnext_block_1 = convnext()
        self.convnext_block_2 = convnext()
        self.convnext_block_3 = convnext()
        self.decoder_1 = S(nn.Conv2d(256, 128, 3, 1, 1), nn.BatchNorm2d(128), nn.GELU(), nn.Upsample((117, 117)))
        self.decoder_2 = S(nn.Conv2d(128 + 128, 64, 3, 1, 1), nn.BatchNorm2d(64), nn.GELU(),
                           nn.ConvTranspose2d(64, 64, 4, 2, 1), nn.BatchNorm2d(64), nn.GELU(),
                           nn.Conv2d(64, 64, 1, 1, 0), nn.BatchNorm2d(64), nn.GELU(), nn.Upsample(scale_factor=2))
        self.fusion_sparse = S(nn.Conv2d(64, 64, 1, 1, 0), nn.BatchNorm2d(num_input_features), nn.GELU())
        self.fusion_dense = S(nn.Conv2d(64, 64, 1, 1, 0), nn.BatchNorm2d(64), nn.GELU())
        # PCR module (train only; kept for state-dict compatibility)
        self.generator = S(nn.Conv3d(64, 32, 1, 1, 0), nn.BatchNorm3d(32), nn.GELU(),
                           nn.Conv3d(32, 16, 1, 1, 0), nn.BatchNorm3d(16), nn.GELU())
        self.gen_out = S(nn.Conv3d(16, 3, 1, 1, 0))
        self.gen_mask = S(nn.Conv3d(16, 8, 1, 1, 0), nn.BatchNorm3d(8), nn.GELU(), nn.Conv3d(8, 1, 1, 1, 0))
        from .dense import DenseOps
        self._dense = DenseOps()

    def set_precision(self, precision):
        from .dense import DenseOps
        self._dense = DenseOps(precision)

    def forward_rows(self, voxel_features, coords, batch_size, input_shape):
        """-> (F_S_a rows [B*ny*nx, 64], F_S_b rows, (ny, nx))."""
        from .dense import ACT_GELU, ACT_NONE
        if self.training:
            raise NotImplementedError("PointPillarsScatter_S2D is inference only in this build (call .eval())")
        D = self._dense
        nx, ny = int(input_shape[0]), int(input_shape[1])
        B = batch_size
        assert (ny, nx) == (468, 468), "LayerNorm([256,59,59]) / Upsample((117,117)) fix the canvas to 468 x 468"
        canvas = ops.dense_bev_rows(voxel_features.contiguous(), coords.int().contiguous(), B, (1, ny, nx))   # scatter
        e1, e2 = self.encoder_1, self.encoder_2
        a, H1, W1 = D.maxpool2(canvas, B, ny, nx)                                                   # 234
        a, _, _ = D.conv("encoder_1.1", a, B, H1, W1, e1[1], e1[2], ACT_GELU)
        a, H2, W2 = D.conv("encoder_1.4", a, B, H1, W1, e1[4], e1[5], ACT_GELU)                    # 117
        y_3 = torch.empty((B * H2 * W2, 256), dtype=torch.float32, device=canvas.device)            # cat([decoder_1, y_1])
        y_1, _, _ = D.conv("encoder_1.7", a, B, H2, W2, e1[7], e1[8], ACT_GELU, out=y_3[:, 128:])
        a, H3, W3 = D.conv("encoder_2.0", y_1, B, H2, W2, e2[0], e2[1], ACT_GELU)                   # 59
        att, _, _ = D.conv("encoder_2.3", a, B, H3, W3, e2[3], e2[4], ACT_GELU)
        for bi, blk in enumerate((self.convnext_block_1, self.convnext_block_2, self.convnext_block_3)):
            t = D.dwconv(att, B, H3, W3, blk[0])
            t = D.layernorm(t, B, H3, W3, blk[1])
            t, _, _ = D.conv(f"convnext_block_{bi + 1}.2", t, B, H3, W3, blk[2], None, ACT_GELU)
            att, _, _ = D.conv(f"convnext_block_{bi + 1}.4", t, B, H3, W3, blk[4], None, ACT_NONE, residual=att)
        d1 = self.decoder_1
        a, _, _ = D.conv("decoder_1.0", att, B, H3, W3, d1[0], d1[1], ACT_GELU)
        D.upsample_nearest(a, B, H3, W3, H2, W2, out=y_3[:, :128])                                  # Upsample((117,117))
        d2 = self.decoder_2
        a, _, _ = D.conv("decoder_2.0", y_3, B, H2, W2, d2[0], d2[1], ACT_GELU)
        a, H4, W4 = D.tconv("decoder_2.3", a, B, H2, W2, d2[3], d2[4], ACT_GELU)                   # 234
        a, _, _ = D.conv("decoder_2.6", a, B, H4, W4, d2[6], d2[7], ACT_GELU)
        F_S_b = D.upsample_nearest(a, B, H4, W4, ny, nx, scale=2)                                   # 468
        fs, _, _ = D.conv("fusion_sparse.0", canvas, B, ny, nx, self.fusion_sparse[0], self.fusion_sparse[1], ACT_GELU)
        F_S_a, _, _ = D.conv("fusion_dense.0", F_S_b, B, ny, nx, self.fusion_dense[0], self.fusion_dense[1], ACT_GELU,
                             residual=fs, res_after_act=True)
        return F_S_a, F_S_b, (ny, nx)

    def forward(self, voxel_features, coords, batch_size, input_shape):
        """Reference return value (eval): (F_S_a, F_S_b, None, None) as NCHW maps."""
        from .dense import to_nchw
        a, b, (H, W) = self.forward_rows(voxel_features, coords, batch_size, input_shape)
        return to_nchw(a, batch_size, H, W), to_nchw(b, batch_size, H, W), None, None


@BACKBONES.register_module
class PointPillarsScatter(nn.Module):
    """Plain pillar scatter of the teacher / baseline pillar models (pillar_encoder.py:157-217): the canvas
    ``[B, 64, ny, nx]`` with each pillar's feature at ``(y, x)``; one launch (``s2d_dense_bev_nhwc`` with D = 1)."""

    def __init__(self, num_input_features=64, norm_cfg=None, name="PointPillarsScatter", **kwargs):
        super().__init__()
        self.name = "PointPillarsScatter"
        self.nchannels = num_input_features

    def forward_rows(self, voxel_features, coords, batch_size, input_shape):
        nx, ny = int(input_shape[0]), int(input_shape[1])
        rows = ops.dense_bev_rows(voxel_features.contiguous(), coords.int().contiguous(), batch_size, (1, ny, nx))
        return rows, (ny, nx)

    def forward(self, voxel_features, coords, batch_size, input_shape):
        from .dense import to_nchw
        rows, (ny, nx) = self.forward_rows(voxel_features, coords, batch_size, input_shape)
        return to_nchw(rows, batch_size, ny, nx)
