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
        if spconv._is_eval_bn(self.bn1) and spconv._is_eval_bn(self.bn2):
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

        if as_rows:
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
