"""BEV necks of CenterPoint / Sparse2Dense (det3d/models/necks/rpn.py).

``RPN`` (:24-162) and ``S2D_RPN`` (:164-337) keep the reference's constructor arguments, module tree and
state-dict keys (``encoder_1.0.weight``, ``blocks.0.1.weight``, ``deblocks.1.0.weight`` …) so that reference
checkpoints load; the eval-mode forward runs on NHWC rows through the tcgen05 gather-GEMM kernel with
BatchNorm / bias / GELU / ReLU / residual / concat fused (``sparse2dense_b200/dense.py``).  In training mode the same
layer sequence runs through the autograd operators of ``autograd.py`` (batch-statistics BatchNorm, backward kernels of
csrc/train.cu) and ``S2D_RPN`` adds the PCR branch (rpn.py:314-323).
"""
import logging

import numpy as np
import torch
from torch import nn
from torch.nn import Conv2d

from . import ops
from .dense import ACT_GELU, ACT_NONE, ACT_RELU, DenseOps, commit_twin, rows_with_twin, to_nchw, to_rows
from .registry import NECKS, build_norm_layer


@NECKS.register_module
class RPN(nn.Module):
    def __init__(self, layer_nums, ds_layer_strides, ds_num_filters, us_layer_strides, us_num_filters,
                 num_input_features, norm_cfg=None, name="rpn", logger=None, **kwargs):
        super(RPN, self).__init__()
        self._layer_strides = ds_layer_strides
        self._num_filters = ds_num_filters
        self._layer_nums = layer_nums
        self._upsample_strides = us_layer_strides
        self._num_upsample_filters = us_num_filters
        self._num_input_features = num_input_features
        if norm_cfg is None:
            norm_cfg = dict(type="BN", eps=1e-3, momentum=0.01)
        self._norm_cfg = norm_cfg
        assert len(self._layer_strides) == len(self._layer_nums)
        assert len(self._num_filters) == len(self._layer_nums)
        assert len(self._num_upsample_filters) == len(self._upsample_strides)
        self._upsample_start_idx = len(self._layer_nums) - len(self._upsample_strides)
        must_equal_list = []
        for i in range(len(self._upsample_strides)):
            must_equal_list.append(self._upsample_strides[i]
                                   / np.prod(self._layer_strides[: i + self._upsample_start_idx + 1]))
        for val in must_equal_list:
            assert val == must_equal_list[0]

        in_filters = [self._num_input_features, *self._num_filters[:-1]]
        blocks, deblocks = [], []
        for i, layer_num in enumerate(self._layer_nums):
            block, num_out_filters = self._make_layer(in_filters[i], self._num_filters[i], layer_num,
                                                      stride=self._layer_strides[i])
            blocks.append(block)
            if i - self._upsample_start_idx >= 0:
                stride = self._upsample_strides[i - self._upsample_start_idx]
                cout = self._num_upsample_filters[i - self._upsample_start_idx]
                if stride > 1:
                    deblock = nn.Sequential(nn.ConvTranspose2d(num_out_filters, cout, stride, stride=stride, bias=False),
                                            build_norm_layer(self._norm_cfg, cout)[1], nn.ReLU())
                else:
                    stride = int(np.round(1 / stride))
                    deblock = nn.Sequential(nn.Conv2d(num_out_filters, cout, stride, stride=stride, bias=False),
                                            build_norm_layer(self._norm_cfg, cout)[1], nn.ReLU())
                deblocks.append(deblock)
        self.blocks = nn.ModuleList(blocks)
        self.deblocks = nn.ModuleList(deblocks)
        self._dense = DenseOps()
        (logger or logging.getLogger("RPN")).info("Finish RPN Initialization")

    @property
    def downsample_factor(self):
        factor = np.prod(self._layer_strides)
        if len(self._upsample_strides) > 0:
            factor /= self._upsample_strides[-1]
        return factor

    def _make_layer(self, inplanes, planes, num_blocks, stride=1):
        mods = [nn.ZeroPad2d(1), nn.Conv2d(inplanes, planes, 3, stride=stride, bias=False),
                build_norm_layer(self._norm_cfg, planes)[1], nn.ReLU()]
        for j in range(num_blocks):
            mods.append(nn.Conv2d(planes, planes, 3, padding=1, bias=False))
            mods.append(build_norm_layer(self._norm_cfg, planes)[1])
            if j < num_blocks - 1:
                mods.append(nn.ReLU())
        return nn.Sequential(*mods), planes

    def set_precision(self, precision):
        self._dense = DenseOps(precision)

    # ---------------------------------------------------------------------------------------
    def _check_eval(self):
        """Select the execution mode of the dense operators: eval = fused inference kernels, train = autograd.py."""
        self._dense.training = self.training

    @staticmethod
    def _rows_in(x, training):
        B, C, H, W = x.shape
        if training:                              # differentiable layout change (torch permute: plumbing)
            return x.permute(0, 2, 3, 1).reshape(B * H * W, C)
        return to_rows(x)

    @staticmethod
    def _nchw_out(rows, B, H, W, training):
        if rows is None:
            return None
        if training:
            return rows.reshape(B, H, W, rows.shape[1]).permute(0, 3, 1, 2)
        return to_nchw(rows, B, H, W)

    def _run_block(self, i, x, B, H, W, outer_relu):
        """blocks[i] on rows: ZeroPad2d(1)+Conv3x3(stride) + BN + ReLU, then conv3x3 + BN (+ReLU) ... (rpn.py:126-145)."""
        mods = list(self.blocks[i])
        j, first = 0, True
        while j < len(mods):
            m = mods[j]
            if isinstance(m, nn.Conv2d):
                bn = mods[j + 1]
                relu = j + 2 < len(mods) and isinstance(mods[j + 2], nn.ReLU)
                last = not any(isinstance(q, nn.Conv2d) for q in mods[j + 1:])
                act = ACT_RELU if (relu or (last and outer_relu)) else ACT_NONE
                x, H, W = self._dense.conv(f"blocks.{i}.{j}", x, B, H, W, m, bn, act, pad=1 if first else None)
                first = False
                j += 2 + int(relu)
            else:
                j += 1
        return x, H, W

    def _run_deblock(self, i, x, B, H, W, out):
        m, bn = self.deblocks[i][0], self.deblocks[i][1]
        if isinstance(m, nn.ConvTranspose2d):
            return self._dense.tconv(f"deblocks.{i}", x, B, H, W, m, bn, ACT_RELU, out=out)
        return self._dense.conv(f"deblocks.{i}", x, B, H, W, m, bn, ACT_RELU, out=out)

    def _rpn_rows(self, x, B, H, W, outer_relu):
        """The block / deblock pyramid; returns the concatenated `ups` rows and its size."""
        n_up = len(self.deblocks)
        ups, H0, W0, off = None, None, None, 0
        for i in range(len(self.blocks)):
            x, H, W = self._run_block(i, x, B, H, W, outer_relu)
            d = i - self._upsample_start_idx
            if d >= 0:
                cout = self._num_upsample_filters[d]
                if ups is None:
                    s = self._upsample_strides[d]
                    H0, W0 = (int(H * s), int(W * s)) if s >= 1 else (int(H / round(1 / s)), int(W / round(1 / s)))
                    ups = rows_with_twin(B * H0 * W0, sum(self._num_upsample_filters), x.device)
                    launches = 0
                m = self.deblocks[d][0]
                launches += m.stride[0] * m.stride[1] if isinstance(m, nn.ConvTranspose2d) else 1   # sub-pixel classes
                self._run_deblock(d, x, B, H, W, ups[:, off:off + cout])
                off += cout
        if n_up == 0:
            return x, H, W
        if not self._dense.training:
            commit_twin(ups, launches)
        return ups, H0, W0

    def forward_rows(self, x, B, H, W):
        """RPN.forward on NHWC rows -> (ups rows, (Hu, Wu))."""
        self._check_eval()
        ups, Hu, Wu = self._rpn_rows(x, B, H, W, outer_relu=True)
        return ups, (Hu, Wu)

    def forward(self, x):
        """rpn.py:153-162 (note the outer F.relu after every block, which S2D_RPN.forward omits)."""
        self._check_eval()
        B, _, H, W = x.shape
        rows, H, W = self._rpn_rows(self._rows_in(x, self.training), B, H, W, outer_relu=True)
        return self._nchw_out(rows, B, H, W, self.training)


@NECKS.register_module
class S2D_RPN(RPN):
    def __init__(self, layer_nums, ds_layer_strides, ds_num_filters, us_layer_strides, us_num_filters,
                 num_input_features, norm_cfg=None, name="rpn", logger=None, **kwargs):
        super(S2D_RPN, self).__init__(layer_nums, ds_layer_strides, ds_num_filters, us_layer_strides, us_num_filters,
                                      num_input_features, norm_cfg, name, logger)
        C = num_input_features
        # S2D module (rpn.py:186-248)
        self.encoder_1 = nn.Sequential(Conv2d(C, 256, 2, 2), nn.BatchNorm2d(256), nn.GELU(),
                                       Conv2d(256, 256, 3, 1, 1), nn.BatchNorm2d(256), nn.GELU())
        self.encoder_2 = nn.Sequential(Conv2d(256, 256, 3, 2, 1), nn.BatchNorm2d(256), nn.GELU(),
                                       Conv2d(256, 256, 3, 1, 1), nn.BatchNorm2d(256), nn.GELU())

        def convnext():
            return nn.Sequential(nn.Conv2d(256, 256, kernel_size=7, padding=3, groups=256),
                                 nn.LayerNorm([256, 47, 47], eps=1e-6), nn.Conv2d(256, 256 * 4, 1, 1, 0), nn.GELU(),
                                 nn.Conv2d(256 * 4, 256, 1, 1, 0))
        self.convnext_block_1 = convnext()
        self.convnext_block_2 = convnext()
        self.convnext_block_3 = convnext()
        self.decoder_1 = nn.Sequential(nn.ConvTranspose2d(256, 256, 4, 2, 1), nn.BatchNorm2d(256), nn.GELU())
        self.decoder_2 = nn.Sequential(nn.Conv2d(512, 256, 3, 1, 1), nn.BatchNorm2d(256), nn.GELU(),
                                       nn.ConvTranspose2d(256, C, 4, 2, 1), nn.BatchNorm2d(C), nn.GELU())
        self.fusion_sparse = nn.Sequential(nn.Conv2d(C, C, 1, 1, 0), nn.BatchNorm2d(C), nn.GELU())
        self.fusion_dense = nn.Sequential(nn.Conv2d(C, C, 1, 1, 0), nn.BatchNorm2d(C), nn.GELU())
        self.out_conv = nn.Sequential(nn.Conv2d(C, 640, 1, 1, 0), nn.BatchNorm2d(640), nn.GELU())
        # PCR module (rpn.py:263-297): parameters only (train-time branch, kept for state-dict compatibility)
        self.generator_1 = nn.Sequential(nn.Conv3d(128, 32, 1, 1, 0), nn.BatchNorm3d(32), nn.ReLU(),
                                         nn.ConvTranspose3d(32, 32, 4, 2, 1), nn.BatchNorm3d(32), nn.ReLU())
        self.gen_out_4 = nn.Sequential(nn.Conv3d(32, 3, 1, 1, 0))
        self.gen_mask_4 = nn.Sequential(nn.Conv3d(32, 1, 1, 1, 0))
        self.generator_2 = nn.Sequential(nn.Conv3d(32, 16, 1, 1, 0), nn.BatchNorm3d(16), nn.ReLU(),
                                         nn.ConvTranspose3d(16, 3, 4, 2, 1), nn.BatchNorm3d(3), nn.ReLU())
        self.gen_out_2 = nn.Sequential(nn.Conv3d(3, 3, 1, 1, 0))
        self.gen_mask_2 = nn.Sequential(nn.Conv3d(3, 1, 1, 1, 0))
        self.train_pcr = True         # training mode runs the PCR branch (rpn.py:314-323)
        self.pcr_rows = None          # outputs of the last training forward's PCR branch (pcr.pcr_branch)

    def forward_rows(self, x, B, H, W):
        """S2D_RPN.forward (rpn.py:300-337) on NHWC rows ``x [B*H*W, C]`` -> (ups rows, F_S_a rows, F_S_b rows)."""
        self._check_eval()
        D = self._dense
        e1, e2 = self.encoder_1, self.encoder_2
        a, H1, W1 = D.conv("encoder_1.0", x, B, H, W, e1[0], e1[1], ACT_GELU)                      # 94
        y_3 = rows_with_twin(B * H1 * W1, 512, x.device)                                         # cat([decoder_1, y_1])
        y_1, _, _ = D.conv("encoder_1.3", a, B, H1, W1, e1[3], e1[4], ACT_GELU, out=y_3[:, 256:])
        a, H2, W2 = D.conv("encoder_2.0", y_1, B, H1, W1, e2[0], e2[1], ACT_GELU)                  # 47
        att, _, _ = D.conv("encoder_2.3", a, B, H2, W2, e2[3], e2[4], ACT_GELU)
        for bi, blk in enumerate((self.convnext_block_1, self.convnext_block_2, self.convnext_block_3)):
            t = D.dwconv(att, B, H2, W2, blk[0])
            t = D.layernorm(t, B, H2, W2, blk[1])
            t, _, _ = D.conv(f"convnext_block_{bi + 1}.2", t, B, H2, W2, blk[2], None, ACT_GELU)
            att, _, _ = D.conv(f"convnext_block_{bi + 1}.4", t, B, H2, W2, blk[4], None,
                               ACT_GELU if bi == 2 else ACT_NONE, residual=att)                    # (+ att), F.gelu on the last
        D.tconv("decoder_1.0", att, B, H2, W2, self.decoder_1[0], self.decoder_1[1], ACT_GELU, out=y_3[:, :256])
        commit_twin(y_3, 1 + self.decoder_1[0].stride[0] ** 2)                                     # one conv + the sub-pixel classes
        d2 = self.decoder_2
        a, _, _ = D.conv("decoder_2.0", y_3, B, H1, W1, d2[0], d2[1], ACT_GELU)
        F_S_b, _, _ = D.tconv("decoder_2.3", a, B, H1, W1, d2[3], d2[4], ACT_GELU)                 # 188
        self.pcr_rows = None
        if self.training and self.train_pcr:
            from . import pcr
            self.pcr_rows = pcr.pcr_branch(self, F_S_b, B, H, W)
        fs, _, _ = D.conv("fusion_sparse.0", x, B, H, W, self.fusion_sparse[0], self.fusion_sparse[1], ACT_GELU)
        F_S_a, _, _ = D.conv("fusion_dense.0", F_S_b, B, H, W, self.fusion_dense[0], self.fusion_dense[1], ACT_GELU,
                             residual=fs, res_after_act=True)                                      # gelu(.) + gelu(.)
        ups, Hu, Wu = self._rpn_rows(F_S_a, B, H, W, outer_relu=False)                             # rpn.py:327-335
        return ups, (Hu, Wu), F_S_a, F_S_b

    def forward(self, x):
        """x: NCHW [B,C,188,188] -> (x [B,512,188,188], None, None, None, None, F_S_a, F_S_b) like the reference."""
        B, _, H, W = x.shape
        t = self.training
        ups, (Hu, Wu), F_S_a, F_S_b = self.forward_rows(self._rows_in(x, t), B, H, W)
        gen = (None, None, None, None)
        if t and self.pcr_rows is not None:                           # NCDHW maps of the PCR branch (train only)
            from . import pcr
            gen = pcr.as_ncdhw(self.pcr_rows, B, H, W)
        return (self._nchw_out(ups, B, Hu, Wu, t),) + tuple(gen) + (self._nchw_out(F_S_a, B, H, W, t),
                                                                  self._nchw_out(F_S_b, B, H, W, t))
