"""Host-side mirror of the spconv v1.x Python API that the reference backbone is written against.

Names, constructor arguments, attributes and state-dict layout follow spconv @ 7342772 as the
reference uses it (det3d/models/backbones/scn.py:2,8,16-39,42-85,104-152,159-176;
det3d/models/detectors/voxelnet.py:203-215), so reference-style model code keeps working:

    SparseConvTensor(features, indices, spatial_shape, batch_size)  .features .indices
        .spatial_shape .batch_size .indice_dict .dense()
    SubMConv3d / SparseConv3d(in, out, kernel_size, stride=1, padding=0, dilation=1, groups=1,
        bias=True, indice_key=None)      weight: Parameter[kD,kH,kW,Cin,Cout]
    SparseSequential(*modules), SparseModule

Underneath, every step is one call into libs2d_b200.so (output-stationary gather-GEMM with a
fused BatchNorm/ReLU/residual epilogue) -- see include/s2d_b200.h.  In training mode (batch-statistics
BatchNorm) the layers run on the autograd operators of ``autograd.py``, whose backward is the same
gather-GEMM over the transposed rulebook plus ``s2d_conv_wgrad``.  There is no CPU implementation.
"""
import math
from collections import OrderedDict

import numpy as np
import torch
from torch import nn

from . import ops


class SparseModule(nn.Module):
    """Marker base class: modules that consume / produce SparseConvTensor (scn.py:42)."""


class _Indice:
    """What spconv keeps per ``indice_key``: the rulebook and the output coordinate set."""

    def __init__(self, tbl, out_indices, out_index, out_shape, n_pairs=None, build=None):
        self._tbl, self._build = tbl, build
        self.out_indices, self.out_index, self.out_shape, self.n_pairs = out_indices, out_index, out_shape, n_pairs

    @property
    def tbl(self):
        """The scan-order rulebook ``i32 [K, n_out]``; built on first use (the inference path of a 3x3x3 submanifold layer
        only ever needs the grouped table, ``ops.rulebook_subm_grouped``)."""
        if self._tbl is None:
            self._tbl = self._build()
        return self._tbl


class SparseConvTensor:
    def __init__(self, features, indices, spatial_shape, batch_size, grid=None):
        self.features = features
        self.indices = indices
        self.spatial_shape = tuple(int(v) for v in np.asarray(spatial_shape).tolist())
        self.batch_size = int(batch_size)
        self.indice_dict = {}
        self.grid = grid
        self._index = None            # ops.GridIndex of this tensor's coordinate set
        self._planned = {}            # id(module) -> ops.SparseCoords precomputed by plan_coords()

    @property
    def spatial_size(self):
        return int(np.prod(self.spatial_shape))

    def find_indice_pair(self, key):
        if key is None:
            return None
        return self.indice_dict.get(key)

    def index(self):
        if self._index is None:
            self._index = ops.build_grid_index(self.indices, self.batch_size, self.spatial_shape)
        return self._index

    def _like(self, features, indices=None, spatial_shape=None, index=None):
        out = SparseConvTensor(features, self.indices if indices is None else indices,
                               self.spatial_shape if spatial_shape is None else spatial_shape, self.batch_size)
        out.indice_dict = self.indice_dict
        out._planned = self._planned
        out._index = index if indices is not None else self._index
        return out

    def dense(self, channels_first=True):
        """[B, C, D, H, W] like spconv (zeros scattered with the active rows, then permuted)."""
        d, h, w = self.spatial_shape
        bev = ops.dense_bev(self.features, self.indices, self.batch_size, self.spatial_shape)
        out = bev.view(self.batch_size, self.features.shape[1], d, h, w)
        return out if channels_first else out.permute(0, 2, 3, 4, 1).contiguous()


def _fold_bn(bn, bias, cout, device):
    """eval-mode BatchNorm1d (+ conv bias) -> per-channel (scale, shift)."""
    if bn is None:
        if bias is None:
            return None, None
        return None, bias.detach().float().contiguous()
    inv = torch.rsqrt(bn.running_var.float() + bn.eps)
    gamma = bn.weight.float() if bn.weight is not None else torch.ones(cout, device=device)
    beta = bn.bias.float() if bn.bias is not None else torch.zeros(cout, device=device)
    scale = gamma * inv
    shift = beta - bn.running_mean.float() * scale
    if bias is not None:
        shift = shift + bias.float() * scale
    return scale.detach().contiguous(), shift.detach().contiguous()


# group the rows of 27-offset rulebooks by neighbour pattern before the tile kernel (ops.table_group_rows):
# 0 never, 1 submanifold rulebooks only (shared by the four convolutions of a stage), 2 strided rulebooks too (default:
# both kinds are built directly in grouped order, which costs about as much as the scan-order table)
GROUP_ROWS = int(__import__("os").environ.get("S2D_GROUP_ROWS", "2"))


class SparseConvolution(SparseModule):
    def __init__(self, ndim, in_channels, out_channels, kernel_size=3, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, subm=False, indice_key=None):
        super().__init__()
        assert ndim == 3 and groups == 1
        self.ndim = ndim
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size = ops._triple(kernel_size)
        self.stride = ops._triple(stride)
        self.padding = ops._triple(padding)
        self.dilation = ops._triple(dilation)
        self.groups = groups
        self.subm = subm
        self.indice_key = indice_key
        self.precision = ops.PRECISION_FP32
        self._packed = None           # (key, tensor): tcgen05 weight image, rebuilt when the weight changes
        self.weight = nn.Parameter(torch.empty(*self.kernel_size, in_channels, out_channels))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in, _ = nn.init._calculate_fan_in_and_fan_out(self.weight)
            bound = 1 / math.sqrt(fan_in)
            nn.init.uniform_(self.bias, -bound, bound)

    def extra_repr(self):
        return (f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, stride={self.stride}, "
                f"padding={self.padding}, subm={self.subm}, indice_key={self.indice_key}")

    # -- rulebook ---------------------------------------------------------------------------
    def _indice(self, x):
        cached = x.find_indice_pair(self.indice_key)
        if cached is not None:
            return cached
        if self.subm:
            index = x.index()
            coors, ks, dl = x.indices, self.kernel_size, self.dilation
            ind = _Indice(None, coors, index, x.spatial_shape, build=lambda: ops.rulebook_subm(coors, index, ks, dl))
        else:
            coords = x._planned.get(id(self))
            if coords is None:
                coords = ops.sparse_out_coords(x.indices, x.indices.shape[0], x.batch_size, x.spatial_shape,
                                               self.kernel_size, self.stride, self.padding, self.dilation)
            out_indices = coords.coors          # host count read here unless plan_coords() already did
            index_in, ks, st, pd, dl = x.index(), self.kernel_size, self.stride, self.padding, self.dilation
            ind = _Indice(None, out_indices, coords.index, coords.shape,
                          build=lambda: ops.rulebook_sparse(out_indices, index_in, ks, st, pd, dl))
            ind.index_in = index_in             # for the grouped builder (ops.rulebook_sparse_grouped)
        if self.indice_key is not None:
            x.indice_dict[self.indice_key] = ind
        return ind

    # -- forward ----------------------------------------------------------------------------
    def fused_forward(self, x, bn=None, relu=False, residual=None):
        """conv (+ eval BatchNorm1d) (+ residual) (+ ReLU) in one kernel launch."""
        assert isinstance(x, SparseConvTensor)
        ind = self._indice(x)
        if (bn is not None and bn.training) or (bn is None and self.training and torch.is_grad_enabled()):
            return self._train_forward(x, ind, bn, relu, residual)
        if torch.is_grad_enabled() and (self.weight.requires_grad and self.training or x.features.requires_grad):
            # an eval-mode (frozen) BatchNorm inside a training graph would silently cut the gradient here
            raise NotImplementedError("frozen (eval-mode) BatchNorm inside a training graph is not built "
                                      "(use torch.no_grad() for inference or bn.train() for training)")
        scale, shift = self._folded(bn, x.features.device)
        n_out = ind.out_indices.shape[0]
        K = self.kernel_size[0] * self.kernel_size[1] * self.kernel_size[2]
        feats_in, weight = x.features.contiguous(), self.weight.detach()
        cin = self.in_channels
        auto = self.precision in (ops.PRECISION_AUTO, ops.PRECISION_BF16X2) or self.__dict__.get("pad_narrow_input", False)
        if auto and cin < 16 and ops.bf2_shape_ok(16, self.out_channels, K):
            # the 5-channel input layer: zero-pad to 16 channels so that it runs on the tensor-core kernel over the stage's
            # grouped rulebook like every other layer (0.33 -> 0.12 ms at batch 8; the CUDA-core kernel needs the scan-order table)
            cin = 16
            feats_in = torch.nn.functional.pad(feats_in, (0, 16 - self.in_channels))
            weight = self._padded_weight(16)
        if auto and ops.bf2_shape_ok(cin, self.out_channels, K):
            prec = ops.PRECISION_BF16X2                       # every rulebook of this module comes from ops.alloc_table
        else:
            prec = ops.effective_precision(self.precision, cin, self.out_channels, ind.tbl)
        packed = self._packed_weights(prec, cin) if prec != ops.PRECISION_FP32 else None
        masks, tbl, out_rows = None, None, None
        if prec == ops.PRECISION_BF16X2:
            # once per rulebook: rows grouped by their live offset triples (bit-identical results, far fewer live
            # (tile, offset) pairs -- conv_bf2.cu "Row grouping") and the per-tile live-offset masks of the grouped table;
            # a 3x3x3 submanifold rulebook is built directly in grouped order, the scan-order table is never written
            grouped = ind.__dict__.get("grouped")
            if grouped is None:
                if K == 27 and GROUP_ROWS >= (1 if self.subm else 2):
                    direct = ind._tbl is None and self.kernel_size == (3, 3, 3)
                    if direct and self.subm:
                        grouped = ops.rulebook_subm_grouped(ind.out_indices, ind.out_index, self.dilation)
                    elif direct and "index_in" in ind.__dict__:
                        grouped = ops.rulebook_sparse_grouped(ind.out_indices, ind.index_in, self.stride, self.padding,
                                                              self.dilation)
                    else:
                        grouped = ops.table_group_rows(ind.tbl, n_out)
                else:
                    grouped = (ind.tbl, None, ops.table_tile_masks(ind.tbl, n_out))
                ind.grouped = grouped
            tbl, out_rows, masks = grouped
        else:
            tbl = ind.tbl
        feats = ops.spconv_fwd(feats_in, weight, tbl, n_out, scale, shift,
                               None if residual is None else residual.contiguous(), relu, prec,
                               packed=packed, tile_masks=masks, out_rows=out_rows)
        if self.subm:
            return x._like(feats)
        return x._like(feats, ind.out_indices, ind.out_shape, ind.out_index)

    def _train_forward(self, x, ind, bn, relu, residual):
        """Training mode: conv -> BatchNorm1d with batch statistics over the active rows (scn.py:100-107) -> (+ residual)
        -> ReLU, each with its backward in libs2d_b200.so (autograd.py)."""
        from . import autograd as AG
        table = ind.__dict__.get("table")
        K = self.kernel_size[0] * self.kernel_size[1] * self.kernel_size[2]
        if table is None:
            n_out = ind.out_indices.shape[0]
            grouped = ind.__dict__.get("grouped")
            if grouped is not None and grouped[1] is None:
                grouped = None                              # masks of an ungrouped table: not what the training path wants
            if grouped is None and K == 27 and n_out > 0 and GROUP_ROWS >= (1 if self.subm else 2):
                grouped = ops.table_group_rows(ind.tbl, n_out)
            table = ind.table = AG.Table(ind.tbl, x.features.shape[0], n_out, symmetric=self.subm, grouped=grouped)
        w = self.weight.view(K, self.in_channels, self.out_channels)
        y = AG.GatherConv.apply(x.features, w, table, self.precision)
        act = 1 if relu else 0
        if bn is None:
            if self.bias is not None or relu or residual is not None:
                y = AG.norm_act(y, None, self.bias, act, residual)
        else:
            y = AG.norm_act(y, bn, None, act, residual, pre_bias=self.bias)
        if self.subm:
            return x._like(y)
        return x._like(y, ind.out_indices, ind.out_shape, ind.out_index)

    def forward(self, x):
        return self.fused_forward(x)

    def _folded(self, bn, device):
        """(scale, shift) of the eval-mode BatchNorm (+ conv bias), cached until a parameter changes."""
        tensors = [self.bias] if bn is None else [self.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var]
        key = (id(bn), str(device)) + tuple((t.data_ptr(), t._version) if t is not None else None for t in tensors)
        cache = self.__dict__.get("_fold_cache")
        if cache is None or cache[0] != key:
            cache = (key,) + _fold_bn(bn, self.bias, self.out_channels, device)
            self.__dict__["_fold_cache"] = cache
        return cache[1], cache[2]

    def _padded_weight(self, cin):
        """The weight with its input-channel axis zero-padded to ``cin`` (cached until the weight changes)."""
        key = (self.weight.data_ptr(), self.weight._version, str(self.weight.device), cin)
        hit = self.__dict__.get("_pad_cache")
        if hit is None or hit[0] != key:
            w = torch.nn.functional.pad(self.weight.detach(), (0, 0, 0, cin - self.in_channels)).contiguous()
            hit = self.__dict__["_pad_cache"] = (key, w)
        return hit[1]

    def _packed_weights(self, precision=None, cin=None):
        precision = self.precision if precision is None else precision
        cin = self.in_channels if cin is None else cin
        key = (self.weight.data_ptr(), self.weight._version, str(self.weight.device), precision, cin)
        if self._packed is None or self._packed[0] != key:
            w = self.weight if cin == self.in_channels else self._padded_weight(cin)
            self._packed = (key, ops.pack_weights_tf32(w, precision))
        return self._packed[1]


class SubMConv3d(SparseConvolution):
    """Submanifold conv: output set == input set; `padding`/`stride` are ignored as in spconv
    (scn.py:16-26 passes padding=1, scn.py:105 passes nothing; both mean a centred kernel)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, True,
                         indice_key)


class SparseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True,
                 indice_key=None):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias, False,
                         indice_key)


def _is_eval_bn(m):
    return isinstance(m, (nn.BatchNorm1d, nn.SyncBatchNorm)) and not m.training and m.track_running_stats


def _is_fusable_bn(m):
    """eval-mode BN folds into the conv epilogue; training-mode BN runs the batch-statistics kernels of train.cu."""
    return isinstance(m, (nn.BatchNorm1d, nn.SyncBatchNorm)) and m.track_running_stats


class SparseSequential(SparseModule):
    """spconv.SparseSequential: sparse modules get the tensor, others are applied to ``.features``.

    In eval mode a ``conv -> BatchNorm1d -> ReLU`` run is executed as ONE fused kernel; the
    result is identical up to fp32 rounding of the folded affine.
    """

    def __init__(self, *args, **kwargs):
        super().__init__()
        if len(args) == 1 and isinstance(args[0], OrderedDict):
            for key, module in args[0].items():
                self.add_module(key, module)
        else:
            for idx, module in enumerate(args):
                self.add_module(str(idx), module)
        for name, module in kwargs.items():
            if name in self._modules:
                raise ValueError("name exists.")
            self.add_module(name, module)

    def __getitem__(self, idx):
        if not (-len(self) <= idx < len(self)):
            raise IndexError("index {} is out of range".format(idx))
        if idx < 0:
            idx += len(self)
        return list(self._modules.values())[idx]

    def __len__(self):
        return len(self._modules)

    def add(self, module, name=None):
        if name is None:
            name = str(len(self._modules))
            if name in self._modules:
                raise KeyError("name exists")
        self.add_module(name, module)

    def forward(self, input):
        mods = list(self._modules.values())
        i = 0
        while i < len(mods):
            m = mods[i]
            if isinstance(m, SparseConvolution) and isinstance(input, SparseConvTensor):
                bn = mods[i + 1] if i + 1 < len(mods) and _is_fusable_bn(mods[i + 1]) else None
                relu = bn is not None and i + 2 < len(mods) and isinstance(mods[i + 2], nn.ReLU)
                input = m.fused_forward(input, bn=bn, relu=relu)
                i += 1 + (bn is not None) + int(relu)
            elif isinstance(m, SparseModule):
                input = m(input)
                i += 1
            else:
                if isinstance(input, SparseConvTensor):
                    if input.indices.shape[0] != 0:
                        input.features = m(input.features)
                else:
                    input = m(input)
                i += 1
        return input


def plan_coords(x, modules):
    """Run the coordinate phase of every strided conv in ``modules`` (in forward order) without a
    host round trip, then read all output counts in ONE device->host copy.

    Coordinates do not depend on features, so the whole chain voxels -> conv2.0 -> conv3.0 -> ...
    can be resolved before the first feature kernel runs.  Results are parked on the tensor and
    picked up by ``SparseConv3d`` when it executes.
    """
    convs = [m for m in modules if isinstance(m, SparseConvolution) and not m.subm]
    coors, n, n_dev, shape = x.indices, x.indices.shape[0], None, x.spatial_shape
    planned = []
    for m in convs:
        sc = ops.sparse_out_coords(coors, n, x.batch_size, shape, m.kernel_size, m.stride, m.padding, m.dilation,
                                   n_in_dev=n_dev)
        planned.append(sc)
        coors, n, n_dev, shape = sc.coors_buffer, sc.coors_buffer.shape[0], sc.n_dev, sc.shape
    if planned:
        counts = ops.read_ints(torch.cat([sc.n_dev for sc in planned]))       # the one sync
        for m, sc, c in zip(convs, planned, counts):
            sc.set_n(c)
            x._planned[id(m)] = sc
    return planned
