"""Dense BEV stage on NHWC rows.

A BEV map ``[B,C,H,W]`` is stored as rows ``[B*H*W, C]`` (row = (b*H + y)*W + x, channel fastest).  In that
layout ``Conv2d`` / ``ConvTranspose2d`` are the same gather-GEMM as the sparse convolution, driven by a
regular-grid neighbour table, so they run on the tcgen05 kernel with BatchNorm / bias / GELU / ReLU /
residual fused in the epilogue and ``torch.cat`` replaced by writing into channel slices of a wider buffer.

Replaces the ``torch.nn`` (cuDNN) modules of det3d/models/necks/rpn.py:186-259,300-337 and
det3d/models/bbox_heads/center_head.py:209-244.  Everything here is a thin wrapper over
``s2d_conv_fwd`` / ``s2d_grid2d_*`` / ``s2d_dwconv2d`` / ``s2d_layernorm_chw`` (include/s2d_b200.h).
"""
import ctypes

import torch

from . import _lib, ops

ACT_NONE, ACT_RELU, ACT_GELU = 0, 1, 2

_TABLES = {}          # (device, kind, B, H, W, ...) -> tensors; regular-grid tables depend on the shape only


def _stream():
    return torch.cuda.current_stream().cuda_stream


def conv_table(device, B, H, W, k, stride, pad):
    """-> (tbl i32 [k*k, B*Ho*Wo], Ho, Wo) for Conv2d(k, stride, pad) on a [B,H,W] grid (cached)."""
    key = (str(device), "conv", B, H, W, k, stride, pad)
    hit = _TABLES.get(key)
    if hit is None:
        Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        n = B * Ho * Wo
        tbl = ops.alloc_table(k * k, n, device)
        _lib.check(_lib.load().s2d_grid2d_table(B, H, W, k, k, stride, pad, tbl.data_ptr(), tbl.stride(0), _stream()),
                   "s2d_grid2d_table")
        hit = _TABLES[key] = (tbl, Ho, Wo)
    return hit


def tconv_tables(device, B, H, W, k, pad, stride=2):
    """The stride**2 sub-pixel classes of ConvTranspose2d(k, stride, pad) with k - stride == 2*pad:
    list of (py, px, tbl, out_rows) (cached)."""
    key = (str(device), "tconv", B, H, W, k, pad, stride)
    hit = _TABLES.get(key)
    if hit is None:
        n = B * H * W
        hit = []
        for py in range(stride):
            for px in range(stride):
                tbl = ops.alloc_table((k // stride) ** 2, n, device)
                rows = torch.empty((n,), dtype=torch.int32, device=device)
                _lib.check(_lib.load().s2d_grid2d_tconv_table_s(B, H, W, k, stride, pad, py, px, tbl.data_ptr(),
                                                                tbl.stride(0), rows.data_ptr(), _stream()),
                           "s2d_grid2d_tconv_table_s")
                hit.append((py, px, tbl, rows))
        _TABLES[key] = hit
    return hit


def nearest_index(device, B, H, W, Ho, Wo, scale=None):
    """Row map of nn.Upsample(mode='nearest') from [B,H,W] to [B,Ho,Wo] (cached): src = min(floor(dst * scale), in - 1)
    with scale = in / out in float32 when a size is given, 1 / scale_factor when a factor is given (torch semantics)."""
    key = (str(device), "nearest", B, H, W, Ho, Wo, scale)
    hit = _TABLES.get(key)
    if hit is None:
        import numpy as np
        sy = np.float32(H) / np.float32(Ho) if scale is None else np.float32(1.0 / scale)
        sx = np.float32(W) / np.float32(Wo) if scale is None else np.float32(1.0 / scale)
        ys = np.minimum(np.floor(np.arange(Ho, dtype=np.float32) * sy).astype(np.int64), H - 1)
        xs = np.minimum(np.floor(np.arange(Wo, dtype=np.float32) * sx).astype(np.int64), W - 1)
        idx = (np.arange(B)[:, None, None] * H + ys[None, :, None]) * W + xs[None, None, :]
        hit = _TABLES[key] = torch.from_numpy(idx.reshape(-1).astype(np.int32)).to(device)
    return hit


def to_rows(x_nchw, out=None):
    """torch NCHW -> rows [B*H*W, C] (or into `out`, a 2-D view with stride(1) == 1)."""
    B, C, H, W = x_nchw.shape
    x_nchw = x_nchw.contiguous().float()
    if out is None:
        out = torch.empty((B * H * W, C), dtype=torch.float32, device=x_nchw.device)
    _lib.check(_lib.load().s2d_nchw_to_nhwc(x_nchw.data_ptr(), B, C, H * W, out.data_ptr(), out.stride(0), _stream()),
               "s2d_nchw_to_nhwc")
    return out


def to_nchw(rows, B, H, W):
    C = rows.shape[1]
    out = torch.empty((B, C, H, W), dtype=torch.float32, device=rows.device)
    _lib.check(_lib.load().s2d_nhwc_to_nchw(rows.data_ptr(), rows.stride(0), B, C, H * W, out.data_ptr(), _stream()),
               "s2d_nhwc_to_nchw")
    return out


def fold_bn(bn, bias, cout, device):
    """eval-mode BatchNorm2d (+ conv bias) -> (scale, shift); bn None -> (None, bias)."""
    if bn is None:
        return None, (None if bias is None else bias.detach().float().contiguous())
    inv = torch.rsqrt(bn.running_var.float() + bn.eps)
    gamma = bn.weight.float() if bn.weight is not None else torch.ones(cout, device=device)
    beta = bn.bias.float() if bn.bias is not None else torch.zeros(cout, device=device)
    scale = gamma * inv
    shift = beta - bn.running_mean.float() * scale
    if bias is not None:
        shift = shift + bias.float() * scale
    return scale.detach().contiguous(), shift.detach().contiguous()


def rows_with_twin(n, c, device):
    """A row buffer that several launches fill by column slices (``torch.cat`` by offset) together with its split-row twin:
    every BF16-pair launch that writes a slice of the buffer also writes the same slice of the twin (``conv_rows`` finds it
    through the view's base), and ``commit_twin`` publishes the twin once all slices are in, so the consumer's
    ``ops.rows_split`` is a cache hit instead of a pass over the whole concatenation (1.2 GB of traffic for the 512-channel
    188 x 188 maps of a batch of 8)."""
    buf = torch.empty((n, c), dtype=torch.float32, device=device)
    buf._s2d_twin = torch.empty((n, c), dtype=torch.int32, device=device)
    buf._s2d_twin_state = [0, True]                      # [channels x launches written with a twin, no launch without one]
    return buf


def commit_twin(buf, expected):
    """Publish the twin of ``rows_with_twin`` when ``expected`` twin-writing launches happened and none wrote fp32 only."""
    st = getattr(buf, "_s2d_twin_state", None)
    if st is not None and st[1] and st[0] == expected:
        ops.set_split(buf, buf._s2d_twin)


def _twin_slice(out, cout):
    """The slice of the base buffer's twin that corresponds to the row view ``out`` (or None)."""
    base = out._base if out._base is not None else out
    twin = getattr(base, "_s2d_twin", None)
    if twin is None or base.dim() != 2:
        return None, None
    off = (out.data_ptr() - base.data_ptr()) // 4
    r0, c0 = divmod(off, base.stride(0))
    if c0 % 32 or cout % 32 or out.stride(0) != base.stride(0):
        return None, base
    return twin[r0:r0 + out.shape[0], c0:c0 + cout], base


GRID_TMA = int(__import__("os").environ.get("S2D_GRID_TMA", "1"))     # 0: table-driven gathers for the regular grids too (A-B)


def conv_rows(x, weight_kio, tbl, n_out, scale=None, shift=None, act=ACT_NONE, residual=None, res_after_act=False,
              out=None, out_rows=None, precision=ops.PRECISION_TF32X3, packed=None, out_split=None, want_split=True,
              grid=None, tile_masks=None):
    """One gather-GEMM launch.  x / out / residual: 2-D views with stride(1) == 1; weight_kio: [K,Cin,Cout].
    ``packed``: a weight image for ``ops.effective_precision(precision, ...)`` or a dict precision -> image.
    With the BF16-pair kernel the input is read in split-row form (the producer's twin when ``x`` carries one, else
    converted here) and a freshly allocated result gets its own twin (``want_split``) for the next layer."""
    K, cin, cout = weight_kio.shape
    assert x.stride(1) == 1 and x.shape[1] == cin and tbl.shape[0] == K
    fresh = out is None
    if out is None:
        out = torch.empty((n_out if out_rows is None else int(out_rows.numel()), cout), dtype=torch.float32,
                          device=x.device)
    assert out.stride(1) == 1 and out.shape[1] == cout
    prec = ops.effective_precision(precision, cin, cout, tbl)
    w = weight_kio
    if prec != ops.PRECISION_FP32:
        if isinstance(packed, dict):
            w = packed.get(prec)
            if w is None:
                w = packed[prec] = ops.pack_weights_tf32(weight_kio, prec)
        else:
            w = packed if packed is not None else ops.pack_weights_tf32(weight_kio, prec)
    xs = None
    twin_base = None
    if not fresh and out_split is None:
        tw, twin_base = _twin_slice(out, cout)
        if prec == ops.PRECISION_BF16X2 and tw is not None:
            out_split = tw
            twin_base._s2d_twin_state[0] += 1
        elif twin_base is not None:
            twin_base._s2d_twin_state[1] = False           # an fp32-only write: the twin must not be published
    if prec == ops.PRECISION_BF16X2:
        xs = ops.rows_split(x, cache=True)       # the twin is reused by the weight gradient / the next consumer of x
        if out_split is None and fresh and want_split and (cout % 32 == 0 or cout == 16):
            out_split = torch.empty((out.shape[0], cout), dtype=torch.int32, device=x.device)
    else:
        out_split = None
    if grid is not None and GRID_TMA and prec == ops.PRECISION_BF16X2 and out_rows is None and cin % 32 == 0 and \
            x.shape[0] == grid[0] * grid[1] * grid[2]:
        # stride-1 Conv2d on a regular map: TMA boxes instead of the neighbour table (s2d_conv_fwd_grid)
        rows, n_tile_rows = ops.grid_tile_rows(x.device, grid[0], grid[1], grid[2])
        ops.conv_launch(x, w, None, n_tile_rows, cin, cout, K, scale, shift, act, residual, res_after_act, out, rows, prec, xs,
                        out_split, grid=grid)
    else:
        assert tile_masks is None or prec == ops.PRECISION_BF16X2, "tile masks need the BF16-pair kernel"
        ops.conv_launch(x, w, tbl, n_out, cin, cout, K, scale, shift, act, residual, res_after_act, out, out_rows, prec, xs,
                        out_split, tile_masks)
    if out_split is not None and fresh:
        ops.set_split(out, out_split)
    return out


class ParamCache:
    """Per-module cache of derived tensors (k-major / packed weights, folded BN), rebuilt when a source
    parameter changes (data_ptr / version / device)."""

    def __init__(self):
        self._store = {}

    def get(self, name, sources, build):
        key = tuple(None if t is None else (t.data_ptr(), t._version, str(t.device)) for t in sources)
        hit = self._store.get(name)
        if hit is None or hit[0] != key:
            hit = (key, build())
            self._store[name] = hit
        return hit[1]


def _bn_sources(bn):
    return [] if bn is None else [bn.weight, bn.bias, bn.running_mean, bn.running_var]


class DenseOps:
    """Executes torch ``Conv2d`` / ``ConvTranspose2d`` (+BN +activation) modules on NHWC rows."""

    def __init__(self, precision=ops.PRECISION_TF32X3):
        self.precision = precision
        self.cache = ParamCache()
        self.training = False         # set by the owning module: batch-statistics BN + autograd (autograd.py)

    # -- training mode ---------------------------------------------------------------------
    @staticmethod
    def _train_table(key, tbl, n_in, n_out):
        from . import autograd as AG
        hit = _TABLES.get(("train",) + key)
        if hit is None:
            hit = _TABLES[("train",) + key] = AG.Table(tbl, n_in, n_out)
        return hit

    @staticmethod
    def _train_norm(y, bias, bn, act, residual, res_after_act):
        from . import autograd as AG
        if bn is None:
            return AG.norm_act(y, None, bias, act, residual, res_after_act)
        return AG.norm_act(y, bn, None, act, residual, res_after_act, pre_bias=bias)

    def _conv_train(self, x, B, H, W, conv, bn, act, residual, res_after_act, out, pad):
        from . import autograd as AG
        k, s = conv.kernel_size[0], conv.stride[0]
        tbl, Ho, Wo = conv_table(x.device, B, H, W, k, s, pad)
        table = self._train_table((str(x.device), "conv", B, H, W, k, s, pad), tbl, B * H * W, B * Ho * Wo)
        w = conv.weight
        kio = w.permute(2, 3, 1, 0).reshape(k * k, w.shape[1], w.shape[0])
        y = AG.GatherConv.apply(x, kio, table, self.precision)
        y = self._train_norm(y, conv.bias, bn, act, residual, res_after_act)
        if out is not None:
            out.copy_(y)
        return y, Ho, Wo

    def _tconv_train(self, x, B, H, W, conv, bn, act, out):
        from . import autograd as AG
        k, pad, st = conv.kernel_size[0], conv.padding[0], conv.stride[0]
        classes = []
        for py, px, tbl, rows in tconv_tables(x.device, B, H, W, k, pad, st):
            ky0, kx0 = (py + pad) % st, (px + pad) % st
            taps = [(ky0 + st * a, kx0 + st * c) for a in range(k // st) for c in range(k // st)]
            classes.append((taps, tbl, rows))
        adj_tbl, Hc, Wc = conv_table(x.device, B, st * H, st * W, k, st, pad)
        assert (Hc, Wc) == (H, W)
        adj = self._train_table((str(x.device), "conv", B, st * H, st * W, k, st, pad), adj_tbl, B * st * H * st * W,
                                B * H * W)
        y = AG.TransposedConv.apply(x, conv.weight, classes, adj, self.precision)
        y = self._train_norm(y, conv.bias, bn, act, None, False)
        if out is not None:
            out.copy_(y)
        return y, st * H, st * W

    # -- parameter preparation -------------------------------------------------------------
    def _conv_weights(self, name, conv, transposed_class=None):
        def build():
            w = conv.weight.detach().float()
            if transposed_class is None:                      # Conv2d [Cout,Cin,kh,kw] -> [K,Cin,Cout]
                kio = w.permute(2, 3, 1, 0).reshape(-1, w.shape[1], w.shape[0]).contiguous()
            else:                                             # ConvTranspose2d [Cin,Cout,kh,kw], taps of one class
                py, px, pad, st = transposed_class
                kh, kw = w.shape[2], w.shape[3]
                ky0, kx0 = (py + pad) % st, (px + pad) % st
                taps = [w[:, :, ky0 + st * a, kx0 + st * c] for a in range(kh // st) for c in range(kw // st)]
                kio = torch.stack(taps, 0).contiguous()       # [K, Cin, Cout]
            return kio, {}                                    # effective precision -> packed image (filled by conv_rows)
        return self.cache.get(("w", name, transposed_class, self.precision), [conv.weight], build)

    def _affine(self, name, conv, bn):
        return self.cache.get(("bn", name), [conv.bias] + _bn_sources(bn),
                              lambda: fold_bn(bn, conv.bias, conv.weight.shape[0] if not isinstance(
                                  conv, torch.nn.ConvTranspose2d) else conv.weight.shape[1], conv.weight.device))

    # -- layers ----------------------------------------------------------------------------
    def conv(self, name, x, B, H, W, conv, bn=None, act=ACT_NONE, residual=None, res_after_act=False, out=None,
             pad=None):
        """Conv2d on rows; returns (out_rows_tensor, Ho, Wo)."""
        k, s = conv.kernel_size[0], conv.stride[0]
        pad = conv.padding[0] if pad is None else pad
        if self.training:
            return self._conv_train(x, B, H, W, conv, bn, act, residual, res_after_act, out, pad)
        tbl, Ho, Wo = conv_table(x.device, B, H, W, k, s, pad)
        kio, packed = self._conv_weights(name, conv)
        scale, shift = self._affine(name, conv, bn)
        # stride-1 3x3 layers: TMA boxes instead of the neighbour table (3-7 % faster: no table reads, one issuing thread);
        # 1x1 layers stay on the table path (the 16 x 8-pixel tiles pad the map edges: 4 % more rows, measured 5-10 % slower)
        grid = (B, H, W, k, pad) if (s == 1 and k == 3 and pad == 1 and conv.dilation[0] == 1) else None
        y = conv_rows(x, kio, tbl, B * Ho * Wo, scale, shift, act, residual, res_after_act, out, None,
                      self.precision, packed, grid=grid)
        return y, Ho, Wo

    def tconv(self, name, x, B, H, W, conv, bn=None, act=ACT_NONE, out=None):
        """ConvTranspose2d(k, stride s, pad) with k - s == 2*pad on rows as s*s sub-pixel convolutions;
        returns (out, s*H, s*W)."""
        k, pad, st = conv.kernel_size[0], conv.padding[0], conv.stride[0]
        assert k % st == 0 and k - st == 2 * pad, "ConvTranspose2d shape outside the sub-pixel scheme"
        if self.training:
            return self._tconv_train(x, B, H, W, conv, bn, act, out)
        cout = conv.weight.shape[1]
        if out is None:
            out = torch.empty((B * st * st * H * W, cout), dtype=torch.float32, device=x.device)
        scale, shift = self._affine(name, conv, bn)
        cin = conv.weight.shape[0]
        split = None
        for py, px, tbl, rows in tconv_tables(x.device, B, H, W, k, pad, st):
            kio, packed = self._conv_weights(name, conv, (py, px, pad, st))
            if split is None and out.is_contiguous() and cout % 32 == 0 and \
                    ops.effective_precision(self.precision, cin, cout, tbl) == ops.PRECISION_BF16X2:
                split = torch.empty(out.shape, dtype=torch.int32, device=x.device)   # one twin for the s*s classes
            conv_rows(x, kio, tbl, B * H * W, scale, shift, act, None, False, out, rows, self.precision, packed,
                      out_split=split)
        if split is not None:
            ops.set_split(out, split)
        return out, st * H, st * W

    def maxpool2(self, x, B, H, W):
        """nn.MaxPool2d(2, 2) on rows -> (out, H//2, W//2)."""
        C = x.shape[1]
        out = torch.empty((B * (H // 2) * (W // 2), C), dtype=torch.float32, device=x.device)
        _lib.check(_lib.load().s2d_maxpool2d_rows(x.data_ptr(), x.stride(0), B, H, W, C, out.data_ptr(), out.stride(0),
                                                  _stream()), "s2d_maxpool2d_rows")
        return out, H // 2, W // 2

    def upsample_nearest(self, x, B, H, W, Ho, Wo, scale=None, out=None):
        """nn.Upsample(mode='nearest') on rows (into `out`, a 2-D view with stride(1) == 1, when given)."""
        C = x.shape[1]
        idx = nearest_index(x.device, B, H, W, Ho, Wo, scale)
        if out is None:
            out = torch.empty((B * Ho * Wo, C), dtype=torch.float32, device=x.device)
        _lib.check(_lib.load().s2d_gather_rows(x.data_ptr(), x.stride(0), idx.data_ptr(), B * Ho * Wo, C, out.data_ptr(),
                                               out.stride(0), _stream()), "s2d_gather_rows")
        return out

    def dwconv(self, x, B, H, W, conv):
        C, k = conv.weight.shape[0], conv.kernel_size[0]
        if self.training:
            from . import autograd as AG
            assert conv.groups == C
            return AG.DepthwiseConv.apply(x, conv.weight, conv.bias, B, H, W, conv.padding[0])
        assert conv.groups == C and x.is_contiguous()
        out = torch.empty_like(x)
        w = conv.weight.detach().float().contiguous()
        b = None if conv.bias is None else conv.bias.detach().float().contiguous()
        _lib.check(_lib.load().s2d_dwconv2d(x.data_ptr(), w.data_ptr(), None if b is None else b.data_ptr(), B, H, W, C,
                                            k, conv.padding[0], out.data_ptr(), _stream()), "s2d_dwconv2d")
        return out

    def layernorm(self, x, B, H, W, ln):
        C = x.shape[1]
        if self.training:
            from . import autograd as AG
            assert tuple(ln.normalized_shape) == (C, H, W)
            return AG.LayerNormCHW.apply(x, ln.weight, ln.bias, B, ln.eps)
        assert tuple(ln.normalized_shape) == (C, H, W) and x.is_contiguous()
        out = torch.empty_like(x)
        lib = _lib.load()
        nbytes = lib.s2d_layernorm_workspace_bytes(B)
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=x.device)
        g = None if ln.weight is None else ln.weight.detach().float().contiguous()
        b = None if ln.bias is None else ln.bias.detach().float().contiguous()
        _lib.check(lib.s2d_layernorm_chw(x.data_ptr(), None if g is None else g.data_ptr(),
                                         None if b is None else b.data_ptr(), B, C, H * W, float(ln.eps),
                                         out.data_ptr(), ws.data_ptr(), nbytes, _stream()), "s2d_layernorm_chw")
        return out
