"""Training-mode operators: ``torch.autograd.Function`` wrappers whose forward AND backward run in libs2d_b200.so.

torch's autograd engine is used as the tape only (which op follows which, gradient accumulation of fan-outs); every
convolution, normalisation and activation gradient is one of the kernels of ``csrc/train.cu`` or the forward gather-GEMM
itself run over a transposed table.  Reference semantics: spconv ``indice_conv`` backward, ``nn.Conv2d`` /
``nn.ConvTranspose2d`` / ``nn.BatchNorm1d|2d`` (train mode, det3d/models/utils/norm.py:59-108) / ``nn.GELU`` / ``nn.ReLU`` /
``nn.LayerNorm`` autograd as exercised by ``TS_Trainer.batch_processor_inline`` (det3d/torchie/trainer/trainer.py:775-811).
"""
import torch

from . import _lib, ops
from .dense import ACT_NONE, conv_rows


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


class Table:
    """A k-major neighbour table with its lazily built transpose (for the data gradient)."""

    def __init__(self, tbl, n_in, n_out, symmetric=False, out_rows=None, grouped=None):
        self.tbl, self.n_in, self.n_out, self.symmetric, self.out_rows = tbl, int(n_in), int(n_out), symmetric, out_rows
        self._inv = None
        # (table in grouped row order, perm, tile masks) of a 27-offset sparse rulebook (conv_bf2.cu "Row grouping"): the
        # forward -- and for a submanifold table, which is its own transpose with the taps reversed, the data gradient --
        # skip the (tile, offset) blocks without neighbours, with bit-identical results.  The scan-order ``tbl`` stays for
        # the weight gradient.
        self.grouped = grouped

    def transposed(self):
        """inv[k][j] = i  <=>  tbl[k][i] = j.  A SubM table is its own transpose with the taps reversed."""
        if self._inv is None:
            if self.symmetric:
                self._inv = torch.flip(self.tbl, dims=[0])
            else:
                K = self.tbl.shape[0]
                inv = torch.empty((K, max(self.n_in, 1)), dtype=torch.int32, device=self.tbl.device)
                _lib.check(_lib.load().s2d_table_transpose(self.tbl.data_ptr(), self.tbl.stride(0), K, self.n_out,
                                                           _ptr(self.out_rows), inv.data_ptr(), inv.stride(0), self.n_in,
                                                           _stream()), "s2d_table_transpose")
                self._inv = inv
        return self._inv


WGRAD_TC = int(__import__("os").environ.get("S2D_WGRAD_TC", "1"))     # 0: CUDA-core weight gradient everywhere (debug / A-B)


def conv_wgrad(g, d, tbl, n_rows, d_rows=None, precision=ops.PRECISION_AUTO):
    """out[k][a][b] = sum_i g[tbl[k][i]][a] * d[d_rows[i] or i][b]  ->  f32 [K, Cg, Cd].  Channel counts that are multiples of 32
    run on the tensor cores (wgrad_tc.cu) from the split-row twins of both operands unless ``precision`` is FP32."""
    assert g.stride(1) == 1 and d.stride(1) == 1
    K, cg, cd = tbl.shape[0], g.shape[1], d.shape[1]
    lib = _lib.load()
    out = torch.empty((K, cg, cd), dtype=torch.float32, device=g.device)
    if WGRAD_TC and precision != ops.PRECISION_FP32 and n_rows >= 256 and lib.s2d_conv_wgrad_bf2_supported(cg, cd):
        gs, ds = ops.rows_split(g, cache=True), ops.rows_split(d, cache=True)
        nbytes = lib.s2d_conv_wgrad_bf2_workspace_bytes(n_rows, K, cg, cd)
        ws = torch.empty((max(nbytes, 16),), dtype=torch.uint8, device=g.device)
        _lib.check(lib.s2d_conv_wgrad_bf2(gs.data_ptr(), gs.stride(0), gs.shape[0], cg, ds.data_ptr(), ds.stride(0), _ptr(d_rows),
                                          cd, tbl.data_ptr(), tbl.stride(0), n_rows, K, out.data_ptr(), 0, ws.data_ptr(), nbytes,
                                          _stream()), "s2d_conv_wgrad_bf2")
        return out
    nbytes = lib.s2d_conv_wgrad_workspace_bytes(n_rows, K, cg, cd)
    ws = torch.empty((max(nbytes, 16),), dtype=torch.uint8, device=g.device)
    _lib.check(lib.s2d_conv_wgrad(g.data_ptr(), g.stride(0), g.shape[0], cg, d.data_ptr(), d.stride(0), _ptr(d_rows), cd,
                                  tbl.data_ptr(), tbl.stride(0), n_rows, K, out.data_ptr(), 0, ws.data_ptr(), nbytes,
                                  _stream()), "s2d_conv_wgrad")
    return out


_MAX_K = 27          # taps per s2d_conv_fwd launch (kMaxK of the library)


def _pad_channels(c, is_input):
    """Smallest channel count >= c the tcgen05 kernel takes (Cin: 16 or a multiple of 32; Cout: a multiple of 16)."""
    if is_input:
        return 16 if c <= 16 else -(-c // 32) * 32
    return -(-c // 16) * 16


def conv_rows_tc(x, w, tbl, n_out, precision, out_rows=None, n_total=None, grouped=None):
    """conv_rows that keeps small-channel layers (the 3-channel generators of the PCR branch, the 1..3-channel head
    convolutions and their data gradients) on the tensor-core kernel by zero-padding the channel counts, when there is
    enough work per row to pay for the padded traffic; otherwise the exact-shape kernel.  With ``out_rows`` the result
    is scattered into a fresh ``[n_total, Cout(_padded)]`` buffer.  Returns a 2-D tensor whose first Cout columns hold
    the result (possibly a strided view of the padded buffer)."""
    K, cin, cout = w.shape
    tc = precision != ops.PRECISION_FP32
    if tc and not ops.tf32_supported(cin, cout) and K * cin >= 64:
        cin_p, cout_p = _pad_channels(cin, True), _pad_channels(cout, False)
        if ops.tf32_supported(cin_p, cout_p) and cin_p <= 8 * cin and cout_p <= 16 * cout:
            if cin_p != cin:
                xp = torch.zeros((x.shape[0], cin_p), dtype=torch.float32, device=x.device)
                xp[:, :cin] = x
                x = xp
            wp = torch.zeros((K, cin_p, cout_p), dtype=torch.float32, device=w.device)
            wp[:, :cin, :cout] = w
            w = wp
    cout_eff = w.shape[2]
    tile_masks = None
    if grouped is not None and out_rows is None and \
            ops.effective_precision(precision, x.shape[1], cout_eff, grouped[0]) == ops.PRECISION_BF16X2:
        # ``grouped`` = (tbl in grouped row order, perm, tile masks) of the same rulebook: the BF16-pair kernel skips the
        # (tile, offset) blocks without neighbours and scatters through perm -- bit-identical to the plain launch
        tbl, out_rows, tile_masks = grouped
        n_total = n_out
    out = None
    if out_rows is not None:
        out = torch.empty((n_total, cout_eff), dtype=torch.float32, device=x.device)
    y = conv_rows(x, w.contiguous(), tbl, n_out, out=out, out_rows=out_rows, precision=precision, tile_masks=tile_masks)
    return y if cout_eff == cout else y[:, :cout]


def conv_rows_any_k(x, w, tbl, n_out, precision):
    """conv_rows for any number of taps: more than 27 (the 4x4x4 adjoint of ConvTranspose3d) run as successive launches
    that accumulate through the residual input."""
    K = tbl.shape[0]
    if K <= _MAX_K:
        return conv_rows_tc(x, w, tbl, n_out, precision)
    cin, cout = w.shape[1], w.shape[2]
    if precision != ops.PRECISION_FP32 and not ops.tf32_supported(cin, cout):
        cin_p, cout_p = _pad_channels(cin, True), _pad_channels(cout, False)
        if ops.tf32_supported(cin_p, cout_p) and cin_p <= 8 * cin and cout_p <= 16 * cout:     # pad once for all chunks
            xp = torch.zeros((x.shape[0], cin_p), dtype=torch.float32, device=x.device)
            xp[:, :cin] = x
            wp = torch.zeros((K, cin_p, cout_p), dtype=torch.float32, device=w.device)
            wp[:, :cin, :cout] = w
            x, w = xp, wp
    out = None
    for k0 in range(0, K, _MAX_K):
        k1 = min(K, k0 + _MAX_K)
        out = conv_rows(x, w[k0:k1].contiguous(), tbl[k0:k1], n_out, residual=out, precision=precision)
    return out if out.shape[1] == cout else out[:, :cout]


class GatherConv(torch.autograd.Function):
    """out[i] = sum_k x[tbl[k][i]] . w[k]   (x [n_in, Cin], w [K, Cin, Cout])."""

    @staticmethod
    def forward(ctx, x, w, table, precision):
        x = x if x.stride(1) == 1 else x.contiguous()
        w = w.contiguous()
        ctx.save_for_backward(x, w)
        ctx.table, ctx.precision = table, precision
        return conv_rows_tc(x, w, table.tbl, table.n_out, precision, grouped=table.grouped)

    @staticmethod
    def backward(ctx, dy):
        x, w = ctx.saved_tensors
        t = ctx.table
        dy = dy.contiguous()
        dx = dw = None
        if ctx.needs_input_grad[0]:
            wt = w.transpose(1, 2).contiguous()                       # [K, Cout, Cin]
            if t.symmetric:
                wt = torch.flip(wt, dims=[0])
                dx = conv_rows_tc(dy, wt, t.tbl, t.n_in, ctx.precision, grouped=t.grouped)   # same table, taps of W reversed
            else:
                dx = conv_rows_tc(dy, wt, t.transposed(), t.n_in, ctx.precision)
        if ctx.needs_input_grad[1]:
            dw = conv_wgrad(x, dy, t.tbl, t.n_out, precision=ctx.precision)
        return dx, dw, None, None


class TransposedConv(torch.autograd.Function):
    """ConvTranspose2d(k, stride s, pad) with k - s == 2 pad on rows: forward as s*s sub-pixel gather-GEMMs, backward
    through the adjoint Conv2d table on the fine grid (data gradient = that convolution, weight gradient = the same
    row-reduction GEMM with the roles of input and output swapped)."""

    @staticmethod
    def forward(ctx, x, w_t, classes, adj_table, precision):
        # x [n_coarse, Cin]; w_t torch layout [Cin, Cout, *kernel] (2-D or 3-D); classes: [(taps, tbl, out_rows)] with
        # taps = kernel index tuples in the order of the class table's rows
        x = x if x.stride(1) == 1 else x.contiguous()
        cin, cout = w_t.shape[:2]
        n_fine = adj_table.n_in
        K0 = len(classes[0][0])
        pad = (precision != ops.PRECISION_FP32 and not ops.tf32_supported(cin, cout) and K0 * cin >= 64 and
               ops.tf32_supported(cin, _pad_channels(cout, False)))
        cout_p = _pad_channels(cout, False) if pad else cout       # small Cout (generator_2: 16 -> 3): padded, tensor cores
        out = torch.empty((n_fine, cout_p), dtype=torch.float32, device=x.device)
        for taps, tbl, rows in classes:
            kio = torch.stack([w_t[(slice(None), slice(None)) + tuple(t)] for t in taps], 0)
            if cout_p != cout:
                kp = torch.zeros((kio.shape[0], cin, cout_p), dtype=torch.float32, device=x.device)
                kp[:, :, :cout] = kio
                kio = kp
            conv_rows(x, kio.contiguous(), tbl, x.shape[0], out=out, out_rows=rows, precision=precision)
        ctx.save_for_backward(x, w_t)
        ctx.adj, ctx.precision = adj_table, precision
        return out if cout_p == cout else out[:, :cout]

    @staticmethod
    def backward(ctx, dy):
        x, w_t = ctx.saved_tensors
        adj = ctx.adj                                                  # fine grid (n_in) -> coarse grid (n_out)
        cin, cout = w_t.shape[:2]
        dy = dy.contiguous()
        dx = dw = None
        if ctx.needs_input_grad[0]:
            kio = w_t.flatten(2).permute(2, 1, 0).contiguous()         # [K, Cout, Cin], K row-major over the kernel dims
            dx = conv_rows_any_k(dy, kio, adj.tbl, adj.n_out, ctx.precision)
        if ctx.needs_input_grad[1]:
            g = conv_wgrad(dy, x, adj.tbl, adj.n_out, precision=ctx.precision)                  # [K, Cout, Cin]
            dw = g.permute(2, 1, 0).reshape(w_t.shape).contiguous()
        return dx, dw, None, None, None


def _rows_ws(C, device):
    nbytes = _lib.load().s2d_rows_workspace_bytes(C)
    return torch.empty((nbytes,), dtype=torch.uint8, device=device), nbytes


def _all_reduce_sums(ws, C, n, group):
    """[2*C column sums, row count] of this rank (float64) summed over the ranks: ONE small all-reduce, no host sync."""
    import torch.distributed as dist
    off = _lib.load().s2d_rows_workspace_sums_offset(C)
    buf = torch.empty((2 * C + 1,), dtype=torch.float64, device=ws.device)
    buf[:2 * C].copy_(ws[off:off + 16 * C].view(torch.float64))
    buf[2 * C:].fill_(float(n))
    dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=None if group is True else group)
    return buf


class RowsNormAct(torch.autograd.Function):
    """y = act(norm(x) (+ res))  or  act(norm(x)) + res  on rows [n, C].
    ``bn`` = (running_mean, running_var, momentum, eps, group): training-mode BatchNorm with affine (gamma, beta);
    ``group`` None = statistics of this rank, True / a process group = SyncBatchNorm over that group;
    ``bn`` = None: plain bias (gamma ignored, beta = bias or None)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, residual, bn, act, res_after_act):
        x = x.contiguous()
        n, C = x.shape
        lib = _lib.load()
        residual = None if residual is None else (residual if residual.stride(1) == 1 else residual.contiguous())
        out = torch.empty_like(x)
        group = None
        if bn is not None:
            running_mean, running_var, momentum, eps, group = bn
            stats = torch.empty((4, C), dtype=torch.float32, device=x.device)      # mean, invstd, scale, shift
            ws, nbytes = _rows_ws(C, x.device)
            if group is None:
                _lib.check(lib.s2d_bn_train_stats(x.data_ptr(), x.stride(0), n, C, float(eps), float(momentum), _ptr(gamma),
                                                  _ptr(beta), _ptr(running_mean), _ptr(running_var), stats[0].data_ptr(),
                                                  stats[1].data_ptr(), stats[2].data_ptr(), stats[3].data_ptr(),
                                                  ws.data_ptr(), nbytes, _stream()), "s2d_bn_train_stats")
            else:                                     # SyncBatchNorm: statistics over the rows of every rank
                _lib.check(lib.s2d_bn_train_sums(x.data_ptr(), x.stride(0), n, C, ws.data_ptr(), nbytes, _stream()),
                           "s2d_bn_train_sums")
                glob = _all_reduce_sums(ws, C, n, group)
                _lib.check(lib.s2d_bn_train_finalize(glob.data_ptr(), glob[2 * C:].data_ptr(), C, float(eps),
                                                     float(momentum), _ptr(gamma), _ptr(beta), _ptr(running_mean),
                                                     _ptr(running_var), stats[0].data_ptr(), stats[1].data_ptr(),
                                                     stats[2].data_ptr(), stats[3].data_ptr(), _stream()),
                           "s2d_bn_train_finalize")
            scale, shift = stats[2], stats[3]
        else:
            stats, scale = None, None
            shift = None if beta is None else beta.detach().contiguous()
        _lib.check(lib.s2d_rows_affine_act(x.data_ptr(), x.stride(0), n, C, _ptr(scale), _ptr(shift), _ptr(residual),
                                           0 if residual is None else residual.stride(0), act, int(res_after_act),
                                           out.data_ptr(), out.stride(0), _stream()), "s2d_rows_affine_act")
        ctx.save_for_backward(x, gamma, beta, residual, stats)
        ctx.is_bn, ctx.act, ctx.res_after_act, ctx.group = bn is not None, act, bool(res_after_act), group
        return out

    @staticmethod
    def backward(ctx, dy):
        x, gamma, beta, residual, stats = ctx.saved_tensors
        n, C = x.shape
        lib = _lib.load()
        dy = dy if dy.stride(1) == 1 else dy.contiguous()
        dz = torch.empty_like(x)
        ws, nbytes = _rows_ws(C, x.device)
        if ctx.is_bn:
            scale, shift = stats[2], stats[3]
        else:
            scale, shift = None, (None if beta is None else beta.detach().contiguous())
        dbias = None
        if not ctx.is_bn and beta is not None and ctx.needs_input_grad[2]:
            dbias = torch.empty((C,), dtype=torch.float32, device=x.device)
        _lib.check(lib.s2d_rows_affine_act_bwd(x.data_ptr(), x.stride(0), n, C, _ptr(scale), _ptr(shift), _ptr(residual),
                                               0 if residual is None else residual.stride(0), ctx.act,
                                               int(ctx.res_after_act), dy.data_ptr(), dy.stride(0), dz.data_ptr(),
                                               dz.stride(0), _ptr(dbias), ws.data_ptr(), nbytes, _stream()),
                   "s2d_rows_affine_act_bwd")
        dres = None
        if residual is not None and ctx.needs_input_grad[3]:
            dres = dy if ctx.res_after_act else dz
        if not ctx.is_bn:
            return dz, None, dbias, dres, None, None, None
        dx = torch.empty_like(x)
        dgamma = torch.empty((C,), dtype=torch.float32, device=x.device)
        dbeta = torch.empty((C,), dtype=torch.float32, device=x.device)
        if ctx.group is None:
            _lib.check(lib.s2d_bn_train_bwd(x.data_ptr(), x.stride(0), n, C, dz.data_ptr(), dz.stride(0),
                                            stats[0].data_ptr(), stats[1].data_ptr(), _ptr(gamma), dx.data_ptr(),
                                            dx.stride(0), dgamma.data_ptr(), dbeta.data_ptr(), ws.data_ptr(), nbytes,
                                            _stream()), "s2d_bn_train_bwd")
        else:
            off = lib.s2d_rows_workspace_sums_offset(C)
            local = ws[off:off + 16 * C].view(torch.float64)
            _lib.check(lib.s2d_bn_train_bwd_params(local.data_ptr(), C, stats[0].data_ptr(), stats[1].data_ptr(),
                                                   dgamma.data_ptr(), dbeta.data_ptr(), _stream()), "s2d_bn_train_bwd_params")
            glob = _all_reduce_sums(ws, C, n, ctx.group)
            _lib.check(lib.s2d_bn_train_bwd_dx(x.data_ptr(), x.stride(0), n, C, dz.data_ptr(), dz.stride(0),
                                               stats[0].data_ptr(), stats[1].data_ptr(), _ptr(gamma), glob.data_ptr(),
                                               glob[2 * C:].data_ptr(), dx.data_ptr(), dx.stride(0), ws.data_ptr(), nbytes,
                                               _stream()), "s2d_bn_train_bwd_dx")
        return dx, (dgamma if gamma is not None else None), (dbeta if beta is not None else None), dres, None, None, None


class _ZeroGradBias(torch.autograd.Function):
    """A convolution bias in front of a training-mode BatchNorm cancels in the normalisation: its gradient is exactly
    zero.  This node hands the parameter that zero so that optimizers see a gradient as they do in the reference."""

    @staticmethod
    def forward(ctx, y, bias):
        return y.view_as(y)

    @staticmethod
    def backward(ctx, dy):
        return dy, torch.zeros(dy.shape[1], dtype=dy.dtype, device=dy.device)


def norm_act(x, norm, bias=None, act=ACT_NONE, residual=None, res_after_act=False, pre_bias=None):
    """Apply a torch BatchNorm module in TRAINING mode (batch statistics, running-stat update) or a plain ``bias``,
    then the activation / residual.  ``pre_bias``: the bias of the convolution in front of the BatchNorm; it only
    shifts the running mean."""
    if norm is None:
        return RowsNormAct.apply(x, None, bias, residual, None, act, res_after_act)
    assert isinstance(norm, torch.nn.modules.batchnorm._BatchNorm)
    if not norm.training:
        raise NotImplementedError("frozen (eval-mode) BatchNorm inside a training graph is not built")
    if norm.momentum is None:
        raise NotImplementedError("BatchNorm(momentum=None) (cumulative average) is not used by the reference configs")
    momentum = norm.momentum
    if norm.num_batches_tracked is not None:
        norm.num_batches_tracked += 1
    if pre_bias is not None and pre_bias.requires_grad:
        x = _ZeroGradBias.apply(x, pre_bias)
    group = None
    if isinstance(norm, torch.nn.SyncBatchNorm):                  # nn.SyncBatchNorm.convert_sync_batchnorm(model)
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(norm.process_group) > 1:
            group = True if norm.process_group is None else norm.process_group
    y = RowsNormAct.apply(x, norm.weight, norm.bias, residual,
                          (norm.running_mean, norm.running_var, momentum, norm.eps, group), act, res_after_act)
    if pre_bias is not None:
        with torch.no_grad():
            norm.running_mean.add_(pre_bias.detach(), alpha=momentum)
    # the statistics kernels updated the running buffers through raw pointers: bump their version counters so that the
    # folded-BN caches keyed on (data_ptr, _version) are rebuilt by the next eval-mode forward
    torch.autograd.graph.increment_version([norm.running_mean, norm.running_var])
    return y


class LayerNormCHW(torch.autograd.Function):
    """nn.LayerNorm([C, H, W]) on rows x [B*HW, C] (weight / bias [C, H, W])."""

    @staticmethod
    def forward(ctx, x, weight, bias, B, eps):
        x = x.contiguous()
        C, HW = x.shape[1], x.shape[0] // B
        lib = _lib.load()
        out = torch.empty_like(x)
        nbytes = lib.s2d_layernorm_workspace_bytes(B)
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=x.device)
        w = None if weight is None else weight.detach().contiguous()
        b = None if bias is None else bias.detach().contiguous()
        _lib.check(lib.s2d_layernorm_chw(x.data_ptr(), _ptr(w), _ptr(b), B, C, HW, float(eps), out.data_ptr(),
                                         ws.data_ptr(), nbytes, _stream()), "s2d_layernorm_chw")
        ctx.save_for_backward(x, weight)
        ctx.B, ctx.eps, ctx.has_bias = B, eps, bias is not None
        return out

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        B, C, HW = ctx.B, x.shape[1], x.shape[0] // ctx.B
        lib = _lib.load()
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        dw = torch.empty((C, HW), dtype=torch.float32, device=x.device)
        db = torch.empty((C, HW), dtype=torch.float32, device=x.device)
        nbytes = lib.s2d_layernorm_bwd_workspace_bytes(B)
        ws = torch.empty((nbytes,), dtype=torch.uint8, device=x.device)
        w = None if weight is None else weight.detach().contiguous()
        _lib.check(lib.s2d_layernorm_chw_bwd(x.data_ptr(), _ptr(w), B, C, HW, float(ctx.eps), dy.data_ptr(), dx.data_ptr(),
                                             dw.data_ptr(), db.data_ptr(), ws.data_ptr(), nbytes, _stream()),
                   "s2d_layernorm_chw_bwd")
        return (dx, None if weight is None else dw.view_as(weight), db.view_as(weight) if ctx.has_bias else None,
                None, None)


class DepthwiseConv(torch.autograd.Function):
    """Depthwise k x k Conv2d (groups = C, stride 1) on rows x [B*H*W, C]; weight [C, 1, k, k]."""

    @staticmethod
    def forward(ctx, x, weight, bias, B, H, W, pad):
        x = x.contiguous()
        C, k = weight.shape[0], weight.shape[2]
        out = torch.empty_like(x)
        w = weight.detach().contiguous()
        _lib.check(_lib.load().s2d_dwconv2d(x.data_ptr(), w.data_ptr(), _ptr(None if bias is None else bias.detach()),
                                            B, H, W, C, k, pad, out.data_ptr(), _stream()), "s2d_dwconv2d")
        ctx.save_for_backward(x, weight)
        ctx.dims, ctx.has_bias = (B, H, W, pad), bias is not None
        return out

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        B, H, W, pad = ctx.dims
        C, k = weight.shape[0], weight.shape[2]
        lib = _lib.load()
        dy = dy.contiguous()
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            assert 2 * pad == k - 1
            wf = torch.flip(weight.detach(), dims=[2, 3]).contiguous()
            dx = torch.empty_like(x)
            _lib.check(lib.s2d_dwconv2d(dy.data_ptr(), wf.data_ptr(), None, B, H, W, C, k, pad, dx.data_ptr(), _stream()),
                       "s2d_dwconv2d")
        if ctx.needs_input_grad[1]:
            nbytes = lib.s2d_dwconv2d_wgrad_workspace_bytes(B, H, W, C, k)
            ws = torch.empty((nbytes,), dtype=torch.uint8, device=x.device)
            dw = torch.empty_like(weight)
            _lib.check(lib.s2d_dwconv2d_wgrad(x.data_ptr(), dy.data_ptr(), B, H, W, C, k, pad, dw.data_ptr(), ws.data_ptr(),
                                              nbytes, _stream()), "s2d_dwconv2d_wgrad")
        if ctx.has_bias and ctx.needs_input_grad[2]:
            # column sums of dy: the bias-only mode of the rows backward (act = none, dz = dy)
            wsb, nb = _rows_ws(C, x.device)
            db = torch.empty((C,), dtype=torch.float32, device=x.device)
            scratch = torch.empty_like(dy)
            _lib.check(lib.s2d_rows_affine_act_bwd(dy.data_ptr(), dy.stride(0), dy.shape[0], C, None, None, None, 0,
                                                   ACT_NONE, 0, dy.data_ptr(), dy.stride(0), scratch.data_ptr(),
                                                   scratch.stride(0), db.data_ptr(), wsb.data_ptr(), nb, _stream()),
                       "s2d_rows_affine_act_bwd")
        return dx, dw, db, None, None, None, None


class DenseBEV(torch.autograd.Function):
    """``SparseConvTensor.dense()`` + ``view(N, C*D, H, W)`` (scn.py:173-176) as NHWC rows or NCHW; the backward is the
    gather of the active cells (pure data movement, torch indexing)."""

    @staticmethod
    def forward(ctx, feats, coors, batch, spatial_shape, as_rows):
        d, h, w = (int(v) for v in spatial_shape)
        ctx.save_for_backward(coors)
        ctx.dims, ctx.as_rows = (batch, feats.shape[1], d, h, w), as_rows
        if as_rows:
            return ops.dense_bev_rows(feats, coors, batch, spatial_shape)
        return ops.dense_bev(feats, coors, batch, spatial_shape).view(batch, feats.shape[1] * d, h, w)

    @staticmethod
    def backward(ctx, dy):
        (coors,) = ctx.saved_tensors
        B, C, D, H, W = ctx.dims
        b, z, y, x = (coors[:, i].long() for i in range(4))
        if ctx.as_rows:
            g = dy.reshape(B * H * W, C, D)[(b * H + y) * W + x, :, z]
        else:
            g = dy.reshape(B, C, D, H, W)[b, :, z, y, x]
        return g.contiguous(), None, None, None, None

