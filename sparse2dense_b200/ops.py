"""Tensor-level wrappers over the C ABI (``include/s2d_b200.h``).

PyTorch is plumbing here: it owns device memory and the CUDA stream; every arithmetic step is
one call into ``libs2d_b200.so``.  All functions require CUDA tensors and raise otherwise --
there is no CPU path.
"""
import numpy as np
import torch

from . import _lib

PRECISION_FP32 = 0
PRECISION_TF32 = 1
PRECISION_TF32X3 = 2
PRECISION_TF32_BF16C = 3      # TF32 main term + BF16 correction terms: TF32X3 accuracy class at 2/3 of the tensor time
PRECISION_AUTO = 4            # the library picks TF32_BF16C or TF32X3 per layer shape (both fp32-level accuracy)
PRECISION_BF16X2 = 5          # activations and weights pre-split into BF16 pairs, three BF16 MMAs (conv_bf2.cu): ~4e-6 per layer
PRECISION_NAMES = {"fp32": PRECISION_FP32, "tf32": PRECISION_TF32, "tf32x3": PRECISION_TF32X3,
                   "tf32_bf16c": PRECISION_TF32_BF16C, "auto": PRECISION_AUTO, "bf16x2": PRECISION_BF16X2}
PRECISION_NAMES_INV = {v: k for k, v in PRECISION_NAMES.items()}


# When set to a list, spconv_fwd appends (key, start_event, end_event) per launch so that bench.py can
# time the dominant kernel live on its own stream (profiling aid; off by default).
KERNEL_EVENTS = None


def _stream():
    return torch.cuda.current_stream().cuda_stream


def kernel_launches():
    """Kernels launched by libs2d_b200.so in this process so far."""
    return int(_lib.load().s2d_kernel_launches())


def _ptr(t):
    return None if t is None else t.data_ptr()


_EXPORT = {}


def read_ints(t):
    """Device int32 tensor (a few control values: row counts) -> python list, synchronising the current stream.
    The values travel through mapped pinned host memory written by a kernel (``s2d_export_i32``), not a DMA copy, so
    the read does not queue behind a bulk device->host transfer running on the copy engine."""
    _need_cuda(t)
    t = t.contiguous()
    assert t.dtype == torch.int32
    n = t.numel()
    key = (t.device.index, torch.cuda.current_stream(t.device).cuda_stream)
    slot = _EXPORT.get(key)
    if slot is None or slot[0].numel() < n:
        slot = (torch.empty((max(n, 64),), dtype=torch.int32, pin_memory=True), torch.cuda.Event())
        _EXPORT[key] = slot
    buf, ev = slot
    _lib.check(_lib.load().s2d_export_i32(_ptr(t), n, buf.data_ptr(), _stream()), "s2d_export_i32")
    ev.record()
    ev.synchronize()
    return buf[:n].tolist()


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.S2DError("sparse2dense_b200 ops need CUDA tensors (no CPU fallback)")


def _triple(v):
    if isinstance(v, (int, np.integer)):
        return (int(v),) * 3
    v = tuple(int(a) for a in v)
    assert len(v) == 3, v
    return v


# ------------------------------------------------------------------------------------------
# voxelizer
# ------------------------------------------------------------------------------------------
class VoxelBatch:
    """Device-side result of :func:`voxelize` with capacity-sized buffers and a lazy host count."""

    def __init__(self, voxels, coors, num_points, mean, voxel_offsets, batch):
        self._voxels, self._coors, self._num_points, self._mean = voxels, coors, num_points, mean
        self.voxel_offsets = voxel_offsets          # device i32 [batch+1]
        self.batch = batch
        self._offsets_host = None

    @property
    def n_dev(self):
        return self.voxel_offsets[self.batch:]

    def offsets_host(self):
        if self._offsets_host is None:
            self._offsets_host = read_ints(self.voxel_offsets)       # the one host sync
        return self._offsets_host

    @property
    def n(self):
        return self.offsets_host()[-1]

    @property
    def capacity(self):
        return self._coors.shape[0]

    @property
    def coors_buffer(self):
        return self._coors

    @property
    def mean_buffer(self):
        return self._mean

    @property
    def voxels(self):
        return None if self._voxels is None else self._voxels[: self.n]

    @property
    def coors(self):
        return self._coors[: self.n]

    @property
    def num_points(self):
        return self._num_points[: self.n]

    @property
    def mean(self):
        return None if self._mean is None else self._mean[: self.n]


def voxelize(points, scene_offsets, voxel_size, coors_range, max_points, max_voxels, want_voxels=True,
             mean_channels=None):
    """Batched GPU voxelizer (C ABI ``s2d_voxelize``; reference point_cloud_ops.py:112-184).

    points: cuda f32 [N,F] (scenes concatenated); scene_offsets: host ints [B+1].
    """
    _need_cuda(points)
    lib = _lib.load()
    assert points.dtype == torch.float32 and points.dim() == 2
    points = points.contiguous()
    n, f = points.shape
    batch = len(scene_offsets) - 1
    dev = points.device
    cap = min(batch * max_voxels, max(n, 1))
    voxels = torch.empty((cap, max_points, f), dtype=torch.float32, device=dev) if want_voxels else None
    coors = torch.empty((cap, 4), dtype=torch.int32, device=dev)
    num = torch.empty((cap,), dtype=torch.int32, device=dev)
    mean = torch.empty((cap, mean_channels), dtype=torch.float32, device=dev) if mean_channels else None
    offs = torch.empty((batch + 1,), dtype=torch.int32, device=dev)
    ws_bytes = lib.s2d_voxelize_workspace_bytes(n, batch, max_points, max_voxels)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    _lib.check(lib.s2d_voxelize(_ptr(points), _lib.ints(scene_offsets), n, batch, f, _lib.floats(coors_range),
                                _lib.floats(voxel_size), max_points, max_voxels, _ptr(voxels), _ptr(coors),
                                _ptr(num), _ptr(mean), mean_channels or 0, _ptr(offs), _ptr(ws), ws_bytes,
                                _stream()), "s2d_voxelize")
    return VoxelBatch(voxels, coors, num, mean, offs, batch)


def voxel_mean(voxels, num_points, channels=None):
    """VoxelFeatureExtractorV3.forward (voxel_encoder.py:17-24)."""
    _need_cuda(voxels, num_points)
    m, p, f = voxels.shape
    c = f if channels is None else channels
    out = torch.empty((m, c), dtype=torch.float32, device=voxels.device)
    _lib.check(_lib.load().s2d_voxel_mean(_ptr(voxels.contiguous()), _ptr(num_points.int().contiguous()), m, p, f, c,
                                          _ptr(out), _stream()), "s2d_voxel_mean")
    return out


# ------------------------------------------------------------------------------------------
# coordinate index + rulebooks
# ------------------------------------------------------------------------------------------
class GridIndex:
    """Occupancy bitmap + popcount prefix (+ rank->row permutation) of one sparse tensor."""

    def __init__(self, batch, shape, capacity, device):
        self.batch, self.shape, self.capacity = int(batch), _triple(shape), int(capacity)
        self.nbytes = _lib.load().s2d_grid_index_bytes(self.batch, _lib.ints(self.shape), self.capacity)
        self.buf = torch.empty((self.nbytes,), dtype=torch.uint8, device=device)


def build_grid_index(coors, batch, shape, n=None, n_dev=None):
    _need_cuda(coors)
    assert coors.dtype == torch.int32 and coors.is_contiguous() and coors.shape[1] == 4
    n = coors.shape[0] if n is None else n
    idx = GridIndex(batch, shape, n, coors.device)
    _lib.check(_lib.load().s2d_grid_index_build(_ptr(coors), n, _ptr(n_dev), idx.batch, _lib.ints(idx.shape),
                                                _ptr(idx.buf), idx.nbytes, _stream()), "s2d_grid_index_build")
    return idx


def rulebook_subm(coors, index, ksize=3, dilation=1, count_pairs=False):
    """-> tbl i32 [K, n] (k-major), and the pair count tensor (device u64 as int64) if requested."""
    _need_cuda(coors)
    n = coors.shape[0]
    ks, dl = _triple(ksize), _triple(dilation)
    k = ks[0] * ks[1] * ks[2]
    tbl = alloc_table(k, n, coors.device)
    pairs = torch.zeros((1,), dtype=torch.int64, device=coors.device) if count_pairs else None
    _lib.check(_lib.load().s2d_rulebook_subm(_ptr(coors), n, index.batch, _lib.ints(index.shape), _lib.ints(ks),
                                             _lib.ints(dl), _ptr(index.buf), _ptr(tbl), tbl.stride(0), _ptr(pairs),
                                             _stream()), "s2d_rulebook_subm")
    return (tbl, pairs) if count_pairs else tbl


def rulebook_subm_grouped(coors, index, dilation=1):
    """The 3x3x3 submanifold rulebook built directly in grouped row order (``s2d_rulebook_subm_grouped``) ->
    (tbl i32 [27, n], perm i32 [n], tile_masks): identical to ``table_group_rows(rulebook_subm(...))``."""
    _need_cuda(coors)
    n = coors.shape[0]
    lib = _lib.load()
    tbl = alloc_table(27, n, coors.device)
    perm = torch.empty((max(n, 1),), dtype=torch.int32, device=coors.device)
    masks = torch.empty((max((n + 127) // 128, 1),), dtype=torch.int32, device=coors.device)
    nbytes = lib.s2d_rulebook_subm_grouped_workspace_bytes(n)
    ws = torch.empty((max(nbytes, 16),), dtype=torch.uint8, device=coors.device)
    _lib.check(lib.s2d_rulebook_subm_grouped(_ptr(coors), n, index.batch, _lib.ints(index.shape), _lib.ints(_triple(dilation)),
                                             _ptr(index.buf), _ptr(perm), _ptr(tbl), tbl.stride(0), _ptr(masks), _ptr(ws),
                                             nbytes, _stream()), "s2d_rulebook_subm_grouped")
    return tbl, perm, masks


def rulebook_sparse_grouped(out_coors, index_in, stride, pad, dilation=1):
    """The strided 3x3x3 rulebook built directly in grouped row order -> (tbl i32 [27, n_out], perm, tile_masks)."""
    _need_cuda(out_coors)
    n = out_coors.shape[0]
    lib = _lib.load()
    tbl = alloc_table(27, n, out_coors.device)
    perm = torch.empty((max(n, 1),), dtype=torch.int32, device=out_coors.device)
    masks = torch.empty((max((n + 127) // 128, 1),), dtype=torch.int32, device=out_coors.device)
    nbytes = lib.s2d_rulebook_subm_grouped_workspace_bytes(n)
    ws = torch.empty((max(nbytes, 16),), dtype=torch.uint8, device=out_coors.device)
    _lib.check(lib.s2d_rulebook_sparse_grouped(_ptr(out_coors), n, index_in.batch, _lib.ints(index_in.shape),
                                               _lib.ints(_triple(stride)), _lib.ints(_triple(pad)), _lib.ints(_triple(dilation)),
                                               _ptr(index_in.buf), _ptr(perm), _ptr(tbl), tbl.stride(0), _ptr(masks), _ptr(ws),
                                               nbytes, _stream()), "s2d_rulebook_sparse_grouped")
    return tbl, perm, masks


def conv_out_shape(shape, ksize, stride, pad, dilation=1):
    out = (_lib.ctypes.c_int * 3)()
    _lib.check(_lib.load().s2d_conv_out_shape(_lib.ints(_triple(shape)), _lib.ints(_triple(ksize)),
                                              _lib.ints(_triple(stride)), _lib.ints(_triple(pad)),
                                              _lib.ints(_triple(dilation)), out), "s2d_conv_out_shape")
    return tuple(out)


def out_capacity_bound(n_in, batch, shape_out, ksize, stride):
    """Upper bound of the number of output sites of a strided conv over n_in active inputs."""
    per_input = 1
    for k, s in zip(_triple(ksize), _triple(stride)):
        per_input *= -(-k // s)
    return int(min(n_in * per_input, batch * shape_out[0] * shape_out[1] * shape_out[2]))


class SparseCoords:
    """Output coordinate set of a strided conv: capacity buffers, device count, lazy host count."""

    def __init__(self, coors_buf, n_dev, index, shape):
        self.coors_buffer, self.n_dev, self.index, self.shape = coors_buf, n_dev, index, shape
        self._n = None

    @property
    def n(self):
        if self._n is None:
            self._n = int(read_ints(self.n_dev)[0])
            if self._n > self.coors_buffer.shape[0]:
                raise _lib.S2DError(f"sparse conv produced {self._n} outputs > capacity {self.coors_buffer.shape[0]}")
        return self._n

    def set_n(self, n):
        self._n = int(n)
        if self._n > self.coors_buffer.shape[0]:
            raise _lib.S2DError(f"sparse conv produced {self._n} outputs > capacity {self.coors_buffer.shape[0]}")

    @property
    def coors(self):
        return self.coors_buffer[: self.n]


def sparse_out_coords(coors_in, n_in, batch, shape_in, ksize, stride, pad, dilation=1, n_in_dev=None):
    """Phase 1 of SparseConv3d: output coordinates (ascending) + their index; no host sync."""
    _need_cuda(coors_in)
    ks, st, pd, dl = _triple(ksize), _triple(stride), _triple(pad), _triple(dilation)
    shape_out = conv_out_shape(shape_in, ks, st, pd, dl)
    cap = max(out_capacity_bound(n_in, batch, shape_out, ks, st), 1)
    dev = coors_in.device
    index = GridIndex(batch, shape_out, cap, dev)
    out_coors = torch.empty((cap, 4), dtype=torch.int32, device=dev)
    n_out = torch.zeros((1,), dtype=torch.int32, device=dev)
    _lib.check(_lib.load().s2d_sparse_out_coords(_ptr(coors_in), n_in, _ptr(n_in_dev), batch,
                                                 _lib.ints(_triple(shape_in)), _lib.ints(ks), _lib.ints(st),
                                                 _lib.ints(pd), _lib.ints(dl), _ptr(index.buf), index.nbytes,
                                                 _ptr(out_coors), cap, _ptr(n_out), _stream()),
               "s2d_sparse_out_coords")
    return SparseCoords(out_coors, n_out, index, shape_out)


def rulebook_sparse(out_coors, index_in, ksize, stride, pad, dilation=1, count_pairs=False):
    """Phase 2 of SparseConv3d: gather table i32 [K, n_out]."""
    _need_cuda(out_coors)
    n_out = out_coors.shape[0]
    ks, st, pd, dl = _triple(ksize), _triple(stride), _triple(pad), _triple(dilation)
    k = ks[0] * ks[1] * ks[2]
    tbl = alloc_table(k, n_out, out_coors.device)
    pairs = torch.zeros((1,), dtype=torch.int64, device=out_coors.device) if count_pairs else None
    _lib.check(_lib.load().s2d_rulebook_sparse(_ptr(out_coors), n_out, index_in.batch, _lib.ints(index_in.shape),
                                               _lib.ints(ks), _lib.ints(st), _lib.ints(pd), _lib.ints(dl),
                                               _ptr(index_in.buf), _ptr(tbl), tbl.stride(0), _ptr(pairs), _stream()),
               "s2d_rulebook_sparse")
    return (tbl, pairs) if count_pairs else tbl


# ------------------------------------------------------------------------------------------
# convolution + densify
# ------------------------------------------------------------------------------------------
def round_up(n, m):
    return (int(n) + m - 1) // m * m


def alloc_table(k, n, device):
    """Neighbour table i32 [k, n] whose row stride is a multiple of 128 entries (16 B aligned int4 index loads)."""
    return torch.empty((k, round_up(max(n, 1), 128)), dtype=torch.int32, device=device)[:, :max(n, 1)]


def get_split(t):
    """The split-row twin of an fp32 row tensor (written by the bf16x2 conv epilogue or rows_split), if still valid."""
    hit = getattr(t, "_s2d_split", None)
    if hit is None or hit[1] != t._version or hit[0].shape != t.shape:
        return None
    return hit[0]


def set_split(t, split):
    t._s2d_split = (split, t._version)


def rows_split(x, cache=False):
    """fp32 rows [n, C] (stride(1) == 1) -> split rows int32 [n, C]: per 32-channel chunk (or 16-channel row)
    [hi words | lo words], two BF16 per word (x = hi + lo to 16 mantissa bits) -- the operand format of PRECISION_BF16X2."""
    _need_cuda(x)
    hit = get_split(x)               # written by the epilogue of the layer that produced x
    if hit is not None:
        return hit
    assert x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
    n, c = x.shape
    out = torch.empty((n, c), dtype=torch.int32, device=x.device)
    _lib.check(_lib.load().s2d_rows_split(_ptr(x), n, c, x.stride(0), _ptr(out), out.stride(0), _stream()), "s2d_rows_split")
    if cache:
        set_split(x, out)
    return out


def table_tile_masks(tbl, n_rows):
    """i32 [ceil(n_rows/128)]: bit k set iff some row of the 128-row tile has a neighbour at kernel offset k."""
    _need_cuda(tbl)
    masks = torch.empty((max((n_rows + 127) // 128, 1),), dtype=torch.int32, device=tbl.device)
    _lib.check(_lib.load().s2d_table_tile_masks(_ptr(tbl), tbl.stride(0), tbl.shape[0], n_rows, _ptr(masks), _stream()),
               "s2d_table_tile_masks")
    return masks


def table_group_rows(tbl, n_rows):
    """Group the rows of a neighbour table by their live (dz, dy) offset triples (``s2d_table_group_rows``) ->
    (tbl_out i32 [K, n_rows], perm i32 [n_rows], tile_masks): launch the conv with tbl_out, out_rows=perm, tile_masks."""
    _need_cuda(tbl)
    k = tbl.shape[0]
    lib = _lib.load()
    out = alloc_table(k, n_rows, tbl.device)
    perm = torch.empty((max(n_rows, 1),), dtype=torch.int32, device=tbl.device)
    masks = torch.empty((max((n_rows + 127) // 128, 1),), dtype=torch.int32, device=tbl.device)
    nbytes = lib.s2d_table_group_rows_workspace_bytes(n_rows)
    ws = torch.empty((max(nbytes, 16),), dtype=torch.uint8, device=tbl.device)
    _lib.check(lib.s2d_table_group_rows(_ptr(tbl), tbl.stride(0), k, n_rows, _ptr(perm), _ptr(out), out.stride(0), _ptr(masks),
                                        _ptr(ws), nbytes, _stream()), "s2d_table_group_rows")
    return out, perm, masks


def bf2_shape_ok(cin, cout, k):
    """Shape part of ``bf2_ok`` (every table from ``alloc_table`` satisfies the layout part)."""
    nchunk = 1 if cin == 16 else cin // 32
    return (cin == 16 or (cin >= 32 and cin % 32 == 0)) and cout % 16 == 0 and k <= 27 and nchunk <= 128 and \
        -(-k // (2 if cin == 16 else 1)) * nchunk <= 1024


def bf2_ok(cin, cout, tbl):
    return bf2_shape_ok(cin, cout, tbl.shape[0]) and tbl.stride(0) % 4 == 0 and tbl.data_ptr() % 16 == 0


def effective_precision(precision, cin, cout, tbl):
    """What a conv launch actually runs: AUTO -> BF16X2 (conv_bf2.cu) where that kernel exists for the shape and table
    layout, else the TF32-based AUTO of spconv_tc.cu; shapes without any tensor-core kernel -> FP32 (CUDA cores)."""
    if precision in (PRECISION_AUTO, PRECISION_BF16X2):
        if bf2_ok(cin, cout, tbl):
            return PRECISION_BF16X2
        precision = PRECISION_AUTO
    if precision != PRECISION_FP32 and not tf32_supported(cin, cout):
        return PRECISION_FP32
    return precision


def conv_launch(x, w_arg, tbl, n_out, cin, cout, k, scale=None, shift=None, act=0, residual=None, res_after_act=False,
                out=None, out_rows=None, precision=PRECISION_FP32, x_split=None, out_split=None, tile_masks=None,
                kind="dense", grid=None):
    """One ``s2d_conv_fwd`` call.  x / out / residual / x_split / out_split: 2-D row views with stride(1) == 1 (or None).
    ``grid = (B, H, W, k, pad)``: the dense-grid form ``s2d_conv_fwd_grid`` (TMA boxes instead of a neighbour table;
    ``tbl`` is None, ``n_out`` / ``out_rows`` come from ``grid_tile_rows``)."""
    p = _lib.ConvParams()
    p.in_, p.weights, p.tbl = _ptr(x), _ptr(w_arg), _ptr(tbl)
    p.scale, p.shift, p.residual = _ptr(scale), _ptr(shift), _ptr(residual)
    p.out, p.out_rows = _ptr(out), _ptr(out_rows)
    p.in_ld = 0 if x is None else x.stride(0)
    p.out_ld = 0 if out is None else out.stride(0)
    p.res_ld = 0 if residual is None else residual.stride(0)
    p.tbl_stride, p.K = (0 if tbl is None else tbl.stride(0)), k
    p.n_in = (x if x is not None else x_split).shape[0]
    p.n_out, p.Cin, p.Cout = n_out, cin, cout
    p.act, p.res_after_act, p.precision = int(act), int(bool(res_after_act)), int(precision)
    p.in_split, p.out_split, p.tile_masks = _ptr(x_split), _ptr(out_split), _ptr(tile_masks)
    p.in_split_ld = 0 if x_split is None else x_split.stride(0)
    p.out_split_ld = 0 if out_split is None else out_split.stride(0)
    ev = None
    if KERNEL_EVENTS is not None:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    if grid is not None:
        B, H, W, ksz, pad = grid
        _lib.check(_lib.load().s2d_conv_fwd_grid(_lib.ctypes.byref(p), B, H, W, ksz, pad, _stream()), "s2d_conv_fwd_grid")
    else:
        _lib.check(_lib.load().s2d_conv_fwd(_lib.ctypes.byref(p), _stream()), "s2d_conv_fwd")
    if ev is not None:
        ev[1].record()
        n_rows = n_out if grid is None else grid[0] * grid[1] * grid[2]       # real output rows (tile order pads the edges)
        KERNEL_EVENTS.append(((cin, cout, k, residual is not None, p.n_in, n_rows, int(precision), kind), ev[0], ev[1]))


_GRID_ROWS = {}


def grid_tile_rows(device, B, H, W):
    """(out_rows i32 [n_tiles * 128], n_tile_rows) of the dense-grid conv: tile-order row -> pixel row, -1 outside the map
    (cached per shape and device)."""
    key = (str(device), B, H, W)
    hit = _GRID_ROWS.get(key)
    if hit is None:
        lib = _lib.load()
        n = lib.s2d_grid2d_tile_rows_count(B, H, W)
        rows = torch.empty((n,), dtype=torch.int32, device=device)
        _lib.check(lib.s2d_grid2d_tile_rows(B, H, W, _ptr(rows), _stream()), "s2d_grid2d_tile_rows")
        hit = _GRID_ROWS[key] = (rows, n)
    return hit


def spconv_fwd(feats, weight, tbl, n_out, scale=None, shift=None, residual=None, relu=False,
               precision=PRECISION_FP32, out=None, packed=None, tile_masks=None, want_split=True, out_rows=None):
    """out[o] = act((sum_k feats[tbl[k][o]] @ W[k]) * scale + shift (+ residual[o])).

    weight: [kD,kH,kW,Cin,Cout] (spconv layout) or [K,Cin,Cout], fp32 contiguous.
    With PRECISION_BF16X2 the input is read in split-row form (the producer's ``out_split`` when ``feats`` carries one,
    else converted here) and, if ``want_split``, the result carries its own split twin for the next layer.
    """
    _need_cuda(feats, weight, tbl)
    if precision == PRECISION_BF16X2:
        cin, cout = weight.shape[-2], weight.shape[-1]
        k = weight.numel() // (cin * cout)
        assert feats.shape[1] == cin and tbl.shape[0] == k and tbl.dtype == torch.int32
        if not bf2_ok(cin, cout, tbl):
            raise _lib.S2DError(f"no bf16x2 kernel for Cin={cin} Cout={cout} K={k} tbl stride {tbl.stride(0)}")
        xs = rows_split(feats, cache=True)
        if out is None:
            out = torch.empty((n_out, cout), dtype=torch.float32, device=feats.device)
        split_ok = want_split and (cout % 32 == 0 or cout == 16) and out.is_contiguous()
        out_split = torch.empty((n_out, cout), dtype=torch.int32, device=feats.device) if split_ok else None
        w_arg = packed if packed is not None else pack_weights_tf32(weight, precision)
        conv_launch(None, w_arg, tbl, n_out, cin, cout, k, scale, shift, 1 if relu else 0, residual, False, out, out_rows,
                    precision, xs, out_split, tile_masks, kind="sparse")
        if out_split is not None:
            set_split(out, out_split)
        return out
    _need_cuda(feats, weight, tbl)
    assert out_rows is None, "out_rows needs the bf16x2 kernel"
    assert feats.dtype == torch.float32 and feats.is_contiguous() and weight.is_contiguous()
    cin, cout = weight.shape[-2], weight.shape[-1]
    k = weight.numel() // (cin * cout)
    assert feats.shape[1] == cin and tbl.shape[0] == k and tbl.dtype == torch.int32
    if out is None:
        out = torch.empty((n_out, cout), dtype=torch.float32, device=feats.device)
    for t in (scale, shift, residual):
        assert t is None or (t.dtype == torch.float32 and t.is_contiguous())
    w_arg = weight
    if precision != PRECISION_FP32:
        w_arg = packed if packed is not None else pack_weights_tf32(weight, precision)
    ev = None
    if KERNEL_EVENTS is not None:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    _lib.check(_lib.load().s2d_spconv_fwd(_ptr(feats), feats.shape[0], _ptr(w_arg), _ptr(tbl), tbl.stride(0), n_out,
                                          cin, cout, k, _ptr(scale), _ptr(shift), _ptr(residual), int(bool(relu)),
                                          _ptr(out), int(precision), _stream()), "s2d_spconv_fwd")
    if ev is not None:
        ev[1].record()
        KERNEL_EVENTS.append(((cin, cout, k, residual is not None, feats.shape[0], n_out, int(precision), "sparse"), ev[0], ev[1]))
    return out


def tf32_supported(cin, cout):
    return bool(_lib.load().s2d_spconv_tf32_supported(int(cin), int(cout)))


def pack_weights_tf32(weight, precision=PRECISION_TF32X3):
    """Weights [kD,kH,kW,Cin,Cout] -> the pre-split, pre-swizzled shared-memory image the tcgen05 kernel
    bulk-copies per contraction step (TF32 hi | TF32 lo, or TF32 hi | BF16 [w | lo] for PRECISION_TF32_BF16C).
    Done once per layer and precision."""
    _need_cuda(weight)
    weight = weight.detach().contiguous().float()
    cin, cout = weight.shape[-2], weight.shape[-1]
    k = weight.numel() // (cin * cout)
    nbytes = _lib.load().s2d_spconv_packed_bytes(k, cin, cout)
    if nbytes == 0:
        raise _lib.S2DError(f"no tcgen05 packing for Cin={cin} Cout={cout}")
    packed = torch.empty((nbytes // 4,), dtype=torch.float32, device=weight.device)
    _lib.check(_lib.load().s2d_spconv_pack_weights(_ptr(weight), k, cin, cout, int(precision), _ptr(packed), _stream()),
               "s2d_spconv_pack_weights")
    return packed


def dense_bev(feats, coors, batch, spatial_shape):
    """SparseConvTensor.dense() + view(N, C*D, H, W) (scn.py:173-176) -> [B, C*D, H, W]."""
    _need_cuda(feats, coors)
    d, h, w = _triple(spatial_shape)
    n, c = feats.shape
    bev = torch.empty((batch, c * d, h, w), dtype=torch.float32, device=feats.device)
    lib = _lib.load()
    nbytes = lib.s2d_dense_bev_workspace_bytes(batch, d, h, w)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=feats.device)
    _lib.check(lib.s2d_dense_bev_tiled(_ptr(feats.contiguous()), _ptr(coors.contiguous()), n, c, batch, d, h, w,
                                       _ptr(bev), _ptr(ws), nbytes, _stream()), "s2d_dense_bev_tiled")
    return bev


def dense_bev_rowwise(feats, coors, batch, spatial_shape):
    """The row-stationary kernel (memset + scatter); kept for comparison / as the simple reference implementation."""
    _need_cuda(feats, coors)
    d, h, w = _triple(spatial_shape)
    n, c = feats.shape
    bev = torch.empty((batch, c * d, h, w), dtype=torch.float32, device=feats.device)
    _lib.check(_lib.load().s2d_dense_bev(_ptr(feats.contiguous()), _ptr(coors.contiguous()), n, c, batch, d, h, w,
                                         _ptr(bev), _stream()), "s2d_dense_bev")
    return bev


def dense_bev_rows(feats, coors, batch, spatial_shape, out=None):
    """dense() + view(N, C*D, H, W) written directly as NHWC rows [B*H*W, C*D] (channel = c*D + z)."""
    _need_cuda(feats, coors)
    d, h, w = _triple(spatial_shape)
    n, c = feats.shape
    if out is None:
        out = torch.empty((batch * h * w, c * d), dtype=torch.float32, device=feats.device)
    _lib.check(_lib.load().s2d_dense_bev_nhwc(_ptr(feats.contiguous()), _ptr(coors.contiguous()), n, c, batch, d, h, w,
                                              _ptr(out), out.stride(0), _stream()), "s2d_dense_bev_nhwc")
    return out


# ------------------------------------------------------------------------------------------
# CenterHead.predict: decode + candidate selection + rotated NMS (center_head.py:293-495)
# ------------------------------------------------------------------------------------------
def centerhead_decode(heads, B, H, W, out_size_factor, voxel_size, pc_range, score_threshold, post_center_range):
    """heads: dict name -> 2-D row view [B*H*W, c] (stride(1) == 1) for reg/height/dim/rot/hm.
    -> boxes [B*H*W,7], scores, labels (i32), keys (u64 as int64)."""
    for k in ("reg", "height", "dim", "rot", "hm"):
        t = heads[k]
        _need_cuda(t)
        assert t.dtype == torch.float32 and t.dim() == 2 and t.stride(1) == 1 and t.shape[0] == B * H * W, k
    dev = heads["hm"].device
    n = B * H * W
    boxes = torch.empty((n, 7), dtype=torch.float32, device=dev)
    scores = torch.empty((n,), dtype=torch.float32, device=dev)
    labels = torch.empty((n,), dtype=torch.int32, device=dev)
    keys = torch.empty((n,), dtype=torch.int64, device=dev)
    p = _lib.DecodeParams()
    p.reg, p.height, p.dim, p.rot, p.hm = (heads[k].data_ptr() for k in ("reg", "height", "dim", "rot", "hm"))
    p.ld_reg, p.ld_height, p.ld_dim, p.ld_rot, p.ld_hm = (heads[k].stride(0) for k in ("reg", "height", "dim", "rot", "hm"))
    p.B, p.H, p.W, p.num_cls = B, H, W, heads["hm"].shape[1]
    p.out_size_factor, p.voxel_x, p.voxel_y = float(out_size_factor), float(voxel_size[0]), float(voxel_size[1])
    p.pc_x, p.pc_y, p.score_threshold = float(pc_range[0]), float(pc_range[1]), float(score_threshold)
    for i, v in enumerate(post_center_range):
        p.range[i] = float(v)
    _lib.check(_lib.load().s2d_centerhead_decode(_lib.ctypes.byref(p), _ptr(boxes), _ptr(scores), _ptr(labels),
                                                 _ptr(keys), _stream()), "s2d_centerhead_decode")
    return boxes, scores, labels, keys


def centerhead_select(keys, boxes, scores, labels, B, cells, pre_max, iou_threshold, post_max):
    """-> out_boxes [B,post_max,7], out_scores [B,post_max], out_labels i32, out_cells i32, n_out i32 [B] (device)."""
    _need_cuda(keys, boxes, scores, labels)
    lib = _lib.load()
    dev = boxes.device
    ob = torch.empty((B, post_max, 7), dtype=torch.float32, device=dev)
    os_ = torch.empty((B, post_max), dtype=torch.float32, device=dev)
    ol = torch.empty((B, post_max), dtype=torch.int32, device=dev)
    oc = torch.empty((B, post_max), dtype=torch.int32, device=dev)
    n_out = torch.empty((B,), dtype=torch.int32, device=dev)
    ws_bytes = lib.s2d_centerhead_select_workspace_bytes(B, pre_max, post_max)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    _lib.check(lib.s2d_centerhead_select(_ptr(keys), _ptr(boxes), _ptr(scores), _ptr(labels), B, cells, pre_max,
                                         float(iou_threshold), post_max, _ptr(ob), _ptr(os_), _ptr(ol), _ptr(oc),
                                         _ptr(n_out), _ptr(ws), ws_bytes, _stream()), "s2d_centerhead_select")
    return ob, os_, ol, oc, n_out


def nms_sorted(boxes, iou_threshold):
    """Device-side iou3d_nms_cuda.nms_gpu: boxes [n,7] sorted by descending score -> (keep i32 [n], n_keep i32 [1])."""
    _need_cuda(boxes)
    boxes = boxes.contiguous().float()
    n = boxes.shape[0]
    lib = _lib.load()
    keep = torch.empty((max(n, 1),), dtype=torch.int32, device=boxes.device)
    n_keep = torch.zeros((1,), dtype=torch.int32, device=boxes.device)
    ws_bytes = lib.s2d_nms_workspace_bytes(n)
    if ws_bytes == 0 and n > 0:
        raise _lib.S2DError(f"s2d_nms_sorted handles at most 4096 boxes, got {n}")
    ws = torch.empty((max(ws_bytes, 1),), dtype=torch.uint8, device=boxes.device)
    _lib.check(lib.s2d_nms_sorted(_ptr(boxes), n, float(iou_threshold), _ptr(keep), _ptr(n_keep), _ptr(ws), ws_bytes,
                                  _stream()), "s2d_nms_sorted")
    return keep, n_keep


def iou_bev(boxes_a, boxes_b):
    """Pairwise rotated-BEV IoU [n_a, n_b] (boxes [.,7] = x, y, z, dx, dy, dz, heading)."""
    _need_cuda(boxes_a, boxes_b)
    a, b = boxes_a.contiguous().float(), boxes_b.contiguous().float()
    out = torch.empty((a.shape[0], b.shape[0]), dtype=torch.float32, device=a.device)
    _lib.check(_lib.load().s2d_iou_bev(_ptr(a), a.shape[0], _ptr(b), b.shape[0], _ptr(out), _stream()), "s2d_iou_bev")
    return out
