"""GPU parity of the BF16-pair (pre-split operand) tcgen05 convolution, conv_bf2.cu, through the C ABI vs the
fp64-accumulating oracle.  Tolerance 2e-5 of the output scale per layer (north_star bar: 1e-3 end to end)."""
import numpy as np
import pytest
import torch

from oracle import ref_ops as R
from sparse2dense_b200 import ops

from test_gpu_spconv_tc import SHAPES, make_case, rel_err

pytestmark = pytest.mark.gpu
TOL = 2e-5


def split_ref(x):
    """numpy restatement of the split-row format: per 32-channel chunk (16 for C = 16) [hi words | lo words]."""
    t = torch.from_numpy(np.ascontiguousarray(x))
    hi = t.to(torch.bfloat16)
    lo = (t - hi.float()).to(torch.bfloat16)
    n, c = x.shape
    ch = 16 if c == 16 else 32
    h16 = hi.view(torch.int16).numpy().astype(np.uint16).reshape(n, c // ch, ch // 2, 2)
    l16 = lo.view(torch.int16).numpy().astype(np.uint16).reshape(n, c // ch, ch // 2, 2)
    hw = h16[..., 0].astype(np.uint32) | (h16[..., 1].astype(np.uint32) << 16)
    lw = l16[..., 0].astype(np.uint32) | (l16[..., 1].astype(np.uint32) << 16)
    return np.concatenate([hw, lw], -1).reshape(n, c).view(np.int32)


@pytest.mark.parametrize("c", [16, 32, 64, 256])
def test_rows_split_bit_exact(c):
    rng = np.random.default_rng(c)
    x = (rng.normal(size=(1037, c)) * np.exp(rng.normal(size=(1037, c)) * 4)).astype(np.float32)
    x[3, 1] = 0.0
    got = ops.rows_split(torch.from_numpy(x).cuda()).cpu().numpy()
    np.testing.assert_array_equal(got, split_ref(x))


@pytest.mark.parametrize("cin,cout", SHAPES + [(256, 256), (32, 16), (128, 64)])
@pytest.mark.parametrize("fused", [False, True])
@pytest.mark.parametrize("masks", [False, True])
def test_bf2_subm_vs_oracle(cin, cout, fused, masks):
    n = 3001                                                  # ragged: 23 full tiles + 57 rows
    shape, batch, coors, feats, w = make_case(cin, cout, n, cin * 7 + cout)
    rt, _ = R.rulebook_subm(coors, shape, 3)
    c = torch.from_numpy(coors).cuda()
    tbl = ops.rulebook_subm(c, ops.build_grid_index(c, batch, shape), 3)
    ref = R.spconv_fwd(feats, w, rt, wide=True)
    kw = {}
    if fused:
        rng = np.random.default_rng(1)
        scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
        shift = rng.normal(0, 0.2, cout).astype(np.float32)
        res = rng.normal(size=(n, cout)).astype(np.float32)
        ref = R.bn_act(ref, scale, shift, res, True)
        kw = dict(scale=torch.from_numpy(scale).cuda(), shift=torch.from_numpy(shift).cuda(),
                  residual=torch.from_numpy(res).cuda(), relu=True)
    if masks:
        kw["tile_masks"] = ops.table_tile_masks(tbl, n)
        live = (rt >= 0)
        exp = np.array([sum(int(live[k, t * 128:(t + 1) * 128].any()) << k for k in range(27))
                        for t in range((n + 127) // 128)], dtype=np.int64)
        np.testing.assert_array_equal(kw["tile_masks"].cpu().numpy().astype(np.int64), exp)
    out = ops.spconv_fwd(torch.from_numpy(feats).cuda(), torch.from_numpy(w).cuda(), tbl, n,
                         precision=ops.PRECISION_BF16X2, **kw)
    torch.cuda.synchronize()
    assert rel_err(out.cpu().numpy(), ref) < TOL
    sp = ops.get_split(out)
    if cout % 32 == 0 or cout == 16:
        assert sp is not None
        np.testing.assert_array_equal(sp.cpu().numpy(), split_ref(out.cpu().numpy()))


def test_bf2_sparse_rows_skip_offsets():
    """Rows ordered so that whole tiles have no neighbour at most offsets: the skipped steps must not change anything."""
    cin = cout = 32
    n = 2000
    rng = np.random.default_rng(0)
    # isolated voxels (stride 3 lattice): only the centre offset is live for every tile
    lin = rng.choice(13 * 13 * 3, n // 4, replace=False)
    iso = np.stack([np.zeros_like(lin), 3 * (lin // 169), 3 * ((lin // 13) % 13), 3 * (lin % 13)], 1)
    shape, batch, coors, feats, w = make_case(cin, cout, n, 3)
    coors = np.concatenate([iso.astype(np.int32) + np.array([1, 0, 0, 0], np.int32), coors[coors[:, 0] == 0]], 0)
    n = len(coors)
    feats = rng.normal(size=(n, cin)).astype(np.float32)
    rt, _ = R.rulebook_subm(coors, shape, 3)
    c = torch.from_numpy(coors).cuda()
    tbl = ops.rulebook_subm(c, ops.build_grid_index(c, batch, shape), 3)
    masks = ops.table_tile_masks(tbl, n)
    assert int(masks[0].item()) == 1 << 13                    # first tiles: centre offset only
    ref = R.spconv_fwd(feats, w, rt, wide=True)
    out = ops.spconv_fwd(torch.from_numpy(feats).cuda(), torch.from_numpy(w).cuda(), tbl, n,
                         precision=ops.PRECISION_BF16X2, tile_masks=masks)
    assert rel_err(out.cpu().numpy(), ref) < TOL


def test_bf2_strided_and_k3_vs_oracle():
    for (cin, cout, ks, st, pd) in [(16, 32, 3, 2, 1), (32, 64, 3, 2, 1), (64, 128, 3, 2, (0, 1, 1)),
                                    (128, 128, (3, 1, 1), (2, 1, 1), 0)]:
        kst = (ks,) * 3 if isinstance(ks, int) else ks
        shape, batch, coors, feats, w = make_case(cin, cout, 2500, 5, kst)
        oc, rt, oshape, _ = R.rulebook_sparse(coors, shape, ks, st, pd)
        c = torch.from_numpy(coors).cuda()
        idx = ops.build_grid_index(c, batch, shape)
        sc = ops.sparse_out_coords(c, len(coors), batch, shape, ks, st, pd)
        tbl = ops.rulebook_sparse(sc.coors, idx, ks, st, pd)
        ref = R.spconv_fwd(feats, w, rt, wide=True)
        out = ops.spconv_fwd(torch.from_numpy(feats).cuda(), torch.from_numpy(w).cuda(), tbl, len(oc),
                             precision=ops.PRECISION_BF16X2, tile_masks=ops.table_tile_masks(tbl, len(oc)))
        assert rel_err(out.cpu().numpy(), ref) < TOL, (cin, cout)


def test_bf2_chain_uses_producer_split_and_is_deterministic():
    shape, batch, coors, feats, w = make_case(64, 64, 4096, 9)
    c = torch.from_numpy(coors).cuda()
    tbl = ops.rulebook_subm(c, ops.build_grid_index(c, batch, shape), 3)
    f, wt = torch.from_numpy(feats).cuda(), torch.from_numpy(w).cuda()
    a = ops.spconv_fwd(f, wt, tbl, 4096, precision=ops.PRECISION_BF16X2, relu=True)
    assert ops.get_split(a) is not None
    b2 = ops.spconv_fwd(a, wt, tbl, 4096, precision=ops.PRECISION_BF16X2)          # gathers a's split twin
    a_plain = a.clone()                                                              # no twin: converted by rows_split
    b3 = ops.spconv_fwd(a_plain, wt, tbl, 4096, precision=ops.PRECISION_BF16X2)
    assert torch.equal(b2, b3)
    s = ops.spconv_fwd(a, wt, tbl, 4096, precision=ops.PRECISION_FP32)
    assert rel_err(b2.cpu().numpy(), s.cpu().numpy()) < TOL
    a.add_(1.0)                                                                      # stale twin must not be used
    assert ops.get_split(a) is None


def test_grouped_rows_are_a_permutation_and_give_bit_identical_results():
    """s2d_table_group_rows: perm is a permutation, tbl_out[k][p] == tbl[k][perm[p]], the grouped launch (out_rows = perm)
    writes exactly the bits of the plain launch -- with and without residual -- and the grouped tile masks skip more."""
    from sparse2dense_b200 import synth
    from sparse2dense_b200.hotpath import concat_clouds
    pts, offs = concat_clouds([synth.lidar_scene(21), synth.small_scene(22)])
    vb = ops.voxelize(pts.cuda(), offs, synth.WAYMO_VOXEL, synth.WAYMO_RANGE, 5, 150000, want_voxels=False, mean_channels=5)
    n = vb.n
    coors = vb.coors_buffer[:n]
    index = ops.build_grid_index(coors, 2, (41, 1504, 1504))
    tbl = ops.rulebook_subm(coors, index, 3)
    gt, perm, gmasks = ops.table_group_rows(tbl, n)
    assert torch.equal(torch.sort(perm.long())[0], torch.arange(n, device="cuda"))
    assert torch.equal(gt[:, :n], tbl[:, :n][:, perm.long()])
    masks = ops.table_tile_masks(tbl, n)
    live = lambda m: sum(bin(int(v) & 0x7ffffff).count("1") for v in m.cpu().tolist())
    assert live(gmasks) < 0.6 * live(masks), (live(gmasks), live(masks))
    torch.manual_seed(3)
    for c in (16, 32, 64):
        x = torch.relu(torch.randn(n, c, device="cuda"))
        w = torch.randn(27, c, c, device="cuda") / (27 * c) ** 0.5
        res = torch.randn(n, c, device="cuda")
        sc, sh = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda")
        for r in (None, res):
            a = ops.spconv_fwd(x, w, tbl, n, sc, sh, r, True, ops.PRECISION_BF16X2, tile_masks=masks)
            b = ops.spconv_fwd(x, w, gt, n, sc, sh, r, True, ops.PRECISION_BF16X2, tile_masks=gmasks, out_rows=perm)
            assert torch.equal(a, b), c
            assert torch.equal(ops.get_split(a), ops.get_split(b)), c


def test_grouped_submanifold_builder_equals_group_rows_of_the_scan_order_table():
    """s2d_rulebook_subm_grouped (key from the occupancy bitmap, table written in grouped order) == s2d_rulebook_subm followed
    by s2d_table_group_rows, bit for bit: permutation, table and tile masks."""
    from sparse2dense_b200 import synth
    from sparse2dense_b200.hotpath import concat_clouds
    pts, offs = concat_clouds([synth.lidar_scene(31), synth.small_scene(32), synth.small_scene(33)])
    vb = ops.voxelize(pts.cuda(), offs, synth.WAYMO_VOXEL, synth.WAYMO_RANGE, 5, 150000, want_voxels=False, mean_channels=5)
    n = vb.n
    coors = vb.coors_buffer[:n]
    index = ops.build_grid_index(coors, 3, (41, 1504, 1504))
    a_tbl, a_perm, a_masks = ops.table_group_rows(ops.rulebook_subm(coors, index, 3), n)
    b_tbl, b_perm, b_masks = ops.rulebook_subm_grouped(coors, index)
    assert torch.equal(a_perm[:n], b_perm[:n]) and torch.equal(a_tbl[:, :n], b_tbl[:, :n]) and torch.equal(a_masks, b_masks)


def test_grouped_strided_builder_equals_group_rows_of_the_scan_order_table():
    """s2d_rulebook_sparse_grouped == s2d_rulebook_sparse + s2d_table_group_rows (stride 2 pad 1, and pad (0,1,1))."""
    from sparse2dense_b200 import synth
    from sparse2dense_b200.hotpath import concat_clouds
    pts, offs = concat_clouds([synth.lidar_scene(41), synth.small_scene(42)])
    vb = ops.voxelize(pts.cuda(), offs, synth.WAYMO_VOXEL, synth.WAYMO_RANGE, 5, 150000, want_voxels=False, mean_channels=5)
    n = vb.n
    coors, shape = vb.coors_buffer[:n], (41, 1504, 1504)
    index = ops.build_grid_index(coors, 2, shape)
    for pad in (1, (0, 1, 1)):
        sc = ops.sparse_out_coords(coors, n, 2, shape, 3, 2, pad)
        oc, m = sc.coors, sc.coors.shape[0]
        a_tbl, a_perm, a_masks = ops.table_group_rows(ops.rulebook_sparse(oc, index, 3, 2, pad), m)
        b_tbl, b_perm, b_masks = ops.rulebook_sparse_grouped(oc, index, 2, pad)
        assert torch.equal(a_perm[:m], b_perm[:m]) and torch.equal(a_tbl[:, :m], b_tbl[:, :m]) and torch.equal(a_masks, b_masks)


def _set_variant(v):
    import ctypes
    from sparse2dense_b200 import _lib
    _lib.load()
    ctypes.CDLL(_lib.LIB_PATH).s2d_debug_bf2_variant(int(v))


@pytest.mark.parametrize("cin,cout,n", [(128, 128, 562), (64, 128, 1300), (128, 256, 3001)])
def test_swapped_operand_kernel_variants_agree(cin, cout, n):
    """128 output channels run with swapped operands and two alternating issuer warps (conv_bf2.cu).  On a grouped rulebook
    whose tiles differ in their live offsets (so that one tile of a group starts accumulating before the other), with an ODD
    number of tiles (562 rows = 4 tiles + 50 rows: the last group has one tile) and with two output-channel blocks: the result
    matches the fp64 oracle, is bit-identical between the dynamic tile-group scheduler and the static round robin (variant 4),
    and agrees with the unswapped kernel (variant 2) to the per-layer tolerance."""
    shape, batch, coors, feats, w = make_case(cin, cout, n, 11 * cin + cout + n)
    rt, _ = R.rulebook_subm(coors, shape, 3)
    ref = R.spconv_fwd(feats, w, rt, wide=True)
    c = torch.from_numpy(coors).cuda()
    tbl = ops.rulebook_subm(c, ops.build_grid_index(c, batch, shape), 3)
    gt, perm, masks = ops.table_group_rows(tbl, n)
    assert len(set(int(m) for m in masks.cpu().tolist())) > 1            # tiles with different live offsets
    x, wt = torch.from_numpy(feats).cuda(), torch.from_numpy(w).cuda()
    pk = ops.pack_weights_tf32(wt, ops.PRECISION_BF16X2)
    outs = {}
    try:
        for v in (0, 4, 2):
            _set_variant(v)
            outs[v] = ops.spconv_fwd(x, wt, gt, n, precision=ops.PRECISION_BF16X2, packed=pk, tile_masks=masks,
                                     out_rows=perm).cpu().numpy()
    finally:
        _set_variant(0)
    assert rel_err(outs[0], ref) < TOL
    np.testing.assert_array_equal(outs[0], outs[4])
    assert rel_err(outs[2], outs[0]) < TOL
