#!/usr/bin/env python
"""Per-layer timing of the BF16-pair conv kernel (conv_bf2.cu) vs the TF32 kernel on the real rulebooks of a batch of
synthetic scenes.  usage: microbench_bf2.py [--batch 4] [--reps 10] [--variants 0,1,2]"""
import argparse
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparse2dense_b200 import _lib, ops, spconv, synth  # noqa: E402
from sparse2dense_b200.backbones import SpMiddleResNetFHD  # noqa: E402
from sparse2dense_b200.hotpath import concat_clouds  # noqa: E402


def timeit(fn, flush, reps):
    fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        flush.fill_(0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--reps", type=int, default=7)
    ap.add_argument("--variants", default="0,1,2")
    ap.add_argument("--dense", action="store_true", help="also time dense 3x3 / 1x1 convs of the neck shapes")
    ap.add_argument("--ablate", default="", help="comma list of ablation flag sets (see conv_bf2.cu B2Args::dbg), timed with variant 0")
    ap.add_argument("--only", type=int, default=0)
    ap.add_argument("--wgrad", action="store_true", help="also time the tensor-core weight gradient of each layer")
    ap.add_argument("--insitu", action="store_true", help="also time the grouped layer with BN affine + residual + ReLU")
    ap.add_argument("--prof", action="store_true", help="print the in-kernel cycle counters (producer / MMA waits)")
    args = ap.parse_args()
    _lib.load()
    setv = ctypes.CDLL(_lib.LIB_PATH).s2d_debug_bf2_variant
    clouds = synth.lidar_batch(1, args.batch)
    pts, offs = concat_clouds(clouds)
    vb = ops.voxelize(pts.cuda(), offs, synth.WAYMO_VOXEL, synth.WAYMO_RANGE, 5, 150000, want_voxels=False, mean_channels=5)
    n0 = vb.n
    bb = SpMiddleResNetFHD(num_input_features=5).cuda().eval()
    x = spconv.SparseConvTensor(vb.mean_buffer[:n0], vb.coors_buffer[:n0], (41, 1504, 1504), args.batch)
    plan = spconv.plan_coords(x, [bb.conv2[0], bb.conv3[0], bb.conv4[0], bb.extra_conv[0]])
    stages = [(x.indices, x.index())] + [(sc.coors, sc.index) for sc in plan]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    print(f"rows per stage: {[int(s[0].shape[0]) for s in stages]}", flush=True)
    setf = ctypes.CDLL(_lib.LIB_PATH).s2d_debug_bf2_flags
    for si, c in enumerate([16, 32, 64, 128]):
        if args.only and c != args.only:
            continue
        coors, index = stages[si]
        n = coors.shape[0]
        tbl, pairs = ops.rulebook_subm(coors, index, 3, count_pairs=True)
        p = int(pairs.item())
        masks = ops.table_tile_masks(tbl, n)
        live = float(sum(bin(int(m) & 0x7ffffff).count("1") for m in masks.cpu().tolist())) / (27 * masks.numel())
        feats = torch.relu(torch.randn(n, c, device="cuda"))
        w = torch.randn(3, 3, 3, c, c, device="cuda") / (27 * c) ** 0.5
        ref = ops.spconv_fwd(feats, w, tbl, n, precision=ops.PRECISION_TF32X3)
        old = ops.PRECISION_TF32_BF16C if c == 128 else ops.PRECISION_TF32X3
        pk_old = ops.pack_weights_tf32(w, old)
        out = torch.empty(n, c, device="cuda")
        t_old = timeit(lambda: ops.spconv_fwd(feats, w, tbl, n, precision=old, packed=pk_old, out=out), flush, args.reps)
        pk = ops.pack_weights_tf32(w, ops.PRECISION_BF16X2)
        xs = ops.rows_split(feats)
        t_split = timeit(lambda: ops.rows_split(feats), flush, args.reps)
        line = f"subm {c:3d} N={n:7d} P={p:8d} pairs/27N={p / (27.0 * n):.3f} tile-tap live={live:.3f} | v5 {t_old:.3f} ms | split {t_split:.3f} ms"
        for v in [int(s) for s in args.variants.split(",")]:
            setv(v)
            for use_masks in (False, True):
                o2 = ops.spconv_fwd(feats, w, tbl, n, precision=ops.PRECISION_BF16X2, packed=pk,
                                    tile_masks=masks if use_masks else None)
                err = float((o2 - ref).abs().max() / ref.abs().max())
                t = timeit(lambda: ops.spconv_fwd(feats, w, tbl, n, precision=ops.PRECISION_BF16X2, packed=pk, out=out,
                                                  tile_masks=masks if use_masks else None), flush, args.reps)
                line += f" | v{v}{'m' if use_masks else ' '} {t:.3f} ms (err {err:.1e})"
        setv(0)
        gt, gperm, gmasks = ops.table_group_rows(tbl, n)
        t_grp = timeit(lambda: ops.table_group_rows(tbl, n), flush, args.reps)
        glive = float(sum(bin(int(m) & 0x7ffffff).count("1") for m in gmasks.cpu().tolist())) / (27 * gmasks.numel())
        o3 = ops.spconv_fwd(feats, w, gt, n, precision=ops.PRECISION_BF16X2, packed=pk, tile_masks=gmasks, out_rows=gperm)
        o2 = ops.spconv_fwd(feats, w, tbl, n, precision=ops.PRECISION_BF16X2, packed=pk, tile_masks=masks)
        t = timeit(lambda: ops.spconv_fwd(feats, w, gt, n, precision=ops.PRECISION_BF16X2, packed=pk, out=out, tile_masks=gmasks,
                                          out_rows=gperm), flush, args.reps)
        line += f" | grouped live={glive:.3f} {t:.3f} ms (bit-equal {bool(torch.equal(o2, o3))}; grouping itself {t_grp:.3f} ms)"
        print(line, flush=True)
        if True:             # executed blocks of the grouped launch: (offset, chunk) x tile groups of T tiles
            T = 2 if c == 128 else 4
            mk = gmasks.cpu().numpy().astype("int64") & 0x7ffffff
            pad = (-len(mk)) % T
            mg = np.concatenate([mk, np.zeros(pad, "int64")]).reshape(-1, T)
            kps = 2 if c == 16 else 1
            def steps(m):       # live offset steps of a mask (two offsets per step at 16 channels)
                if kps == 1:
                    return np.array([bin(int(v)).count("1") for v in m])
                return np.array([sum(1 for q in range(14) if (int(v) >> (2 * q)) & 3) for v in m])
            nchunk = max(1, c // 32)
            union = np.bitwise_or.reduce(mg, axis=1)
            blocks = steps(union) * nchunk
            tile_blocks = steps(mk).sum() * nchunk
            clk = t * 1e-3 * 1.965e9 * 148
            print(f"   grouped {c:3d}: {len(mg)} groups of {T} tiles, {int(blocks.sum())} blocks ({blocks.min()}..{blocks.max()} per group), "
                  f"{int(tile_blocks)} live tile-blocks = {tile_blocks / max(1, blocks.sum()):.2f} per block; "
                  f"{clk / blocks.sum():.0f} clk per block, {clk / tile_blocks:.0f} clk per live tile-block", flush=True)
        if args.wgrad and c >= 32:       # tensor-core weight gradient of the same layer (wgrad_tc.cu)
            from sparse2dense_b200 import autograd as AG
            dy = torch.randn(n, c, device="cuda")
            ops.rows_split(dy, cache=True)
            ops.rows_split(feats, cache=True)
            t_w = timeit(lambda: AG.conv_wgrad(feats, dy, tbl, n), flush, args.reps)
            print(f"   wgrad {c:3d}x{c:3d} K=27 rows {n}: {t_w:.3f} ms", flush=True)
        if args.insitu:      # as the layer runs inside a residual block: BN affine + residual + ReLU, grouped rulebook
            sc, sh, res = torch.rand(c, device="cuda") + 0.5, torch.randn(c, device="cuda"), torch.randn(n, c, device="cuda")
            line = f"   in-situ {c:3d} (grouped, affine + residual + ReLU):"
            for v in [int(s) for s in args.variants.split(",")]:
                setv(v)
                t = timeit(lambda: ops.spconv_fwd(feats, w, gt, n, scale=sc, shift=sh, residual=res, relu=True,
                                                  precision=ops.PRECISION_BF16X2, packed=pk, out=out, tile_masks=gmasks,
                                                  out_rows=gperm), flush, args.reps)
                t2 = timeit(lambda: ops.spconv_fwd(feats, w, gt, n, scale=sc, shift=sh, relu=True,
                                                   precision=ops.PRECISION_BF16X2, packed=pk, out=out, tile_masks=gmasks,
                                                   out_rows=gperm), flush, args.reps)
                t3 = timeit(lambda: ops.spconv_fwd(feats, w, gt, n, precision=ops.PRECISION_BF16X2, packed=pk, out=out,
                                                   tile_masks=gmasks, out_rows=gperm, want_split=False), flush, args.reps)
                line += f" | v{v} {t:.3f} ms, no residual {t2:.3f}, plain without split {t3:.3f}"
            setv(0)
            print(line, flush=True)
            line = f"   in-situ {c:3d} cache-policy flags (affine + residual + ReLU):"
            for fl in (0, 256, 512, 768, 1024):
                setf(fl)
                t = timeit(lambda: ops.spconv_fwd(feats, w, gt, n, scale=sc, shift=sh, residual=res, relu=True,
                                                  precision=ops.PRECISION_BF16X2, packed=pk, out=out, tile_masks=gmasks,
                                                  out_rows=gperm), flush, max(args.reps, 15))
                line += f" f{fl}={t:.4f}"
            setf(0)
            print(line, flush=True)
        if args.prof:          # grouped launch: who waits for whom
            lib = ctypes.CDLL(_lib.LIB_PATH)
            lib.s2d_debug_bf2_prof.argtypes = [ctypes.c_void_p]
            buf = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
            lib.s2d_debug_bf2_prof(buf.data_ptr())
            ops.spconv_fwd(feats, w, gt, n, precision=ops.PRECISION_BF16X2, packed=pk, out=out, tile_masks=gmasks, out_rows=gperm)
            torch.cuda.synchronize()
            lib.s2d_debug_bf2_prof(None)
            m = buf.view(148, 16).double().mean(0).tolist()
            mx = buf.view(148, 16).double().max(0).values.tolist()
            print(f"   prof {c:3d} GROUPED: producer0 total {m[0]:.0f} clk (max {mx[0]:.0f}): wait_empty {m[1]:.0f} wait_data {m[2]:.0f} wait_list {m[3]:.0f} over {m[4]:.0f} steps "
                  f"| mma0 total {m[5]:.0f}: wait_a {m[6]:.0f} wait_b {m[7]:.0f} wait_acc {m[8]:.0f} wait_list {m[9]:.0f} over {m[10]:.0f} steps", flush=True)
        if args.prof:
            lib = ctypes.CDLL(_lib.LIB_PATH)
            lib.s2d_debug_bf2_prof.argtypes = [ctypes.c_void_p]
            for pv, fl in [(v, f) for v in [int(s) for s in args.variants.split(",")] for f in (0, 227)]:
                buf = torch.zeros(148 * 16, dtype=torch.int64, device="cuda")
                setv(pv)
                setf(fl)
                lib.s2d_debug_bf2_prof(buf.data_ptr())
                ops.spconv_fwd(feats, w, tbl, n, precision=ops.PRECISION_BF16X2, packed=pk, out=out)
                torch.cuda.synchronize()
                lib.s2d_debug_bf2_prof(None)
                setf(0)
                setv(0)
                m = buf.view(148, 16).double().mean(0).tolist()
                print(f"   prof {c:3d} variant {pv} flags {fl}: producer0 total {m[0]:.0f} clk: wait_empty {m[1]:.0f} wait_data {m[2]:.0f} wait_list {m[3]:.0f} over {m[4]:.0f} steps "
                      f"| mma total {m[5]:.0f}: wait_a {m[6]:.0f} wait_b {m[7]:.0f} wait_acc {m[8]:.0f} wait_list {m[9]:.0f} over {m[10]:.0f} steps", flush=True)
        if args.ablate:
            for av in [int(s) for s in args.variants.split(",")]:
                setv(av)
                line = f"   ablate {c:3d} variant {av}:"
                for fl in [int(v) for v in args.ablate.split(",")]:
                    setf(fl)
                    t = timeit(lambda: ops.spconv_fwd(feats, w, tbl, n, precision=ops.PRECISION_BF16X2, packed=pk, out=out), flush, args.reps)
                    line += f" f{fl}={t:.3f}"
                setf(0)
                setv(0)
                print(line, flush=True)
    if args.dense:
        from sparse2dense_b200 import dense
        B = args.batch
        for (H, cin, cout, k) in [(188, 256, 128, 3), (188, 128, 128, 3), (94, 256, 256, 3), (47, 256, 1024, 1), (47, 1024, 256, 1),
                                  (188, 512, 64, 3), (188, 64, 320, 3)]:
            tbl, Ho, Wo = dense.conv_table(torch.device("cuda"), B, H, H, k, 1, k // 2)
            n = B * H * H
            x = torch.relu(torch.randn(n, cin, device="cuda"))
            w = torch.randn(k * k, cin, cout, device="cuda") / (k * k * cin) ** 0.5
            out = torch.empty(n, cout, device="cuda")
            old = ops.PRECISION_TF32_BF16C if cout % 128 == 0 else ops.PRECISION_TF32X3
            pk_old = ops.pack_weights_tf32(w, old)
            ref = dense.conv_rows(x, w, tbl, n, precision=ops.PRECISION_TF32X3)
            t_old = timeit(lambda: dense.conv_rows(x, w, tbl, n, out=out, precision=old, packed=pk_old), flush, args.reps)
            pk = ops.pack_weights_tf32(w, ops.PRECISION_BF16X2)
            xs = ops.rows_split(x)
            line = f"dense {H}x{H} {cin}->{cout} k{k} | v5 {t_old:.3f} ms"
            for v in [int(s) for s in args.variants.split(",")]:
                setv(v)
                ops.conv_launch(None, pk, tbl, n, cin, cout, k * k, out=out, precision=ops.PRECISION_BF16X2, x_split=xs)
                err = float((out - ref).abs().max() / ref.abs().max())
                t = timeit(lambda: ops.conv_launch(None, pk, tbl, n, cin, cout, k * k, out=out, precision=ops.PRECISION_BF16X2,
                                                   x_split=xs), flush, args.reps)
                fl = 2.0 * n * k * k * cin * cout / (t * 1e-3) / 1e12
                line += f" | v{v} {t:.3f} ms {fl:.0f} TF/s (err {err:.1e})"
            setv(0)
            print(line, flush=True)


if __name__ == "__main__":
    main()
