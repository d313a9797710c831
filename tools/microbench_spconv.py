#!/usr/bin/env python
"""Per-layer timing of the sparse conv kernels on the real rulebooks of a batch of synthetic scenes.
usage: microbench_spconv.py [--batch 4] [--reps 10] [--only CIN] [--precisions fp32,tf32,tf32x3]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparse2dense_b200 import ops, spconv, synth  # noqa: E402
from sparse2dense_b200.backbones import SpMiddleResNetFHD  # noqa: E402
from sparse2dense_b200.hotpath import concat_clouds  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--only", type=int, default=0)
    ap.add_argument("--precisions", default="fp32,tf32,tf32x3,tf32_bf16c")
    ap.add_argument("--gather", type=int, default=0, help="0: cp.async gather ring (default), 2: TMA gather4 where possible")
    args = ap.parse_args()
    prec = ops.PRECISION_NAMES
    import ctypes
    from sparse2dense_b200 import _lib
    _lib.load()
    ctypes.CDLL(_lib.LIB_PATH).s2d_debug_tc_gather(args.gather)
    clouds = synth.lidar_batch(1, args.batch)
    pts, offs = concat_clouds(clouds)
    vb = ops.voxelize(pts.cuda(), offs, synth.WAYMO_VOXEL, synth.WAYMO_RANGE, 5, 150000, want_voxels=False,
                      mean_channels=5)
    n0 = vb.n
    bb = SpMiddleResNetFHD(num_input_features=5).cuda().eval()
    x = spconv.SparseConvTensor(vb.mean_buffer[:n0], vb.coors_buffer[:n0], (41, 1504, 1504), args.batch)
    plan = spconv.plan_coords(x, [bb.conv2[0], bb.conv3[0], bb.conv4[0], bb.extra_conv[0]])
    stages = [(x.indices, x.index())] + [(sc.coors, sc.index) for sc in plan]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    print(f"rows per stage: {[int(s[0].shape[0]) for s in stages]}")
    for si, c in enumerate([16, 32, 64, 128]):
        if args.only and c != args.only:
            continue
        coors, index = stages[si]
        n = coors.shape[0]
        tbl, pairs = ops.rulebook_subm(coors, index, 3, count_pairs=True)
        p = int(pairs.item())
        feats = torch.randn(n, c, device="cuda")
        w = torch.randn(3, 3, 3, c, c, device="cuda") / (27 * c) ** 0.5
        for name in args.precisions.split(","):
            pr = prec[name]
            if pr != ops.PRECISION_FP32 and not ops.tf32_supported(c, c):
                continue
            packed = ops.pack_weights_tf32(w, pr) if pr != ops.PRECISION_FP32 else None
            out = ops.spconv_fwd(feats, w, tbl, n, precision=pr, packed=packed)
            torch.cuda.synchronize()
            ms = []
            for _ in range(args.reps):
                flush.fill_(0)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                ops.spconv_fwd(feats, w, tbl, n, precision=pr, packed=packed, out=out)
                e1.record()
                torch.cuda.synchronize()
                ms.append(e0.elapsed_time(e1))
            t = float(np.median(ms))
            tfl = 2.0 * p * c * c / (t * 1e-3) / 1e12
            dense_tfl = 2.0 * n * 27 * c * c / (t * 1e-3) / 1e12
            gbs = (2 * n * c * 4 + n * 27 * 4 + 27 * c * c * 4) / (t * 1e-3) / 1e9
            print(f"subm {c:3d}->{c:3d} N={n:7d} P={p:8d} {name:7s} {t:7.3f} ms  useful {tfl:6.1f} TFLOP/s  "
                  f"dense-tile {dense_tfl:6.1f} TFLOP/s  compulsory {gbs:6.0f} GB/s")


if __name__ == "__main__":
    main()
