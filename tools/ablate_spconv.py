#!/usr/bin/env python
"""Ablation timing of the tcgen05 sparse conv kernel: which pipeline stage bounds a step?
flags: 1 = no gather (cp.async skipped), 2 = no TMEM store, 4 = no MMA issue.  Results are WRONG with any flag set."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparse2dense_b200 import _lib, ops, spconv, synth  # noqa: E402
from sparse2dense_b200.backbones import SpMiddleResNetFHD  # noqa: E402
from sparse2dense_b200.hotpath import concat_clouds  # noqa: E402


def main():
    _lib.load()
    setf = ctypes.CDLL(_lib.LIB_PATH).s2d_debug_tc_flags
    clouds = synth.lidar_batch(1, 4)
    pts, offs = concat_clouds(clouds)
    vb = ops.voxelize(pts.cuda(), offs, synth.WAYMO_VOXEL, synth.WAYMO_RANGE, 5, 150000, want_voxels=False, mean_channels=5)
    bb = SpMiddleResNetFHD(num_input_features=5).cuda().eval()
    x = spconv.SparseConvTensor(vb.mean_buffer[:vb.n], vb.coors_buffer[:vb.n], (41, 1504, 1504), 4)
    plan = spconv.plan_coords(x, [bb.conv2[0], bb.conv3[0], bb.conv4[0], bb.extra_conv[0]])
    stages = [(x.indices, x.index())] + [(sc.coors, sc.index) for sc in plan]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    precs = sys.argv[1].split(",") if len(sys.argv) > 1 else ["tf32", "tf32_bf16c", "tf32x3"]
    FLAGS = [int(v) for v in sys.argv[2].split(",")] if len(sys.argv) > 2 else (0, 1, 2, 4, 3, 5, 6, 7)
    only = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    for si, c in enumerate([16, 32, 64, 128]):
        if only and c != only:
            continue
        coors, index = stages[si]
        n = coors.shape[0]
        tbl = ops.rulebook_subm(coors, index, 3)
        feats = torch.randn(n, c, device="cuda")
        w = torch.randn(3, 3, 3, c, c, device="cuda") / (27 * c) ** 0.5
        for name in precs:
            pr = ops.PRECISION_NAMES[name]
            packed = ops.pack_weights_tf32(w, pr)
            out = torch.empty(n, c, device="cuda")
            row = []
            for flags in FLAGS:
                setf(flags)
                ops.spconv_fwd(feats, w, tbl, n, precision=pr, packed=packed, out=out)
                torch.cuda.synchronize()
                ms = []
                for _ in range(5):
                    flush.fill_(0)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    ops.spconv_fwd(feats, w, tbl, n, precision=pr, packed=packed, out=out)
                    e1.record()
                    torch.cuda.synchronize()
                    ms.append(e0.elapsed_time(e1))
                row.append(f"f{flags}={np.median(ms):.3f}")
            setf(0)
            print(f"subm {c:3d} {name:10s} " + "  ".join(row), flush=True)


if __name__ == "__main__":
    main()
