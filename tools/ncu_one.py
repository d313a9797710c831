#!/usr/bin/env python
"""One profiled launch of the 128->128 submanifold layer (batch 4, scan-order rulebook, no masks) with the conv_bf2 debug
switches from the environment: FLAGS (B2Args::dbg ablation bits), VARIANT (s2d_debug_bf2_variant).
usage: ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/x python tools/ncu_one.py"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparse2dense_b200 import _lib, ops, spconv, synth  # noqa: E402
from sparse2dense_b200.backbones import SpMiddleResNetFHD  # noqa: E402
from sparse2dense_b200.hotpath import concat_clouds  # noqa: E402


def main():
    _lib.load()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    pts, offs = concat_clouds(synth.lidar_batch(1, 4))
    vb = ops.voxelize(pts.cuda(), offs, synth.WAYMO_VOXEL, synth.WAYMO_RANGE, 5, 150000, want_voxels=False, mean_channels=5)
    bb = SpMiddleResNetFHD(num_input_features=5).cuda().eval()
    x = spconv.SparseConvTensor(vb.mean_buffer[:vb.n], vb.coors_buffer[:vb.n], (41, 1504, 1504), 4)
    plan = spconv.plan_coords(x, [bb.conv2[0], bb.conv3[0], bb.conv4[0], bb.extra_conv[0]])
    sc = plan[2]
    n, c = sc.coors.shape[0], 128
    tbl = ops.rulebook_subm(sc.coors, sc.index, 3)
    feats = torch.relu(torch.randn(n, c, device="cuda"))
    w = torch.randn(3, 3, 3, c, c, device="cuda") / (27 * c) ** 0.5
    pk = ops.pack_weights_tf32(w, ops.PRECISION_BF16X2)
    ops.rows_split(feats, cache=True)
    out = torch.empty(n, c, device="cuda")
    lib.s2d_debug_bf2_variant(int(os.environ.get("VARIANT", "0")))
    lib.s2d_debug_bf2_flags(int(os.environ.get("FLAGS", "0")))
    fn = lambda: ops.spconv_fwd(feats, w, tbl, n, precision=ops.PRECISION_BF16X2, packed=pk, out=out)  # noqa: E731
    fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()
