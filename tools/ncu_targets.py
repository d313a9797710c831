#!/usr/bin/env python
"""One profiled launch of every hot kernel on its real shape (for `ncu --set full --profile-from-start off`):
the four submanifold stages of the backbone (grouped rulebooks, batch 4), the strided 16->32 layer, three dense neck / head
shapes and the tensor-core weight gradient.  Each kernel is launched once untimed (warm-up, outside the profiled range) and
once between cudaProfilerStart / Stop.
usage: ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/x python tools/ncu_targets.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparse2dense_b200 import autograd as AG, dense, ops, spconv, synth  # noqa: E402
from sparse2dense_b200.backbones import SpMiddleResNetFHD  # noqa: E402
from sparse2dense_b200.hotpath import concat_clouds  # noqa: E402


def profiled(fn):
    fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    fn()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


def main():
    batch = int(os.environ.get("BATCH", "4"))
    pts, offs = concat_clouds(synth.lidar_batch(1, batch))
    vb = ops.voxelize(pts.cuda(), offs, synth.WAYMO_VOXEL, synth.WAYMO_RANGE, 5, 150000, want_voxels=False, mean_channels=5)
    n0 = vb.n
    bb = SpMiddleResNetFHD(num_input_features=5).cuda().eval()
    x = spconv.SparseConvTensor(vb.mean_buffer[:n0], vb.coors_buffer[:n0], (41, 1504, 1504), batch)
    plan = spconv.plan_coords(x, [bb.conv2[0], bb.conv3[0], bb.conv4[0], bb.extra_conv[0]])
    stages = [(x.indices, x.index())] + [(sc.coors, sc.index) for sc in plan]
    for si, c in enumerate([16, 32, 64, 128]):
        coors, index = stages[si]
        n = coors.shape[0]
        tbl = ops.rulebook_subm(coors, index, 3)
        gt, perm, masks = ops.table_group_rows(tbl, n)
        feats = torch.relu(torch.randn(n, c, device="cuda"))
        w = torch.randn(3, 3, 3, c, c, device="cuda") / (27 * c) ** 0.5
        pk = ops.pack_weights_tf32(w, ops.PRECISION_BF16X2)
        ops.rows_split(feats, cache=True)
        out = torch.empty(n, c, device="cuda")
        print(f"subm {c}->{c} K=27 rows {n}", flush=True)
        profiled(lambda: ops.spconv_fwd(feats, w, gt, n, precision=ops.PRECISION_BF16X2, packed=pk, out=out, tile_masks=masks,
                                        out_rows=perm))
        if c == 128:        # the weight gradient of the same layer (wgrad_tc.cu)
            dy = torch.randn(n, c, device="cuda")
            ops.rows_split(dy, cache=True)
            print(f"wgrad {c}x{c} K=27 rows {n}", flush=True)
            profiled(lambda: AG.conv_wgrad(feats, dy, tbl, n))
    B = batch
    for (H, cin, cout, k) in [(94, 256, 256, 3), (188, 128, 128, 3), (188, 256, 256, 1), (188, 512, 64, 3)]:
        tbl, Ho, Wo = dense.conv_table(torch.device("cuda"), B, H, H, k, 1, k // 2)
        n = B * H * H
        xd = torch.relu(torch.randn(n, cin, device="cuda"))
        w = torch.randn(k * k, cin, cout, device="cuda") / (k * k * cin) ** 0.5
        out = torch.empty(n, cout, device="cuda")
        pk = ops.pack_weights_tf32(w, ops.PRECISION_BF16X2)
        xs = ops.rows_split(xd)
        print(f"dense {H}x{H} {cin}->{cout} K={k * k} rows {n}", flush=True)
        profiled(lambda: ops.conv_launch(None, pk, tbl, n, cin, cout, k * k, out=out, precision=ops.PRECISION_BF16X2, x_split=xs))


if __name__ == "__main__":
    main()
