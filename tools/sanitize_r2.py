"""compute-sanitizer targets of round 2 (small shapes): grouped rulebook builders + grouped conv, dense-grid TMA conv,
tensor-core weight gradient, staged NMS, top-k.  usage: compute-sanitizer --tool memcheck python tools/sanitize_r2.py"""
import sys
import numpy as np
import torch
sys.path.insert(0, '/root/repo')
from sparse2dense_b200 import autograd as AG, dense, ops, synth
from sparse2dense_b200.hotpath import concat_clouds

import ctypes, os
from sparse2dense_b200 import _lib
if os.environ.get("S2D_BF2_VARIANT"):            # e.g. 2: the unswapped kernel for 128 output channels (A/B under the tool)
    _lib.load()
    ctypes.CDLL(_lib.LIB_PATH).s2d_debug_bf2_variant(int(os.environ["S2D_BF2_VARIANT"]))
pts, offs = concat_clouds([synth.small_scene(5), synth.small_scene(6)])
vb = ops.voxelize(pts.cuda(), offs, synth.WAYMO_VOXEL, synth.WAYMO_RANGE, 5, 150000, want_voxels=False, mean_channels=5)
n = vb.n
coors, shape = vb.coors_buffer[:n], (41, 1504, 1504)
index = ops.build_grid_index(coors, 2, shape)
tbl, perm, masks = ops.rulebook_subm_grouped(coors, index)
x = torch.relu(torch.randn(n, 32, device="cuda"))
w = torch.randn(27, 32, 32, device="cuda") * 0.05
y = ops.spconv_fwd(x, w, tbl, n, None, None, x, True, ops.PRECISION_BF16X2, tile_masks=masks, out_rows=perm)
sc = ops.sparse_out_coords(coors, n, 2, shape, 3, 2, 1)
t2, p2, m2 = ops.rulebook_sparse_grouped(sc.coors, index, 2, 1)
y2 = ops.spconv_fwd(x, torch.randn(27, 32, 64, device="cuda") * 0.05, t2, sc.coors.shape[0], precision=ops.PRECISION_BF16X2,
                    tile_masks=m2, out_rows=p2)
print("grouped convs", float(y.abs().max()), float(y2.abs().max()))
t3, p3, m3 = ops.table_group_rows(ops.rulebook_subm(coors, index, 3), n)
assert torch.equal(t3[:, :n], tbl[:, :n])
B, H, W = 2, 21, 37
xd = torch.relu(torch.randn(B * H * W, 64, device="cuda"))
wd = torch.randn(9, 64, 128, device="cuda") * 0.05
tb, _, _ = dense.conv_table(torch.device("cuda"), B, H, W, 3, 1, 1)
a = dense.conv_rows(xd, wd, tb, B * H * W, precision=ops.PRECISION_BF16X2)
b = dense.conv_rows(xd, wd, tb, B * H * W, precision=ops.PRECISION_BF16X2, grid=(B, H, W, 3, 1))
assert torch.equal(a, b)
print("grid conv ok")
g = torch.randn(n, 64, device="cuda")
d = torch.randn(n, 128, device="cuda")
dw = AG.conv_wgrad(g, d, ops.rulebook_subm(coors, index, 3), n)
dw2 = AG.conv_wgrad(torch.randn(n, 32, device="cuda"), torch.randn(n, 32, device="cuda"), ops.rulebook_subm(coors, index, 3), n)
print("wgrad", float(dw.abs().max()), float(dw2.abs().max()))
rng = np.random.default_rng(0)
Hh, Wh = 40, 36
rows = torch.from_numpy(rng.normal(-1, 2, (B * Hh * Wh, 32)).astype(np.float32)).cuda()
heads = dict(reg=rows[:, 0:2], height=rows[:, 2:3], dim=(rows[:, 3:6] * 0.2).contiguous(), rot=rows[:, 6:8], hm=rows[:, 8:11])
boxes, scores, labels, keys = ops.centerhead_decode(heads, B, Hh, Wh, 8, [0.1, 0.1], [-75.2, -75.2], 0.1, [-80, -80, -10, 80, 80, 10])
out = ops.centerhead_select(keys, boxes, scores, labels, B, Hh * Wh, 4096, 0.7, 500)
print("dets", out[4].cpu().tolist())
torch.cuda.synchronize()
print("done")
