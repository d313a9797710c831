import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from sparse2dense_b200 import ops, registry, synth, second_stage
from sparse2dense_b200.hotpath import concat_clouds
# small end-to-end: voxelize -> backbone (auto precision) on one small scene; decode + nms; second stage
cloud = synth.small_scene(5)
pts, offs = concat_clouds([cloud])
state = synth.backbone_state(0)
bb = registry.build_backbone(dict(type="SpMiddleResNetFHD", num_input_features=5, ds_factor=8))
bb.load_state_dict({k: torch.as_tensor(v) for k, v in state.items()}, strict=False)
bb = bb.cuda().eval(); bb.set_precision(ops.PRECISION_AUTO)
vb = ops.voxelize(pts.cuda(), offs, synth.WAYMO_VOXEL, synth.WAYMO_RANGE, 5, 150000, want_voxels=False, mean_channels=5)
with torch.no_grad():
    bev, _ = bb(vb.mean, vb.coors, 1, [1504, 1504, 40])
print("bev", bev.shape, float(bev.abs().max()))
rng = np.random.default_rng(0)
B, H, W = 2, 40, 36
rows = torch.from_numpy(rng.normal(-2, 2, (B * H * W, 32)).astype(np.float32)).cuda()
heads = dict(reg=rows[:, 0:2], height=rows[:, 2:3], dim=rows[:, 3:6] * 0.2, rot=rows[:, 6:8], hm=rows[:, 8:11])
heads["dim"] = (rows[:, 3:6] * 0.2).contiguous()
boxes, scores, labels, keys = ops.centerhead_decode(heads, B, H, W, 8, [0.1, 0.1], [-75.2, -75.2], 0.1, [-80, -80, -10, 80, 80, 10])
out = ops.centerhead_select(keys, boxes, scores, labels, B, H * W, 4096, 0.7, 500)
print("dets", out[4].cpu().tolist())
ext = second_stage.BEVFeatureExtractor([-75.2, -75.2], [0.1, 0.1], 8)
bevr = torch.randn(B * H * W, 64, device="cuda")
f = ext.box_features(bevr, B, H, W, out[0], out[4], 5)
print("feat", f.shape)
torch.cuda.synchronize()
# pillars + losses
from sparse2dense_b200 import losses as L, dense
import logging
v = torch.zeros(300, 20, 5, device="cuda"); n = torch.randint(1, 21, (300,), device="cuda", dtype=torch.int32)
for i in range(300):
    v[i, : int(n[i])] = torch.randn(int(n[i]), 5, device="cuda")
c = torch.stack([torch.zeros(300), torch.zeros(300), torch.arange(300) // 20, torch.arange(300) % 20], 1).int().cuda()
reader = registry.build_reader(dict(type="PillarFeatureNet", num_filters=[64, 64], num_input_features=5, voxel_size=(0.32, 0.32, 6.0),
                                    pc_range=(-74.88, -74.88, -2, 74.88, 74.88, 4.0))).cuda().eval()
print("pfn", reader(v, n, c).shape)
D = dense.DenseOps(ops.PRECISION_AUTO)
x = torch.randn(2 * 14 * 10, 64, device="cuda")
y, h, w = D.maxpool2(x, 2, 14, 10); u = D.upsample_nearest(y, 2, h, w, 27, 19)
conv = torch.nn.ConvTranspose2d(64, 96, 4, 4, 0, bias=False).cuda().eval()
t, _, _ = D.tconv("t", x, 2, 14, 10, conv)
print("dense helpers", y.shape, u.shape, t.shape)
a, b = torch.randn(2, 8, 36, 40, device="cuda"), torch.randn(2, 8, 36, 40, device="cuda")
ind = torch.randint(0, 36 * 40, (2, 60), device="cuda"); mask = torch.rand(2, 60, device="cuda") < 0.5; cat = torch.randint(0, 8, (2, 60), device="cuda")
print("losses", float(L.sparse2dense_loss(a, b, a, b)), float(L.fastfocalloss(a, b, ind, mask, cat, True, True)), L.distill_reg_loss(a, b, mask, ind).shape)
torch.cuda.synchronize()
