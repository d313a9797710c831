#!/usr/bin/env python
"""Timing of the full two-stage forward (voxelize -> backbone -> S2D_RPN -> CenterHead -> decode + NMS -> RoI head), per stage.
usage: bench_full.py [--batch 4] [--steps 5] [--precision auto]"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparse2dense_b200 import ops, synth  # noqa: E402
from sparse2dense_b200.hotpath import FullForwardPath, concat_clouds  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--precision", default="auto")
    ap.add_argument("--model", default="voxelnet", choices=["voxelnet", "pp"])
    a = ap.parse_args()
    prec = ops.PRECISION_NAMES[a.precision]
    if a.model == "pp":
        return bench_pp(a, prec)
    path = FullForwardPath(state=synth.backbone_state(0), precision=prec)
    path.neck.load_state_dict({k: torch.as_tensor(v) for k, v in synth.random_module_state(path.neck, 11).items()}, strict=False)
    path.head.load_state_dict({k: torch.as_tensor(v) for k, v in synth.random_module_state(path.head, 12).items()}, strict=False)
    with torch.no_grad():
        path.head.tasks[0].hm[-1].bias.fill_(-1.0)          # random weights: keep a realistic number of confident cells
    pts, offs = concat_clouds(synth.lidar_batch(1, a.batch))
    pts = pts.cuda()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        path.forward_points(pts, offs)
    torch.cuda.synchronize()
    ms = []
    ops.KERNEL_EVENTS = []
    for _ in range(a.steps):
        flush.fill_(0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        path.forward_points(pts, offs)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    ev, ops.KERNEL_EVENTS = ops.KERNEL_EVENTS, None
    t = float(np.median(ms))
    print(f"full forward batch {a.batch} {a.precision}: {t:.3f} ms/step -> {a.batch / t * 1e3:.1f} scenes/s")
    groups = {}
    for key, s, e in ev:
        g = groups.setdefault(key[:3] + (key[5],), [0.0, 0])
        g[0] += s.elapsed_time(e); g[1] += 1
    tot = 0.0
    for k, (m, n) in sorted(groups.items(), key=lambda kv: -kv[1][0]):
        tot += m / a.steps
        fl = 2.0 * k[3] * k[2] * k[0] * k[1] * n / a.steps
        print(f"  Cin {k[0]:4d} Cout {k[1]:4d} K {k[2]:2d} rows {k[3]:7d}: {m / a.steps:7.3f} ms/step over {n // a.steps:2d} launches, "
              f"dense {fl / (m / a.steps * 1e-3) / 1e12:6.1f} TFLOP/s")
    print(f"  conv kernels total {tot:.3f} ms/step")
    out = path.forward_points(pts, offs)
    print("  detections per scene:", out[3].cpu().tolist())


def bench_pp(a, prec):
    """BASELINE configs[3] shape: pillar voxelize (0.32 m, 20 pts) -> PFN -> scatter + S2D -> RPN[3,5,5] -> CenterHead -> NMS."""
    import logging
    from sparse2dense_b200 import registry
    from sparse2dense_b200.voxel_generator import VoxelGenerator
    VS, RG = (0.32, 0.32, 6.0), (-74.88, -74.88, -2, 74.88, 74.88, 4.0)
    gen = VoxelGenerator(VS, RG, 20, 32000)
    reader = registry.build_reader(dict(type="PillarFeatureNet", num_filters=[64, 64], num_input_features=5, voxel_size=VS, pc_range=RG))
    bb = registry.build_backbone(dict(type="PointPillarsScatter_S2D", ds_factor=1))
    neck = registry.build_neck(dict(type="RPN", layer_nums=[3, 5, 5], ds_layer_strides=[1, 2, 2], ds_num_filters=[64, 128, 256],
                                    us_layer_strides=[1, 2, 4], us_num_filters=[128, 128, 128], num_input_features=64,
                                    logger=logging.getLogger("RPN")))
    head = registry.build_head(dict(type="CenterHead", in_channels=384, tasks=[dict(num_class=3, class_names=["V", "P", "C"])],
                                    dataset="waymo", weight=2, code_weights=[1.0] * 8,
                                    common_heads={"reg": (2, 2), "height": (1, 2), "dim": (3, 2), "rot": (2, 2)}))
    for m, seed in ((reader, 31), (bb, 32), (neck, 33), (head, 34)):
        m.load_state_dict({k: torch.as_tensor(v) for k, v in synth.random_module_state(m, seed).items()}, strict=False)
        m.cuda().eval()
        if hasattr(m, "set_precision"):
            m.set_precision(prec)
    test_cfg = dict(post_center_limit_range=[-80, -80, -10.0, 80, 80, 10.0],
                    nms=dict(nms_pre_max_size=4096, nms_post_max_size=500, nms_iou_threshold=0.7),
                    score_threshold=0.1, pc_range=[-74.88, -74.88], out_size_factor=1, voxel_size=[0.32, 0.32])
    pts, offs = concat_clouds(synth.lidar_batch(1, a.batch))
    pts = pts.cuda()
    B = a.batch

    @torch.no_grad()
    def step():
        vb = gen.generate_batch(pts, offs, want_voxels=True)
        n = vb.n
        f = reader(vb.voxels, vb.num_points, vb.coors)
        fa, fb, (H, W) = bb.forward_rows(f, vb.coors, B, [468, 468, 1])
        ups, (Hu, Wu) = neck.forward_rows(fa, B, H, W)
        preds = head.forward_rows(ups, B, Hu, Wu)
        return head.select_rows(preds, B, Hu, Wu, test_cfg)[0]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    ms = []
    for _ in range(a.steps):
        flush.fill_(0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = step(); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    t = float(np.median(ms))
    print(f"pillar student forward batch {B} {a.precision}: {t:.3f} ms/step -> {B / t * 1e3:.1f} scenes/s; "
          f"detections {out[4].cpu().tolist()}")


if __name__ == "__main__":
    main()
