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
    a = ap.parse_args()
    prec = ops.PRECISION_NAMES[a.precision]
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


if __name__ == "__main__":
    main()
