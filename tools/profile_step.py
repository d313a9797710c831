#!/usr/bin/env python
"""Per-kernel device time of one step of a bench.py workload (torch.profiler, CUDA activities only).
usage: profile_step.py [--config full|backbone|pillar|train] [--batch N]"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sparse2dense_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="full")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--top", type=int, default=40)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    batch = a.batch or bench.WORKLOADS[a.config]["batch"]
    wl = bench.TrainWorkload(a.config, batch, ops.PRECISION_AUTO, dev, 0, 1) if a.config == "train" else \
        bench.ForwardWorkload(a.config, batch, ops.PRECISION_AUTO, dev, 0)
    for _ in range(3):
        wl.step()
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        wl.step()
        torch.cuda.synchronize()
    agg = {}
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            v = agg.setdefault(e.name.split("(")[0][:90], [0.0, 0])
            v[0] += e.device_time / 1e3
            v[1] += 1
    tot = sum(v[0] for v in agg.values())
    print(f"{a.config} batch {batch}: kernel time of one step {tot:.2f} ms over {sum(v[1] for v in agg.values())} launches")
    for k, (m, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:a.top]:
        print(f"  {m:8.3f} ms {n:5d}  {k}")


if __name__ == "__main__":
    main()
