#!/usr/bin/env python
"""Timing of one distillation training step (teacher forward, student forward + backward, gradient clip, Adam) on
synthetic Waymo-shaped batches (BASELINE configs[4]).  usage: bench_train.py [--batch 2] [--steps 5] [--no-pcr] [--small]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparse2dense_b200 import ops, synth  # noqa: E402
from sparse2dense_b200.trainer import DistillTrainer  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--precision", default="auto")
    ap.add_argument("--no-pcr", action="store_true")
    ap.add_argument("--small", action="store_true")
    ap.add_argument("--sync-bn", action="store_true", help="nn.SyncBatchNorm.convert_sync_batchnorm (tools/train.py:92-96)")
    ap.add_argument("--profile", action="store_true", help="print the per-kernel device time of one step (torch.profiler)")
    a = ap.parse_args()
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    if world > 1:                                              # DDP: one process per GPU, scenes sharded over ranks
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    teacher, student = synth.build_distill_models("cuda", ops.PRECISION_NAMES[a.precision])
    student.neck.train_pcr = not a.no_pcr
    if a.sync_bn:
        student = torch.nn.SyncBatchNorm.convert_sync_batchnorm(student)
    tr = DistillTrainer(teacher, student, total_steps=1000)
    ex = synth.distill_example(a.batch, cfg=1 + rank, small=a.small)
    print("voxels student/dense/recon:", ex["voxels"].shape[0], ex["dense_voxels"].shape[0], ex["reconstruction_voxels"].shape[0])
    for i in range(a.warmup):
        log = tr.step(ex)
        print("warmup", i, {k: (round(float(v), 5)) for k, v in log.items()})
    torch.cuda.synchronize()
    ms = []
    for i in range(a.steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        log = tr.step(ex)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
        print("step", i, f"{ms[-1]:.1f} ms (wall {1e3 * (time.perf_counter() - t0):.1f})", {k: round(float(v), 5) for k, v in log.items()})
    if a.profile:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            tr.step(ex)
            torch.cuda.synchronize()
        agg = {}
        for e in prof.events():
            if e.device_type == torch.autograd.DeviceType.CUDA:
                v = agg.setdefault(e.name.split("(")[0][:100], [0.0, 0])
                v[0] += e.device_time / 1e3
                v[1] += 1
        tot = sum(v[0] for v in agg.values())
        print(f"kernel time of one step: {tot:.1f} ms over {sum(v[1] for v in agg.values())} launches")
        for k, (m, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:32]:
            print(f"  {m:8.2f} ms {n:5d}  {k}")
    t = float(np.median(ms))
    if world > 1:
        tt = torch.tensor([t], device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t = float(tt)
        w0 = next(student.parameters()).detach().flatten()[:1000].clone()
        w1 = w0.clone()
        dist.broadcast(w1, src=0)
        print(f"rank {rank}: weights identical to rank 0 after {a.warmup + a.steps} steps: {bool(torch.equal(w0, w1))}")
        if rank != 0:
            return
        a.batch *= world
    print(f"distillation step batch {a.batch} {a.precision} pcr={not a.no_pcr}: {t:.1f} ms/step -> {a.batch / t * 1e3:.2f} scenes/s; "
          f"peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")


if __name__ == "__main__":
    main()
