#!/usr/bin/env python
"""bench.py -- scenes/s of the hot path on Waymo-shaped synthetic clouds (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config full|backbone|pillar|train]      our arm (CUDA, sm_100a)
    python bench.py --impl reference ...                                                           the CPU arm (host cores)

Workloads (BASELINE.json `configs`):
  full      configs[2]  CenterPoint-VoxelNet + S2D two-stage forward, batch 8 / GPU: points -> voxelize -> SpMiddleResNetFHD ->
                        S2D_RPN -> CenterHead -> decode + rotated NMS -> RoI head -> detections [B,500,*]      (DEFAULT, headline)
  backbone  configs[1]  voxelize -> SpMiddleResNetFHD -> dense BEV, batch 4 / GPU
  pillar    configs[3]  CenterPoint-Pillar + S2D two-stage forward, batch 4 / GPU (16 over 4 GPUs)
  train     configs[4]  distillation training step (teacher fwd, student fwd + bwd, NCCL gradient all-reduce, SyncBN when
                        N > 1, clip + Adam), batch 4 / GPU (32 over 8 GPUs)
The default run prints ONE JSON line for `full` and carries short measurements of the other three as sub-objects
(`also`), so that the driver's 1/2/4/8-GPU runs exercise the collective-free forward AND the all-reduce of the training step.

Keys: see the task contract; `value` = device-resident throughput, `e2e` = through the public API with pinned HOST
buffers (H2D + D2H inside the timed region), `roofline` = dominant sparse kernel timed live with CUDA events (+ `dense` =
the dominant dense kernel), `cpu_baseline` = the CPU restatement of the same workload on the host cores, one scene.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "scenes_per_sec"
UNIT = "scenes/s"
WORKLOADS = {
    "full": dict(batch=8, text="CenterPoint-VoxelNet+S2D two-stage full forward (voxelize -> SpMiddleResNetFHD -> S2D_RPN -> "
                               "CenterHead -> decode + rotated NMS -> RoIHead), batch=8 x ~180k-pt synthetic Waymo clouds / GPU, "
                               "grid 1504x1504x40 [BASELINE configs[2]]"),
    "backbone": dict(batch=4, text="voxelize+SpMiddleResNetFHD backbone -> BEV, batch=4 x ~180k-pt synthetic Waymo clouds / GPU, "
                                   "grid 1504x1504x40 [BASELINE configs[1]]"),
    "pillar": dict(batch=4, text="CenterPoint-Pillar+S2D two-stage forward (pillar voxelize 0.32 m -> PFN -> scatter+S2D -> "
                                 "RPN[3,5,5] -> CenterHead -> NMS -> RoIHead), batch=4 / GPU (16 over 4 GPUs) [BASELINE configs[3]]"),
    "train": dict(batch=4, text="CenterPoint-VoxelNet+S2D distillation training step (teacher fwd, student fwd+bwd incl. PCR, "
                                "NCCL gradient all-reduce + SyncBN when N>1, clip, Adam), batch=4 / GPU (32 over 8 GPUs) "
                                "[BASELINE configs[4]]"),
}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=float(p["hbm_gbs"]), bf16=float(p["bf16_tflops"]),
                    bf16_sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), source="MEASURED_PEAKS.json")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="B200_PROFILING.md fallback")


def measure_tf32_peak(dev):
    """Dense TF32 tensor throughput measured live (cuBLAS TF32 GEMM 8192^3, best of 5, CUDA events): the denominator of the
    tensor roofline (MEASURED_PEAKS.json only holds bf16).  Burst figure: the kernels it is compared with run inside
    sub-second timed regions."""
    import torch
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = True
    try:
        n = 8192
        a = torch.randn(n, n, device=dev)
        b = torch.randn(n, n, device=dev)
        torch.matmul(a, b)
        best = None
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(a, b); e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        return 2.0 * n ** 3 / (best * 1e-3) / 1e12
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


def ncu_traffic(kernel_key):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed
    `ncu --set full` capture (profiles/ncu_traffic.json names the report it was read from)."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        table = json.load(f)
    ent = table.get(kernel_key)
    return None if ent is None else ent["dram_bytes_per_launch"]


def config_dict(name, batch, precision, extra=None):
    cfg = {"workload": WORKLOADS[name]["text"], "name": name, "batch_per_gpu": batch, "points_in_range": "~180k/scene",
           "weights": "seeded random (sparse2dense_b200.synth)", "l2": "flushed between timed steps (256 MiB write)",
           "precision": precision}
    if extra:
        cfg.update(extra)
    return cfg


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------------------------
# CPU arm: the CPU restatement of the reference's algorithm on the host cores (bounded sample: one scene per step)
# ---------------------------------------------------------------------------------------------
class CpuScene:
    """One scene of workload `name` on the host: the C/OpenMP port of the numba voxelizer + spconv-v1 gather-GEMM backbone
    (spconv-CPU is not installable here) and, for the dense stage, torch's own CPU operators over the reference's module
    arithmetic (oracle/neck_head.py, oracle/pillars.py: the reference neck / head ARE torch.nn modules)."""

    SAMPLE = {"backbone": "1 scene: C/OpenMP port of numba voxelizer + spconv-v1 gather-GEMM backbone",
              "full": "1 scene: C/OpenMP port (voxelizer, sparse backbone, decode, rotated NMS, RoI) + torch-CPU S2D_RPN / "
                      "CenterHead (the reference's own nn arithmetic)",
              "pillar": "1 scene: C/OpenMP voxelizer + torch-CPU PFN / scatter+S2D / RPN / CenterHead + C NMS / RoI",
              "train": "1 scene forward only (teacher-shaped full forward); the CPU backward is not timed"}

    def __init__(self, name):
        import torch
        from oracle import ref_ops as R
        from sparse2dense_b200 import synth
        R.build()
        R.set_num_threads(host_cores())        # torchrun exports OMP_NUM_THREADS=1: use every host core explicitly
        torch.set_num_threads(host_cores())
        self.cores = R.num_threads()
        self.name = "full" if name == "train" else name
        self.timings = {}
        if self.name == "backbone":
            self.states = synth.backbone_state(0)
        else:
            self.states = cpu_states(self.name)

    def seconds(self, cloud):
        from oracle import full_forward as FF
        from oracle import backbone as OB
        from oracle import ref_ops as R
        from sparse2dense_b200 import synth
        t0 = time.perf_counter()
        if self.name == "backbone":
            v, c, n = R.points_to_voxel(cloud, synth.WAYMO_VOXEL, synth.WAYMO_RANGE, synth.WAYMO_MAX_POINTS, True,
                                        synth.WAYMO_MAX_VOXELS)
            feats = R.voxel_mean(v, n)
            coors = np.concatenate([np.zeros((len(c), 1), np.int32), c], 1)
            OB.backbone_forward(self.states, feats, coors, 1, (1504, 1504, 40))
        elif self.name == "full":
            FF.scene_forward(self.states, cloud, timings=self.timings)
        else:
            FF.pillar_scene_forward(self.states, cloud, timings=self.timings)
        return time.perf_counter() - t0


def cpu_states(name):
    """The same seeded weights as the GPU arm's synthetic models, as numpy state dicts, built WITHOUT a CUDA device (module
    construction is host-only; the .so is not needed for state dicts)."""
    import logging
    import torch
    from sparse2dense_b200 import registry, synth
    from sparse2dense_b200 import hotpath as HP
    f = lambda m, seed: {**{k: v.detach().numpy() for k, v in m.state_dict().items()},
                         **synth.random_module_state(m, seed)}
    if name == "full":
        neck = registry.build_neck(dict(logger=logging.getLogger("RPN"), **HP.FullForwardPath.NECK_CFG))
        head = registry.build_head(dict(**HP.FullForwardPath.HEAD_CFG))
        roi = registry.build_roi_head(dict(**HP.FullForwardPath.ROI_CFG))
        hs = f(head, 12)
        hs["tasks.0.hm.%d.bias" % (len(head.tasks[0].hm) - 1)] = np.full_like(hs["tasks.0.hm.%d.bias" % (len(head.tasks[0].hm) - 1)], -1.0)
        return dict(backbone=synth.backbone_state(0), neck=f(neck, 11), head=hs, roi=f(roi, 13))
    P = HP.PillarForwardPath
    reader = registry.build_reader(dict(type="PillarFeatureNet", num_filters=[64, 64], num_input_features=5,
                                        with_distance=False, voxel_size=P.VOXEL, pc_range=P.RANGE))
    bb = registry.build_backbone(dict(type="PointPillarsScatter_S2D", ds_factor=1))
    neck = registry.build_neck(dict(type="RPN", layer_nums=[3, 5, 5], ds_layer_strides=[1, 2, 2], ds_num_filters=[64, 128, 256],
                                    us_layer_strides=[1, 2, 4], us_num_filters=[128, 128, 128], num_input_features=64,
                                    logger=logging.getLogger("RPN")))
    head = registry.build_head(dict(type="CenterHead", in_channels=384,
                                    tasks=[dict(num_class=3, class_names=["VEHICLE", "PEDESTRIAN", "CYCLIST"])],
                                    dataset="waymo", weight=2, code_weights=[1.0] * 8,
                                    common_heads={"reg": (2, 2), "height": (1, 2), "dim": (3, 2), "rot": (2, 2)}))
    roi = registry.build_roi_head(dict(type="RoIHead", input_channels=384 * 5, code_size=7,
                                       model_cfg=dict(CLASS_AGNOSTIC=True, SHARED_FC=[256, 256], CLS_FC=[256, 256],
                                                      REG_FC=[256, 256], DP_RATIO=0.3)))
    hs = f(head, 34)
    key = "tasks.0.hm.%d.bias" % (len(head.tasks[0].hm) - 1)
    hs[key] = np.full_like(hs[key], -1.0)
    return dict(reader=f(reader, 31), backbone=f(bb, 32), neck=f(neck, 33), head=hs, roi=f(roi, 35))


def run_reference(args):
    """--impl reference: the reference's CPU algorithm on all host cores; each step = ONE scene of the workload (bounded
    sample).  Under torchrun rank 0 alone works."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from sparse2dense_b200 import synth
    cpu = CpuScene(args.config)
    clouds = synth.lidar_batch(1, 2)
    if args.warmup > 0:
        cpu.seconds(clouds[0])                 # one warm-up scene is enough for a CPU loop
    t0 = time.perf_counter()
    for i in range(args.steps):
        cpu.seconds(clouds[i % len(clouds)])
    total = time.perf_counter() - t0
    value = args.steps / total
    batch = args.batch or WORKLOADS[args.config]["batch"]
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args.config, batch, args.precision, {"sample": "one scene per step"}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cpu.cores, "kind": "port", "sample": CpuScene.SAMPLE[args.config],
                         "stage_seconds_last_scene": {k: round(v, 3) for k, v in cpu.timings.items()}},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit_line(line)


# ---------------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi samples during the timed region (B200_PROFILING.md clocks line)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu, self.mark_at = [], None, gpu_index, 0

    def mark(self):
        """Samples from here on belong to the timed regions (nvidia-smi needs a few hundred ms to start on an 8-GPU box,
        so it is launched before the warm-up)."""
        self.mark_at = len(self.rows)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = self.rows[self.mark_at:] or self.rows          # fall back to the warm-up samples (same workload)
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(names, r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


class ForwardWorkload:
    """backbone / full / pillar: weak scaling, every rank owns its own batch of scenes (seeds differ per rank); no
    data-path collective."""

    def __init__(self, name, batch, precision, dev, rank):
        import torch
        from sparse2dense_b200 import hotpath as HP, sharding, synth
        self.name, self.batch, self.dev = name, batch, dev
        if name == "backbone":
            self.path = HP.VoxelBackbonePath(state=synth.backbone_state(0), precision=precision, device=dev)
        elif name == "full":
            self.path = HP.FullForwardPath.synthetic(precision=precision, device=dev)
        else:
            self.path = HP.PillarForwardPath(precision=precision, device=dev)
        self.clouds = [synth.lidar_scene(seed) for seed in sharding.scene_seeds(1, batch, rank)]
        self.pts_host, self.offs = HP.concat_clouds(self.clouds, pin=True)
        self.pts_dev = self.pts_host.to(dev)
        if name == "backbone":
            mk = lambda: torch.empty((batch, 256, 188, 188), dtype=torch.float32).pin_memory()
        else:
            mk = lambda: (torch.empty((batch, 500, 7), dtype=torch.float32).pin_memory(),
                          torch.empty((batch, 500), dtype=torch.float32).pin_memory(),
                          torch.empty((batch, 500), dtype=torch.int32).pin_memory(),
                          torch.empty((batch,), dtype=torch.int32).pin_memory())
        self.out_host = [mk(), mk()]
        outs = self.out_host[0] if isinstance(self.out_host[0], tuple) else (self.out_host[0],)
        self.h2d = int(self.pts_host.numel() * 4)
        self.d2h = int(sum(o.numel() * o.element_size() for o in outs))
        self.output = ("dense BEV [B,256,188,188] fp32" if name == "backbone" else
                       "detections: boxes [B,500,7] f32, scores [B,500] f32, labels [B,500] i32, counts [B] i32")

    def step(self):
        return self.path.forward_points(self.pts_dev, self.offs)

    def host_step(self, i):
        return self.path.forward_host_async(self.pts_host, self.offs, self.out_host[i % 2])

    def detections(self):
        if self.name == "backbone":
            return None
        return [int(v) for v in self.out_host[0][3].tolist()]


class TrainWorkload:
    """configs[4]: one process per GPU, scenes sharded over ranks, SyncBatchNorm when N > 1 (tools/train.py:92-96), ONE flat
    NCCL all-reduce of the gradients per step (apis/train.py:297-303 wraps the student in DistributedDataParallel)."""

    def __init__(self, name, batch, precision, dev, rank, world):
        import torch
        from sparse2dense_b200 import synth
        from sparse2dense_b200.trainer import DistillTrainer
        self.name, self.batch, self.dev = name, batch, dev
        teacher, student = synth.build_distill_models(dev, precision)
        self.sync_bn = world > 1
        if self.sync_bn:
            student = torch.nn.SyncBatchNorm.convert_sync_batchnorm(student)
        self.trainer = DistillTrainer(teacher, student, total_steps=1000)
        self.trainer.events = []
        self.ex = synth.distill_example(batch, cfg=1 + rank, device=dev)
        self.ex_host = {}
        for k, v in self.ex.items():
            if torch.is_tensor(v) and v.is_cuda:
                self.ex_host[k] = v.cpu().pin_memory()
            elif isinstance(v, list) and v and torch.is_tensor(v[0]) and v[0].is_cuda:
                self.ex_host[k] = [t.cpu().pin_memory() for t in v]
        self.loss_host = torch.zeros((1,), dtype=torch.float32).pin_memory()
        flat = [t for v in self.ex_host.values() for t in (v if isinstance(v, list) else [v])]
        self.h2d = int(sum(t.numel() * t.element_size() for t in flat))
        self.d2h = 4
        self.output = "loss scalar (and the updated student weights on the device)"
        self.params = int(self.trainer.opt.flat_p.numel())

    def step(self):
        return self.trainer.step(self.ex)

    def host_step(self, i):
        import torch
        ex = dict(self.ex)
        for k, v in self.ex_host.items():
            ex[k] = [t.to(self.dev, non_blocking=True) for t in v] if isinstance(v, list) else v.to(self.dev, non_blocking=True)
        log = self.trainer.step(ex)
        self.loss_host.copy_(log["loss"].reshape(1), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return ev

    def detections(self):
        return None


def time_workload(wl, steps, warmup, flush, barrier, sampler=None):
    """W warm-up steps, then K device-resident steps (L2 flushed before each, CUDA events per step), then K end-to-end steps
    (pinned host inputs, results into pinned host memory, all copies inside the timed region).  -> dict of raw numbers."""
    import torch
    from sparse2dense_b200 import ops
    for _ in range(warmup):
        wl.step()
    barrier()
    if sampler is not None:
        sampler.mark()
    ops.KERNEL_EVENTS = []
    if hasattr(wl, "trainer"):
        wl.trainer.events = []
    launches0 = ops.kernel_launches()
    evs = []
    barrier()
    t_wall0 = time.perf_counter()
    for _ in range(steps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        wl.step()
        e1.record()
        evs.append((e0, e1))
    barrier()
    wall = time.perf_counter() - t_wall0
    launches = ops.kernel_launches() - launches0
    kernel_events, ops.KERNEL_EVENTS = ops.KERNEL_EVENTS, None
    train_events = None
    if hasattr(wl, "trainer"):
        train_events, wl.trainer.events = wl.trainer.events, None
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)

    # end to end: the public call is the pipelined one (forward_host_async / trainer.step on a host example): the D2H of
    # step i runs on a copy stream under step i+1; the timed region is the whole K-step loop, closed only when the LAST
    # result has landed in host memory.  The L2 flush writes are timed separately and subtracted (they are bench hygiene, not
    # part of the call; ~0.07 ms each).
    for i in range(2):
        wl.host_step(i).synchronize()
    barrier()
    main = torch.cuda.current_stream(wl.dev)
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    flush_evs = []
    done = None
    for i in range(steps):
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(); flush.fill_(1); f1.record()
        flush_evs.append((f0, f1))
        done = wl.host_step(i)
    main.wait_event(done)
    e1.record()
    barrier()
    e2e_wall = time.perf_counter() - t0
    flush_ms = sum(a.elapsed_time(b) for a, b in flush_evs)
    e2e_ms = e0.elapsed_time(e1) - flush_ms
    return dict(dev_ms=dev_ms, e2e_ms=e2e_ms, wall=wall, e2e_wall=e2e_wall, launches=launches, kernel_events=kernel_events,
                train_events=train_events, flush_ms_subtracted=flush_ms)


def conv_groups(kernel_events, steps):
    groups = {}
    for key, a, b in kernel_events:
        cin, cout, K, has_res, n_in, n_out, prec, kind = key
        g = groups.setdefault((cin, cout, K, kind, n_out), {"ms": 0.0, "n": 0, "n_in": n_in, "n_out": n_out, "res": 0,
                                                             "prec": prec})
        g["ms"] += a.elapsed_time(b)
        g["n"] += 1
        g["res"] += 1 if has_res else 0
    return groups


def rooflines(wl, raw, steps, world, dev):
    """`roofline` for the dominant SPARSE conv group (the kernel north_star's HBM target names) with the dominant DENSE
    (neck / head) group as `dense`, both timed live with CUDA events around every launch of the timed region."""
    from sparse2dense_b200 import ops
    groups = conv_groups(raw["kernel_events"], steps)
    pk = peaks()
    tf32_peak = measure_tf32_peak(dev)
    dev_ms = raw["dev_ms"]
    pairs_cache = {}

    def describe(key, g):
        cin, cout, K, kind, n_out = key
        avg_ms = g["ms"] / g["n"]
        n_in = g["n_in"]
        res_frac = g["res"] / g["n"]
        if kind == "sparse":
            if (cin, cout, K) not in pairs_cache:
                pairs_cache[(cin, cout, K)] = count_pairs_for(wl.path, wl.pts_dev, wl.offs, (cin, cout, K))
            pairs, live_rows = pairs_cache[(cin, cout, K)]
            # compulsory bytes per launch: in rows + out rows + weights (+ residual on the launches that have one) + the
            # neighbour table once per indice_key (shared by 4 SubM launches); activations counted as fp32 (4 B / channel:
            # the split row is 2 x bf16 = 4 B as well)
            bytes_launch = (n_in * cin + n_out * cout + K * cin * cout + res_frac * n_out * cout) * 4 + n_out * K * 4 / 4
        else:
            pairs = n_out * K                                  # dense grid: every neighbour exists up to the border
            live_rows = None
            bytes_launch = (n_in * cin + n_out * cout + K * cin * cout + res_frac * n_out * cout) * 4
        flops = 2.0 * pairs * cin * cout
        tfl = flops / (avg_ms * 1e-3) / 1e12
        gbs = bytes_launch / (avg_ms * 1e-3) / 1e9
        prec = g["prec"]
        eff = ops.PRECISION_NAMES_INV.get(prec, str(prec)) if hasattr(ops, "PRECISION_NAMES_INV") else str(prec)
        return dict(kernel=f"conv {cin}->{cout} K={K} ({kind}, {n_out} rows; precision {eff})", avg_launch_ms=avg_ms,
                    launches_per_step=g["n"] / steps, share_of_step=(g["ms"] / dev_ms) if world == 1 else None,
                    algorithmic_flops=flops, algorithmic_bytes=bytes_launch, pairs=pairs,
                    executed_over_algorithmic_rows=(live_rows if live_rows is not None else n_out * K) / max(pairs, 1),
                    tflops=tfl, gbs=gbs)

    def pick(kind):
        ks = [k for k in groups if k[3] == kind]
        if not ks:
            return None
        k = max(ks, key=lambda kk: groups[kk]["ms"])
        return k, describe(k, groups[k])

    out = None
    sp, de = pick("sparse"), pick("dense")
    peak_note = ("tensor peak = dense TF32 GEMM 8192^3 measured live with cuBLAS (%.1f TFLOP/s; bf16 burst / 2 from %s = %.1f); "
                 "hbm peak from %s" % (tf32_peak, pk["source"], pk["bf16"] / 2.0, pk["source"]))

    def l2_stream(d, key):
        """What the gather-GEMM pulls from L2 into the SMs per launch -- the stream that binds conv_bf2: every real
        (row, offset) pair reads Cin x 4 B of split rows, every executed block one [128-channel x 128 B] weight tile shared by
        the tiles of a group (2 tiles at 128 output channels, else 4) -- against the L2 -> SM rate measured by tools/l2_probe.cu."""
        cin, cout, K, kind, n_out = key
        try:
            with open(os.path.join(ROOT, "profiles", "r2_l2_probe.json")) as f:
                probe = json.load(f)
        except Exception:
            probe = {"l2_to_sm_tbs_resident": 20.3, "l2_to_sm_tbs_large_working_set": 14.5, "source": "fallback constants"}
        cb = 128 if cout % 128 == 0 else 64 if cout % 64 == 0 else 32 if cout % 32 == 0 else 16
        tiles_per_group = 2 if cb == 128 else 4
        chunk_k = 2 if cin == 16 else 1                       # two offsets share a 128 B row at 16 channels
        executed_rows = d["executed_over_algorithmic_rows"] * d["pairs"]
        tile_blocks = executed_rows / 128.0 * max(1, cin // 32) / chunk_k
        weight_bytes = tile_blocks / tiles_per_group * cb * 128.0 * (cout // cb)
        gather_bytes = d["pairs"] * cin * 4.0
        tbs = (gather_bytes + weight_bytes) / (d["avg_launch_ms"] * 1e-3) / 1e12
        return {"achieved_tbs": tbs, "peak_tbs": probe["l2_to_sm_tbs_resident"], "frac": tbs / probe["l2_to_sm_tbs_resident"],
                "peak_tbs_large_working_set": probe["l2_to_sm_tbs_large_working_set"],
                "gather_bytes_per_launch": gather_bytes, "weight_tile_bytes_per_launch": weight_bytes,
                "source": probe.get("source", "")}

    def executed_bf16(d, key):
        """The MMAs the kernel really issues: three kind::f16 (BF16) MMAs per product over the EXECUTED rows (whole 128-row tiles
        of live (tile, offset) pairs), against the dense BF16 GEMM rate of MEASURED_PEAKS.json.  `frac` here says how busy the
        tensor cores are; `roofline.frac` above charges the 1.5 TF32-pass cost of the fp32-level product and the dead rows of
        live tiles to the kernel."""
        cin, cout, K, kind, n_out = key
        fl = 3.0 * 2.0 * d["executed_over_algorithmic_rows"] * d["pairs"] * cin * cout
        tfl = fl / (d["avg_launch_ms"] * 1e-3) / 1e12
        return {"achieved_tflops": tfl, "peak_tflops": pk["bf16"], "frac": tfl / pk["bf16"],
                "note": "3 BF16 MMAs per product x executed rows; peak = dense bf16 GEMM (burst) from " + pk["source"]}

    def tensor_obj(d, key):
        cin, cout, K, kind, n_out = key
        return {"bound": "tensor", "achieved": d["tflops"], "peak": tf32_peak, "unit": "TFLOP/s", "frac": d["tflops"] / tf32_peak,
                "traffic": ncu_traffic(f"conv_bf2_kernel {cin}->{cout} K={K} {kind}"), "kernel": d["kernel"],
                "avg_launch_ms": d["avg_launch_ms"], "launches_per_step": d["launches_per_step"],
                "share_of_step": d["share_of_step"], "algorithmic_flops_per_launch": d["algorithmic_flops"],
                "algorithmic_bytes_per_launch": d["algorithmic_bytes"],
                "executed_over_algorithmic_rows": d["executed_over_algorithmic_rows"],
                "hbm": {"achieved_gbs": d["gbs"], "peak_gbs": pk["hbm"], "frac": d["gbs"] / pk["hbm"]},
                "l2_stream": l2_stream(d, key), "executed_bf16": executed_bf16(d, key)}

    if sp is not None:
        out = tensor_obj(sp[1], sp[0])
        out["peak_source"] = peak_note
        out["note"] = ("achieved = algorithmic flops 2*P*Cin*Cout (real neighbour pairs P only) / live launch time; the "
                       "BF16-pair kernel (conv_bf2.cu) issues 3 BF16 MMAs per product = 1.5 TF32-pass equivalents")
        # every sparse layer group against the HBM roof (north_star's target is stated on HBM)
        table = {}
        for k in sorted(groups):
            if k[3] != "sparse":
                continue
            d = describe(k, groups[k])
            table[f"{k[0]}->{k[1]} K={k[2]}"] = {"avg_launch_ms": round(d["avg_launch_ms"], 4), "launches_per_step": d["launches_per_step"],
                                                "hbm_frac": round(d["gbs"] / pk["hbm"], 4), "tensor_frac": round(d["tflops"] / tf32_peak, 4),
                                                "gbs": round(d["gbs"], 1), "tflops": round(d["tflops"], 1),
                                                "executed_over_algorithmic_rows": round(d["executed_over_algorithmic_rows"], 2)}
        out["sparse_layers"] = table
    if de is not None:
        dense = tensor_obj(de[1], de[0])
        tot_ms = sum(g["ms"] for k, g in groups.items() if k[3] == "dense")
        tot_fl = sum(2.0 * k[4] * k[2] * k[0] * k[1] * g["n"] for k, g in groups.items() if k[3] == "dense")
        dense["all_dense_convs"] = {"ms_per_step": tot_ms / steps, "tflops": tot_fl / (tot_ms * 1e-3) / 1e12,
                                    "frac": tot_fl / (tot_ms * 1e-3) / 1e12 / tf32_peak}
        if out is None:
            out = dense
            out["peak_source"] = peak_note
        else:
            out["dense"] = dense
    per_group = {f"{k[0]}->{k[1]} K={k[2]} {k[3]} rows={k[4]}": round(g["ms"] / steps, 4) for k, g in sorted(groups.items())}
    return out, per_group


def count_pairs_for(path, pts_dev, offs, shape_key):
    """(rulebook pair count P, rows the tile kernel executes = 128 x live (tile, offset) pairs of the grouped rulebook) of the
    sparse layer group (Cin,Cout,K) on this rank's batch (untimed)."""
    import torch
    from sparse2dense_b200 import ops, spconv
    cin, cout, K = shape_key
    bb = path.backbone
    vb = path.generator.generate_batch(pts_dev, offs, want_voxels=False, mean_channels=5)
    n = vb.n
    x = spconv.SparseConvTensor(vb.mean_buffer[:n], vb.coors_buffer[:n], (41, 1504, 1504), len(offs) - 1)
    plan = spconv.plan_coords(x, [bb.conv2[0], bb.conv3[0], bb.conv4[0], bb.extra_conv[0]])
    stages = [(x.indices, x.index(), None)]
    for sc in plan:
        stages.append((sc.coors, sc.index, sc))

    def live(masks):
        m = masks.to(torch.int64) & 0x7FFFFFF
        bits = sum(((m >> b) & 1) for b in range(27))
        return int(bits.sum().item()) * 128

    chans = [16, 32, 64, 128]
    downs = {(16, 32): 1, (32, 64): 2, (64, 128): 3}
    if (cin, cout) in downs and K == 27:
        s = downs[(cin, cout)]
        m = [bb.conv2[0], bb.conv3[0], bb.conv4[0]][s - 1]
        _, pairs = ops.rulebook_sparse(stages[s][0], stages[s - 1][1], m.kernel_size, m.stride, m.padding, count_pairs=True)
        masks = ops.rulebook_sparse_grouped(stages[s][0], stages[s - 1][1], m.stride, m.padding)[2] if spconv.GROUP_ROWS >= 2 \
            else ops.table_tile_masks(ops.rulebook_sparse(stages[s][0], stages[s - 1][1], m.kernel_size, m.stride, m.padding),
                                      stages[s][0].shape[0])
        return int(pairs.item()), live(masks)
    if K == 3:
        m = bb.extra_conv[0]
        tbl, pairs = ops.rulebook_sparse(stages[4][0], stages[3][1], m.kernel_size, m.stride, m.padding, count_pairs=True)
        return int(pairs.item()), live(ops.table_tile_masks(tbl, stages[4][0].shape[0]))
    s = chans.index(cout) if (K == 27 and cout in chans and cin in (cout, 5)) else 0     # SubM group of stage s (5->16: stage 0)
    tbl, pairs = ops.rulebook_subm(stages[s][0], stages[s][1], 3, count_pairs=True)
    masks = ops.rulebook_subm_grouped(stages[s][0], stages[s][1])[2] if spconv.GROUP_ROWS >= 1 else \
        ops.table_tile_masks(tbl, stages[s][0].shape[0])
    return int(pairs.item()), live(masks)


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from sparse2dense_b200 import ops, sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    precision = ops.PRECISION_NAMES[args.precision]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    warmup = max(args.warmup, 3)                 # contract: W >= 3; the value used is the one reported

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def make(name, batch):
        if name == "train":
            return TrainWorkload(name, batch, precision, dev, rank, world)
        return ForwardWorkload(name, batch, precision, dev, rank)

    def summarize(name, wl, raw, steps, with_roofline):
        """Reduce over ranks (slowest rank sets the time) and build the result object of one workload (rank 0 keeps it)."""
        dev_ms, e2e_ms = sharding.max_over_ranks([raw["dev_ms"], raw["e2e_ms"]], device=dev)
        (launches,) = sharding.sum_over_ranks([raw["launches"]], device=dev)
        scenes = wl.batch * world * steps
        obj = {"workload": WORKLOADS[name]["text"], "batch_per_gpu": wl.batch, "global_batch": wl.batch * world, "steps": steps,
               "value": scenes / (dev_ms * 1e-3), "unit": UNIT, "ms_per_step": dev_ms / steps,
               "e2e": {"value": scenes / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": wl.h2d, "d2h_bytes_per_step": wl.d2h,
                       "ms_per_step": e2e_ms / steps, "wall_ms_per_step": 1e3 * raw["e2e_wall"] / steps, "output": wl.output,
                       "l2_flush_ms_subtracted_per_step": raw["flush_ms_subtracted"] / steps},
               "gpu_launches": int(launches), "wall_ms_per_step": 1e3 * raw["wall"] / steps}
        if raw["train_events"]:
            agg = {}
            for nm, a, b in raw["train_events"]:
                agg[nm] = agg.get(nm, 0.0) + a.elapsed_time(b)
            vals = sharding.max_over_ranks([agg.get(k, 0.0) for k in ("forward_backward", "all_reduce", "optimizer")], device=dev)
            obj["train_breakdown_ms_per_step"] = {k: v / steps for k, v in zip(("forward_backward", "all_reduce", "optimizer"), vals)}
            obj["all_reduce_bytes"] = wl.params * 4
            obj["sync_bn"] = wl.sync_bn
            obj["parallelism"] = f"DDP over {world} rank(s): one flat NCCL all-reduce of {wl.params * 4 / 1e6:.1f} MB gradients per step"
        if rank == 0 and with_roofline and raw["kernel_events"] and not hasattr(wl, "trainer"):
            obj["roofline"], obj["kernel_ms_per_step"] = rooflines(wl, raw, steps, world, dev)
        det = wl.detections()
        if det is not None:
            obj["detections_per_scene_rank0"] = det
        return obj

    sampler = ClockSampler(local)
    sampler.start()
    name = args.config
    batch = args.batch or WORKLOADS[name]["batch"]
    wl = make(name, batch)
    raw = time_workload(wl, args.steps, warmup, flush, barrier, sampler)
    clocks = sampler.stop()
    print(f"[bench rank {rank}] {name}: device {raw['dev_ms'] / args.steps:.3f} ms/step, e2e {raw['e2e_ms'] / args.steps:.3f} ms/step "
          f"(wall {1e3 * raw['e2e_wall'] / args.steps:.3f})", file=sys.stderr, flush=True)
    head = summarize(name, wl, raw, args.steps, True)
    clouds0 = getattr(wl, "clouds", None)
    del wl, raw
    torch.cuda.empty_cache()

    also = {}
    others = [] if args.also == "none" else [n for n in (args.also.split(",") if args.also != "default" else
                                                            ["backbone", "pillar", "train"]) if n != name]
    for other in others:
        o_steps = max(3, min(args.steps, 5 if other == "train" else 10))
        try:
            w2 = make(other, WORKLOADS[other]["batch"])
            r2 = time_workload(w2, o_steps, 3, flush, barrier)
            also[other] = summarize(other, w2, r2, o_steps, other != "train")
            print(f"[bench rank {rank}] {other}: device {r2['dev_ms'] / o_steps:.3f} ms/step, e2e {r2['e2e_ms'] / o_steps:.3f} ms/step",
                  file=sys.stderr, flush=True)
            del w2, r2
        except Exception as exc:                              # a sub-measurement must not take the headline down
            if world > 1:
                raise                                           # ... but ranks must not diverge around collectives
            also[other] = {"error": f"{type(exc).__name__}: {exc}"}
        torch.cuda.empty_cache()

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from sparse2dense_b200 import synth
            c = CpuScene(name)
            cloud = clouds0[0] if clouds0 is not None else synth.lidar_scene(1)
            sec = c.seconds(cloud)
            cpu = {"value": 1.0 / sec, "unit": UNIT, "cores": c.cores, "kind": "port",
                   "sample": "scene 0 of the batch, once; " + CpuScene.SAMPLE[name],
                   "stage_seconds": {k: round(v, 3) for k, v in c.timings.items()}}
        sparse_prec = "bf16x3-split (fp32-level: 3 BF16 MMAs per product, fp32 accumulate)" if args.precision == "auto" else args.precision
        line = {
            "metric": METRIC, "value": head["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": head["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if args.precision == "fp32" else "bf16x3", "data": "synthetic",
            "config": config_dict(name, batch, sparse_prec, {
                "global_batch": batch * world,
                "parallelism": head.get("parallelism", f"scenes sharded over {world} rank(s), no data-path collective")}),
            "clocks": clocks, "e2e": head["e2e"], "gpu_launches": head["gpu_launches"],
            "wall_ms_per_step": head["wall_ms_per_step"], "roofline": head.get("roofline"), "cpu_baseline": cpu,
            "kernel_ms_per_step": head.get("kernel_ms_per_step"),
        }
        for k in ("detections_per_scene_rank0", "train_breakdown_ms_per_step", "all_reduce_bytes", "sync_bn"):
            if k in head:
                line[k] = head[k]
        if also:
            line["also"] = also
        emit_line(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_REAL_STDOUT = None


def quiet_stdout():
    """Send everything libraries print on fd 1 (e.g. NCCL's version banner) to stderr; the JSON line is the only
    thing written to the real stdout (emit_line)."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit_line(obj):
    data = (json.dumps(obj) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="full", choices=sorted(WORKLOADS))
    ap.add_argument("--also", default="default", help="comma list of further workloads measured briefly into `also` "
                                                      "(default: the other three; 'none' to skip)")
    ap.add_argument("--batch", type=int, default=0, help="scenes per GPU per step (default: the workload's BASELINE batch)")
    ap.add_argument("--precision", default=os.environ.get("S2D_PRECISION", "auto"),
                    choices=["fp32", "tf32", "tf32x3", "tf32_bf16c", "bf16x2", "auto"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
