#!/usr/bin/env python
"""bench.py -- scenes/s of the hot path on Waymo-shaped synthetic clouds (BASELINE.json metric).

Workload (BASELINE.json configs[1]): voxelize -> reader mean -> SpMiddleResNetFHD sparse backbone ->
dense BEV, batch = 4 synthetic ~180 k-point clouds per GPU on the 1504x1504x40 grid.

    python bench.py [--gpus N] [--steps K] [--warmup W]          our arm (CUDA, sm_100a)
    python bench.py --impl reference ...                          the CPU arm (oracle port, host cores)

One JSON line on stdout (rank 0).  Keys: see the task contract; `value` = device-resident
throughput, `e2e` = through the public API with pinned HOST buffers (H2D + D2H in the timed
region), `roofline` = dominant kernel timed live with CUDA events, `cpu_baseline` = the oracle port
on the host cores for a bounded sample of the same workload.
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
WORKLOAD = "voxelize+SpMiddleResNetFHD backbone -> BEV, batch=4 x ~180k-pt synthetic Waymo clouds, grid 1504x1504x40"
FP32_SIMT_PEAK_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12          # nominal; not in MEASURED_PEAKS.json


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(hbm=float(p["hbm_gbs"]), bf16=float(p["bf16_tflops"]),
                    bf16_sustained=float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), source="measured")
    return dict(hbm=6650.0, bf16=1590.0, bf16_sustained=1400.0, source="fallback")


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


def config_dict(args, extra=None):
    cfg = {"workload": WORKLOAD, "batch_per_gpu": args.batch, "points_in_range": "~180k/scene",
           "voxel_size": [0.1, 0.1, 0.15], "max_voxels": 150000, "weights": "seeded random (synth.backbone_state(0))",
           "l2": "flushed between timed steps (256 MiB write)", "precision": args.precision}
    if extra:
        cfg.update(extra)
    return cfg


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores (bounded sample: one scene per step)
# ---------------------------------------------------------------------------------------------
def cpu_scene_seconds(cloud, state, repeat=1):
    from oracle import backbone as OB
    from oracle import ref_ops as R
    from sparse2dense_b200 import synth
    best = None
    for _ in range(repeat):
        t0 = time.perf_counter()
        v, c, n = R.points_to_voxel(cloud, synth.WAYMO_VOXEL, synth.WAYMO_RANGE, synth.WAYMO_MAX_POINTS, True,
                                    synth.WAYMO_MAX_VOXELS)
        feats = R.voxel_mean(v, n)
        coors = np.concatenate([np.zeros((len(c), 1), np.int32), c], 1)
        OB.backbone_forward(state, feats, coors, 1, (1504, 1504, 40))
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_reference(args):
    """--impl reference: the reference's CPU algorithm (oracle port; spconv itself is not installable,
    see DESIGN.md) on all host cores; each step = one scene of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref_ops as R
    from sparse2dense_b200 import synth
    R.build()
    R.set_num_threads(host_cores())        # torchrun exports OMP_NUM_THREADS=1: use every host core explicitly
    cores = R.num_threads()
    state = synth.backbone_state(0)
    clouds = synth.lidar_batch(1, min(args.batch, 2))
    if args.warmup > 0:
        cpu_scene_seconds(clouds[0], state)                 # one warm-up scene is enough for a CPU loop
    t0 = time.perf_counter()
    for i in range(args.steps):
        cpu_scene_seconds(clouds[i % len(clouds)], state)
    total = time.perf_counter() - t0
    value = args.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(args, {"sample": "one scene per step (voxelize + reader + backbone)"}),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "1 scene per step: C/OpenMP oracle port of numba voxelizer + spconv-v1 gather-GEMM "
                                   "backbone (spconv-CPU is not installable here)"},
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


def run_gpu(args):
    import torch
    import torch.distributed as dist
    from sparse2dense_b200 import ops, sharding, synth
    from sparse2dense_b200.hotpath import VoxelBackbonePath, concat_clouds

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
    path = VoxelBackbonePath(state=synth.backbone_state(0), precision=precision, device=dev)
    # weak scaling: every rank owns its own batch of scenes (seeds differ per rank); no data-path collective
    clouds = [synth.lidar_scene(seed) for seed in sharding.scene_seeds(1, args.batch, rank)]
    pts_host, offs = concat_clouds(clouds, pin=True)
    pts_dev = pts_host.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    bev_host = torch.empty((args.batch, 256, 188, 188), dtype=torch.float32, pin_memory=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        path.forward_points(pts_dev, offs)
    barrier()

    # ---- device-resident timing: K steps, L2 flushed before each, CUDA events per step -------
    sampler.mark()
    ops.KERNEL_EVENTS = []
    launches0 = ops.kernel_launches()
    evs = []
    barrier()
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        path.forward_points(pts_dev, offs)
        e1.record()
        evs.append((e0, e1))
    barrier()
    wall = time.perf_counter() - t_wall0
    launches = ops.kernel_launches() - launches0
    kernel_events, ops.KERNEL_EVENTS = ops.KERNEL_EVENTS, None
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)

    # ---- end to end: pinned host points in, pinned host BEV out, every copy inside the timed region.  The public
    # call is the pipelined one (forward_host_async): the D2H of step i runs on a copy stream under step i+1; the
    # timed region is the whole K-step loop, closed only when the LAST result has landed in host memory.
    bev_hosts = [bev_host, torch.empty_like(bev_host).pin_memory()]
    for i in range(2):
        path.forward_host_async(pts_host, offs, bev_hosts[i]).synchronize()
    barrier()
    main = torch.cuda.current_stream(dev)
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    flush_ms_evs = []
    for i in range(args.steps):
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record(); flush.fill_(1); f1.record()                  # L2 flush: timed separately and subtracted
        flush_ms_evs.append((f0, f1))
        done = path.forward_host_async(pts_host, offs, bev_hosts[i % 2])
    main.wait_event(done)
    e1.record()
    barrier()
    e2e_wall = time.perf_counter() - t0
    e2e_ms = e0.elapsed_time(e1) - sum(a.elapsed_time(b) for a, b in flush_ms_evs)
    clocks = sampler.stop()                                   # samples cover the device-resident AND the end-to-end region

    print(f"[bench rank {rank}] device {dev_ms / args.steps:.3f} ms/step, e2e {e2e_ms / args.steps:.3f} ms/step "
          f"(wall {1e3 * e2e_wall / args.steps:.3f})", file=sys.stderr, flush=True)
    dev_ms, e2e_ms = sharding.max_over_ranks([dev_ms, e2e_ms], device=dev)     # slowest rank sets the time
    (launches,) = sharding.sum_over_ranks([launches], device=dev)

    # ---- per-kernel accounting for the roofline object (rank 0) ------------------------------
    if rank == 0:
        groups = {}
        for key, a, b in kernel_events:
            g = groups.setdefault(key[:4] + (key[6],), {"ms": 0.0, "n": 0, "n_in": key[4], "n_out": key[5]})
            g["ms"] += a.elapsed_time(b)
            g["n"] += 1
        # algorithmic work needs the pair counts: rebuild the rulebooks once, outside any timed region
        pk = max(groups, key=lambda k: groups[k]["ms"])
        cin, cout, K, has_res, prec = pk
        dom = [g for k, g in groups.items() if k[:3] == (cin, cout, K)]
        dom_ms = sum(g["ms"] for g in dom)
        dom_n = sum(g["n"] for g in dom)
        n_in, n_out = groups[pk]["n_in"], groups[pk]["n_out"]
        # mean per-launch compulsory bytes over the launches of this (Cin,Cout,K) group:
        #   in rows + out rows + weights (+ residual on half the SubM launches) + table once per 4 launches
        res_frac = sum(g["n"] for k, g in groups.items() if k[:3] == (cin, cout, K) and k[3]) / max(dom_n, 1)
        bytes_launch = (n_in * cin + n_out * cout + K * cin * cout + res_frac * n_out * cout) * 4 + n_out * K * 4 / 4
        pairs = count_pairs_for(path, pts_dev, offs, (cin, cout, K))
        flops_launch = 2.0 * pairs * cin * cout
        avg_ms = dom_ms / max(dom_n, 1)
        pk_peaks = peaks()
        gbs = bytes_launch / (avg_ms * 1e-3) / 1e9
        tfl = flops_launch / (avg_ms * 1e-3) / 1e12
        if prec == ops.PRECISION_FP32:
            roof = {"bound": "hbm", "achieved": gbs, "peak": pk_peaks["hbm"], "unit": "GB/s",
                    "frac": gbs / pk_peaks["hbm"], "traffic": None,
                    "kernel": f"spconv_simt_kernel<{cin},{cout}> (K={K})", "launches_per_step": dom_n / args.steps,
                    "avg_launch_ms": avg_ms, "share_of_step": dom_ms / dev_ms if world == 1 else None,
                    "peak_source": pk_peaks["source"],
                    "note": "fp32 FFMA path: this layer is compute-bound on the CUDA-core pipe, not HBM",
                    "fp32_simt": {"achieved_tflops": tfl, "nominal_peak_tflops": FP32_SIMT_PEAK_TFLOPS,
                                  "frac": tfl / FP32_SIMT_PEAK_TFLOPS}}
        else:
            # the kernel is timed inside the step (events around every launch of a ~10 ms step), so the sustained
            # cuBLAS figure is the denominator (B200_PROFILING.md); TF32 runs at half the bf16 rate
            tf32_peak = pk_peaks["bf16_sustained"] / 2.0
            if prec == ops.PRECISION_AUTO:                     # what the library resolves AUTO to for this shape
                prec = ops.PRECISION_TF32_BF16C if cout % 128 == 0 else ops.PRECISION_TF32X3
            passes = {ops.PRECISION_TF32: 1, ops.PRECISION_TF32_BF16C: 2, ops.PRECISION_TF32X3: 3}[prec]
            mode_note = {1: "", 2: "; per K-slice one TF32 MMA + BF16 correction MMAs costing one more TF32 pass "
                                   "(TF32 + BF16-correction mode, fp32-level accuracy)",
                         3: " and 3 MMAs per K-slice in the error-compensated TF32x3 mode"}[passes]
            roof = {"bound": "tensor", "achieved": tfl, "peak": tf32_peak, "unit": "TFLOP/s", "frac": tfl / tf32_peak,
                    "traffic": ncu_traffic(f"spconv_tc_kernel<{cout},{passes}> Cin={cin}"),
                    "algorithmic_bytes": bytes_launch, "kernel": f"spconv_tc_kernel<{cin},{cout}> (K={K})",
                    "launches_per_step": dom_n / args.steps, "avg_launch_ms": avg_ms,
                    "share_of_step": dom_ms / dev_ms if world == 1 else None,
                    "peak_source": pk_peaks["source"] + " bf16 sustained / 2 (TF32 runs at half the bf16 rate; burst / 2 = %.1f)" % (pk_peaks["bf16"] / 2.0),
                    "note": "achieved = algorithmic flops 2*P*Cin*Cout (real neighbour pairs only); the kernel "
                            "executes dense 128-row tiles (zero rows for missing neighbours)"
                            + mode_note,
                    "executed_tflops_tf32_equiv": tfl * (n_out * K / max(pairs, 1)) * passes,
                    "hbm": {"achieved_gbs": gbs, "frac": gbs / pk_peaks["hbm"]}}
        per_group = {f"{k[0]}->{k[1]} K={k[2]}{' +res' if k[3] else ''}": round(g["ms"] / args.steps, 4)
                     for k, g in sorted(groups.items())}

        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            from oracle import ref_ops as R
            R.build()
            R.set_num_threads(host_cores())
            sec = cpu_scene_seconds(clouds[0], synth.backbone_state(0))
            cpu = {"value": 1.0 / sec, "unit": UNIT, "cores": R.num_threads(), "kind": "port",
                   "sample": "scene 0 of the batch, once (voxelize + reader + backbone), C/OpenMP oracle port"}

        scenes = args.batch * world * args.steps
        line = {
            "metric": METRIC, "value": scenes / (dev_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32" if prec == ops.PRECISION_FP32 else "tf32",
            "data": "synthetic", "config": config_dict(args, {"global_batch": args.batch * world,
                                                              "parallelism": f"scenes sharded over {world} rank(s), no data-path collective"}),
            "clocks": clocks,
            "e2e": {"value": scenes / (e2e_ms * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(pts_host.numel() * 4),
                    "d2h_bytes_per_step": int(bev_host.numel() * 4), "ms_per_step": e2e_ms / args.steps,
                    "wall_ms_per_step": 1e3 * e2e_wall / args.steps},
            "gpu_launches": int(launches), "wall_ms_per_step": 1e3 * wall / args.steps,
            "roofline": roof, "cpu_baseline": cpu, "kernel_ms_per_step": per_group,
        }
        emit_line(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def count_pairs_for(path, pts_dev, offs, shape_key):
    """Rulebook pair count P of the layer group (Cin,Cout,K) on this rank's batch (untimed)."""
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
    chans = [16, 32, 64, 128]
    if K == 27 and cin == cout and cin in chans:                       # SubM group of stage s
        s = chans.index(cin)
        _, pairs = ops.rulebook_subm(stages[s][0], stages[s][1], 3, count_pairs=True)
        return int(pairs.item())
    downs = {(16, 32): 1, (32, 64): 2, (64, 128): 3}
    if (cin, cout) in downs and K == 27:
        s = downs[(cin, cout)]
        m = [bb.conv2[0], bb.conv3[0], bb.conv4[0]][s - 1]
        _, pairs = ops.rulebook_sparse(stages[s][0], stages[s - 1][1], m.kernel_size, m.stride, m.padding,
                                       count_pairs=True)
        return int(pairs.item())
    if K == 3:
        m = bb.extra_conv[0]
        _, pairs = ops.rulebook_sparse(stages[4][0], stages[3][1], m.kernel_size, m.stride, m.padding,
                                       count_pairs=True)
        return int(pairs.item())
    _, pairs = ops.rulebook_subm(stages[0][0], stages[0][1], 3, count_pairs=True)
    return int(pairs.item())


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
    ap.add_argument("--batch", type=int, default=4, help="scenes per GPU per step")
    ap.add_argument("--precision", default=os.environ.get("S2D_PRECISION", "auto"), choices=["fp32", "tf32", "tf32x3", "tf32_bf16c", "auto"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
