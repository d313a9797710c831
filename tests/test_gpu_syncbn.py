"""SyncBatchNorm over rows ("SyncBN stats over NCCL"): two ranks, each with a different number of
rows, must reproduce the single-process batch-statistics BatchNorm over the concatenated rows (float64 torch) in outputs,
running statistics, input gradients, and -- summed over the ranks -- the affine gradients.  The two ranks share cuda:0 and
talk over gloo so that the test runs on a one-GPU box; the product code path is the same one all-reduce per pass that
NCCL carries in `tools/bench_train.py --sync-bn` under torchrun."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N, C, SPLIT = 3000, 48, 1100


def _data():
    g = torch.Generator().manual_seed(12)
    x = torch.randn(N, C, generator=g, dtype=torch.float64) * 1.5 + 0.3
    res = torch.randn(N, C, generator=g, dtype=torch.float64)
    dy = torch.randn(N, C, generator=g, dtype=torch.float64)
    gamma = torch.rand(C, generator=g, dtype=torch.float64) + 0.5
    beta = torch.randn(C, generator=g, dtype=torch.float64) * 0.2
    return x, res, dy, gamma, beta


def _worker(rank, world, port, out):
    import torch.distributed as dist
    from torch import nn
    from sparse2dense_b200 import autograd as AG
    from sparse2dense_b200.dense import ACT_RELU
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x, res, dy, gamma, beta = _data()
        sl = slice(0, SPLIT) if rank == 0 else slice(SPLIT, N)
        bn = nn.SyncBatchNorm(C, eps=1e-3, momentum=0.01).cuda().train()
        with torch.no_grad():
            bn.weight.copy_(gamma.float())
            bn.bias.copy_(beta.float())
        xl = x[sl].float().cuda().requires_grad_(True)
        rl = res[sl].float().cuda().requires_grad_(True)
        y = AG.norm_act(xl, bn, None, ACT_RELU, rl)
        y.backward(dy[sl].float().cuda())
        out.put((rank, y.detach().cpu().numpy(), xl.grad.cpu().numpy(), rl.grad.cpu().numpy(), bn.weight.grad.cpu().numpy(),
                 bn.bias.grad.cpu().numpy(), bn.running_mean.cpu().numpy(), bn.running_var.cpu().numpy()))
    finally:
        dist.destroy_process_group()


def test_sync_batchnorm_two_ranks_match_full_batch():
    import torch.multiprocessing as mp
    import torch.nn.functional as F
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((out.get(timeout=300) for _ in procs), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0

    x, r, dy, gamma, beta = _data()
    x.requires_grad_(True); r.requires_grad_(True); gamma.requires_grad_(True); beta.requires_grad_(True)
    rm, rv = torch.zeros(C, dtype=torch.float64), torch.ones(C, dtype=torch.float64)
    y = F.relu(F.batch_norm(x, rm, rv, gamma, beta, True, 0.01, 1e-3) + r)
    y.backward(dy)

    def rel(a, b):
        b = b.detach().numpy()
        return float(np.abs(a - b).max() / np.abs(b).max())
    ycat = np.concatenate([res[0][1], res[1][1]])
    assert rel(ycat, y) < 2e-5
    assert rel(np.concatenate([res[0][2], res[1][2]]), x.grad) < 1e-4
    assert rel(np.concatenate([res[0][3], res[1][3]]), r.grad) < 1e-4
    assert rel(res[0][4] + res[1][4], gamma.grad) < 1e-4           # affine gradients: local sums, added by the DDP reduce
    assert rel(res[0][5] + res[1][5], beta.grad) < 1e-4
    for k in (0, 1):
        assert rel(res[k][6], rm) < 1e-5 and rel(res[k][7], rv) < 1e-5      # both ranks hold the global running statistics


def test_converted_model_runs_like_the_original_in_one_process():
    """nn.SyncBatchNorm.convert_sync_batchnorm(student) (tools/train.py:92-96): with no process group the converted modules
    behave as plain BatchNorm -- eval detections unchanged (folded into the conv epilogues), training step finite."""
    from sparse2dense_b200 import ops, synth
    from sparse2dense_b200.trainer import DistillTrainer
    teacher, student = synth.build_distill_models("cuda", ops.PRECISION_AUTO)
    ex = synth.distill_example(1, small=True)
    student.eval()
    with torch.no_grad():
        d0 = student(ex, return_loss=False)
    conv = torch.nn.SyncBatchNorm.convert_sync_batchnorm(student)
    assert any(isinstance(m, torch.nn.SyncBatchNorm) for m in conv.modules())
    conv.eval()
    with torch.no_grad():
        d1 = conv(ex, return_loss=False)
    assert torch.equal(d0[0]["box3d_lidar"], d1[0]["box3d_lidar"]) and torch.equal(d0[0]["scores"], d1[0]["scores"])
    conv.neck.train_pcr = False
    log = DistillTrainer(teacher, conv, total_steps=10).step(ex)
    assert all(np.isfinite(float(v)) for v in log.values())
