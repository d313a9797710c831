"""The distillation training step: ``TS_Trainer.batch_processor_inline`` (det3d/torchie/trainer/trainer.py:726-847, the
CenterPoint branch :775-811) + ``OptimizerHook.after_train_iter`` (hooks/optimizer.py:15-21) + the fastai-style
``OptimWrapper`` over Adam (det3d/solver/fastai_optim.py:118-174, ``build_one_cycle_optimizer`` apis/train.py:168-186)
+ the ``OneCycle`` schedule (det3d/solver/learning_schedules_fastai.py:7-95), B200-first:

* teacher (eval, no grad) and student run on NHWC rows; every forward / backward kernel is libs2d_b200.so (autograd.py);
* all trainable parameters live in ONE flat fp32 buffer and their gradients in another, so that the data-parallel
  exchange is a single NCCL all-reduce over NVLink (74 MB for the 18.4 M-parameter student), gradient clipping is one
  norm reduction and the optimizer is one fused launch (``s2d_grad_norm_clip`` + ``s2d_adam_step``);
* one process per GPU, scenes sharded over ranks, BatchNorm statistics are per rank exactly as the reference's default
  (``torch.nn.parallel.DistributedDataParallel`` without SyncBN conversion unless ``--sync_bn``).
"""
import math

import numpy as np
import torch
import torch.distributed as dist

from . import _lib, losses as L


def annealing_cos(start, end, pct):
    """learning_schedules_fastai.py:68-72."""
    return end + (start - end) / 2 * (np.cos(np.pi * pct) + 1)


class OneCycle:
    """learning_schedules_fastai.py:7-95 (LRSchedulerStep + OneCycle): lr low -> lr_max over ``pct_start`` of the run,
    then lr_max -> low / 1e4; momentum moms[0] -> moms[1] -> moms[0]; both cosine."""

    def __init__(self, total_step, lr_max, moms, div_factor, pct_start):
        self.total_step, self.lr_max, self.moms = int(total_step), lr_max, tuple(moms)
        low = lr_max / div_factor
        a1 = int(self.total_step * pct_start)
        self.lr_phases = [(0, a1, lambda p: annealing_cos(low, lr_max, p)),
                          (a1, self.total_step, lambda p: annealing_cos(lr_max, low / 1e4, p))]
        self.mom_phases = [(0, a1, lambda p: annealing_cos(self.moms[0], self.moms[1], p)),
                           (a1, self.total_step, lambda p: annealing_cos(self.moms[1], self.moms[0], p))]
        self.lr, self.mom = low, self.moms[0]

    def step(self, step):
        for start, end, func in self.lr_phases:
            if step >= start and end > start:
                self.lr = float(func((step - start) / (end - start)))
        for start, end, func in self.mom_phases:
            if step >= start and end > start:
                self.mom = float(func((step - start) / (end - start)))
        return self.lr, self.mom


class FlatAdam:
    """``OptimWrapper(true_wd=True, bn_wd=True)`` over ``torch.optim.Adam(betas=(mom, 0.99))``: every step
    ``p *= 1 - wd*lr`` then Adam (fastai_optim.py:158-174), on one flat buffer with one kernel launch."""

    def __init__(self, params, wd=0.01, beta2=0.99, eps=1e-8):
        self.params = [p for p in params if p.requires_grad]
        assert self.params, "no trainable parameter"
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.flat_p = torch.empty((total,), dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros((total,), dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros_like(self.flat_p)
        self.exp_avg_sq = torch.zeros_like(self.flat_p)
        off = 0
        for p in self.params:                                  # parameters and gradients become views of the flat buffers
            n = p.numel()
            self.flat_p[off:off + n].copy_(p.data.reshape(-1))
            p.data = self.flat_p[off:off + n].view_as(p.data)
            p.grad = self.flat_g[off:off + n].view_as(p.data)
            off += n
        self.wd, self.beta2, self.eps = wd, beta2, eps
        self.steps = 0
        lib = _lib.load()
        self._ws = torch.empty((lib.s2d_grad_norm_workspace_bytes(),), dtype=torch.uint8, device=dev)
        self._norm = torch.zeros((2,), dtype=torch.float32, device=dev)        # ||g||, clip coefficient

    def zero_grad(self):
        self.flat_g.zero_()
        off = 0
        for p in self.params:                                  # autograd may have replaced a .grad: point it back
            n = p.numel()
            if p.grad is None or p.grad.data_ptr() != self.flat_g.data_ptr() + 4 * off:
                p.grad = self.flat_g[off:off + n].view_as(p.data)
            off += n

    def all_reduce(self):
        """DDP gradient averaging: one all-reduce of the flat gradient buffer (NCCL over NVLink / NVSwitch)."""
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat_g, op=dist.ReduceOp.SUM)
            self.flat_g.mul_(1.0 / dist.get_world_size())

    def step(self, lr, mom, max_norm=35.0):
        lib = _lib.load()
        st = torch.cuda.current_stream().cuda_stream
        n = self.flat_p.numel()
        self.steps += 1
        _lib.check(lib.s2d_grad_norm_clip(self.flat_g.data_ptr(), n, float(max_norm if max_norm else 0.0),
                                          self._norm.data_ptr(), self._ws.data_ptr(), self._ws.numel(), st),
                   "s2d_grad_norm_clip")
        _lib.check(lib.s2d_adam_step(self.flat_p.data_ptr(), self.flat_g.data_ptr(), self.exp_avg.data_ptr(),
                                     self.exp_avg_sq.data_ptr(), n, float(lr), float(mom), float(self.beta2),
                                     float(self.eps), float(self.wd), self.steps, self._norm[1:].data_ptr(), st),
                   "s2d_adam_step")
        # the kernel wrote the parameters behind autograd's back: bump their version counters so that derived-weight
        # caches (packed tensor-core images, folded BN) keyed on tensor versions are rebuilt by the next eval forward
        torch.autograd.graph.increment_version(self.params)
        return self._norm                                        # device [2]: gradient norm before clipping, coefficient


    # -- checkpoint / resume (the role of optimizer.state_dict() in det3d/torchie/trainer/checkpoint.py:199-240) -----------
    def state_dict(self):
        """Moments in the flat layout plus the per-parameter shapes they were taken with (so that a load into a model with a
        different parameter order fails loudly)."""
        return dict(exp_avg=self.exp_avg.detach().cpu(), exp_avg_sq=self.exp_avg_sq.detach().cpu(), steps=self.steps,
                    shapes=[tuple(p.shape) for p in self.params], wd=self.wd, beta2=self.beta2, eps=self.eps)

    def load_state_dict(self, state):
        if [tuple(x) for x in state["shapes"]] != [tuple(p.shape) for p in self.params]:
            raise ValueError("optimizer state was saved for a different parameter list")
        self.exp_avg.copy_(state["exp_avg"])
        self.exp_avg_sq.copy_(state["exp_avg_sq"])
        self.steps = int(state["steps"])
        self.wd, self.beta2, self.eps = state["wd"], state["beta2"], state["eps"]


def distill_losses(student, r, T_preds, F_D_a, F_D_b, example, s2d_weights=(10.0, 20.0, 5.0, 20.0)):
    """The CenterPoint branch of batch_processor_inline (trainer.py:775-811) on rows.  ``r`` = KD_VoxelNet.student_rows()."""
    B, H, W, Hu, Wu = r["dims"]
    losses = r["loss"]
    s2d = L.sparse2dense_loss(r["F_S_a"], F_D_a, r["F_S_b"], F_D_b, s2d_weights)
    S, T = r["preds"][0], T_preds[0]
    rows = lambda t: L.Rows(t, B, Hu * Wu)
    ind, mask, cat = example["ind"][0], example["mask"][0], example["cat"][0]
    # the reference's S_preds[0]['hm'] is the clamped sigmoid (CenterHead.loss rewrites the dict entry, center_head.py:254)
    kd_hm = L.fastfocalloss(rows(S["hm"]), rows(T["hm"]), ind, mask, cat, out_is_logits=True, target_is_logits=True)
    names = ["reg", "height", "dim"] + (["vel"] if "vel" in S else []) + ["rot"]
    kd_reg = torch.cat([L.distill_reg_loss(rows(S[n]), rows(T[n]), mask, ind) for n in names])
    head = student.bbox_head
    kd_reg = (kd_reg * kd_reg.new_tensor(head.code_weights)).sum() * head.weight
    distill = kd_hm + kd_reg + s2d
    pcr = r["mask_loss"] + r["comp_loss"]
    # trainer.py:799-803 adds the distillation terms to losses['loss'][0]; parse_second_losses then SUMS losses['loss'] over
    # all tasks, so every task's detection loss reaches backward (single-task Waymo: one entry)
    total = sum(losses["loss"][1:], losses["loss"][0]) + distill + pcr
    with torch.no_grad():
        t_hm = L.fastfocalloss(rows(T["hm"]), example["hm"][0], ind, mask, cat, out_is_logits=True)
    log = dict(loss=total.detach(), hm_loss=losses["hm_loss"][0], loc_loss=losses["loc_loss"][0].detach(),
               sparse2dense_loss=s2d.detach(), kd_hm_loss=kd_hm.detach(), kd_reg_loss=kd_reg.detach(),
               mask_loss=torch.as_tensor(r["mask_loss"]).detach(), reconstruction_loss=torch.as_tensor(r["comp_loss"]).detach(),
               T_hm_loss=t_hm)
    return total, log


class DistillTrainer:
    """One data-parallel training step of the student under a frozen teacher.

    ``step(example)``: teacher forward (no grad) -> student forward -> losses -> backward -> gradient all-reduce ->
    clip(35) -> weight decay + Adam with the one-cycle lr / momentum of this iteration.  Returns the loss log (device
    scalars; nothing is read back to the host)."""

    def __init__(self, teacher, student, total_steps, lr_max=0.003, moms=(0.95, 0.85), div_factor=10.0, pct_start=0.3,
                 wd=0.01, max_norm=35.0):
        self.teacher = teacher.eval()
        for p in self.teacher.parameters():
            p.requires_grad = False
        self.student = student.train()
        self.opt = FlatAdam(self.student.parameters(), wd=wd)
        self.sched = OneCycle(total_steps, lr_max, moms, div_factor, pct_start)
        self.max_norm = max_norm
        self.global_step = 0
        self.events = None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.broadcast(self.opt.flat_p, src=0)               # DDP: every rank starts from rank 0's weights
            for b in self.student.buffers():                     # ... and buffers (DDP broadcast_buffers: BN running stats)
                dist.broadcast(b, src=0)

    def forward_backward(self, example):
        with torch.no_grad():
            T_preds, F_D_a, F_D_b, _ = self.teacher.teacher_rows(example, recon=True)
        assert F_D_b is not None, "the distillation example needs the reconstruction_* voxel family"
        self.opt.zero_grad()
        r = self.student.student_rows(example)
        total, log = distill_losses(self.student, r, T_preds, F_D_a, F_D_b, example)
        total.backward()
        return log

    def _timed(self, name, fn):
        """Run ``fn`` between two CUDA events when ``self.events`` is a list (bench.py's per-phase breakdown)."""
        if self.events is None:
            return fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        self.events.append((name, e0, e1))
        return out

    def step(self, example):
        lr, mom = self.sched.step(self.global_step)
        log = self._timed("forward_backward", lambda: self.forward_backward(example))
        self._timed("all_reduce", self.opt.all_reduce)
        norm = self._timed("optimizer", lambda: self.opt.step(lr, mom, self.max_norm))
        self.global_step += 1
        log["grad_norm"] = norm[0]
        log["lr"], log["mom"] = lr, mom
        return log

    def state_dict(self):
        """What the reference's ``save_checkpoint`` stores (model ``state_dict``, ``optimizer``, ``meta`` with the
        iteration; trainer/checkpoint.py:199-240), for the student."""
        return dict(state_dict=self.student.state_dict(), optimizer=self.opt.state_dict(),
                    meta=dict(iter=self.global_step))

    def load_state_dict(self, ckpt):
        with torch.no_grad():                                # parameters stay views of the flat buffer: copy in place
            own = self.student.state_dict()
            for k, v in ckpt["state_dict"].items():
                own[k].copy_(v)
        torch.autograd.graph.increment_version(self.opt.params)
        self.opt.load_state_dict(ckpt["optimizer"])
        self.global_step = int(ckpt["meta"]["iter"])

