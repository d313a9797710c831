#!/usr/bin/env python
"""Per-parameter gradient difference of the student's training step between two precisions of the conv kernels."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparse2dense_b200 import ops, synth  # noqa: E402
from sparse2dense_b200.trainer import distill_losses  # noqa: E402


def grads(student, teacher, ex, precision):
    student.set_precision(precision)
    student.zero_grad(set_to_none=True)
    with torch.no_grad():
        T_preds, F_D_a, F_D_b, _ = teacher.teacher_rows(ex)
    r = student.student_rows(ex)
    total, log = distill_losses(student, r, T_preds, F_D_a, F_D_b, ex)
    total.backward()
    return {k: p.grad.detach().clone() for k, p in student.named_parameters() if p.grad is not None}, float(total)


def main():
    teacher, student = synth.build_distill_models("cuda", ops.PRECISION_AUTO)
    teacher.eval(); student.train()
    student.neck.train_pcr = False
    ex = synth.distill_example(2, small=True)
    a, la = grads(student, teacher, ex, ops.PRECISION_FP32)
    for name in sys.argv[1:] or ["auto"]:
        b, lb = grads(student, teacher, ex, ops.PRECISION_NAMES[name])
        print(f"== fp32 vs {name}: loss {la:.6f} {lb:.6f}")
        for k in a:
            s = float(a[k].abs().max())
            if s == 0:
                continue
            e = float((a[k] - b[k]).abs().max()) / s
            if e > 1e-3:
                print(f"  {k:50s} {e:.2e}")


if __name__ == "__main__":
    main()
