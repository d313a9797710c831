#!/usr/bin/env python
"""CenterHead in training mode: forward values and gradients, fp32 kernels vs a tensor-core precision."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparse2dense_b200 import ops, synth  # noqa: E402


def run(head, x, dys, precision):
    head.set_precision(precision)
    head.zero_grad(set_to_none=True)
    xx = x.clone().requires_grad_(True)
    out = head.forward_rows(xx, 2, 188, 188)[0]
    loss = sum((out[h] * dys[h]).sum() for h in out)
    loss.backward()
    g = {k: p.grad.detach().clone() for k, p in head.named_parameters() if p.grad is not None}
    g["input"] = xx.grad.detach().clone()
    return {h: v.detach().clone() for h, v in out.items()}, g


def main():
    _, student = synth.build_distill_models("cuda", ops.PRECISION_AUTO)
    head = student.bbox_head.train()
    torch.manual_seed(0)
    x = torch.randn(2 * 188 * 188, 512, device="cuda")
    dys = None
    ref = None
    for name in ["fp32"] + (sys.argv[1:] or ["auto"]):
        if dys is None:
            with torch.no_grad():
                tmp = head.forward_rows(x, 2, 188, 188)[0]
            # sparse upstream gradient (a few hundred cells), like the regression losses
            dys = {h: torch.zeros_like(v) for h, v in tmp.items()}
            idx = torch.randint(0, x.shape[0], (120,), device="cuda")
            for h in dys:
                dys[h][idx] = torch.randn(120, dys[h].shape[1], device="cuda")
        out, g = run(head, x, dys, ops.PRECISION_NAMES[name])
        if ref is None:
            ref = (out, g)
            continue
        print("==", name)
        for h in out:
            print(f"  out {h:8s} {float((out[h] - ref[0][h]).abs().max() / ref[0][h].abs().max()):.2e}")
        for k in g:
            s = float(ref[1][k].abs().max())
            if s > 0:
                print(f"  grad {k:40s} {float((g[k] - ref[1][k]).abs().max()) / s:.2e}")


if __name__ == "__main__":
    main()
