#!/usr/bin/env python
"""Forward error of one gather-GEMM layer against float64, per precision mode (max error / max |out| and rms error / rms out)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sparse2dense_b200 import dense, ops  # noqa: E402


def main():
    torch.manual_seed(0)
    B, H, W = 2, 94, 94
    for cin, cout in ((64, 64), (256, 256), (512, 64)):
        x = torch.randn(B * H * W, cin, dtype=torch.float64).relu()            # post-ReLU activations (non-zero mean)
        w = torch.randn(9, cin, cout, dtype=torch.float64) / np.sqrt(9 * cin)
        tbl, Ho, Wo = dense.conv_table(torch.device("cuda"), B, H, W, 3, 1, 1)
        t = tbl.cpu().long()
        xp = torch.cat([x, x.new_zeros(1, cin)])
        idx = torch.where(t < 0, torch.full_like(t, x.shape[0]), t)
        ref = sum(xp[idx[k]] @ w[k] for k in range(9))
        xc, wc = x.float().cuda(), w.float().cuda()
        line = [f"{cin:4d}->{cout:4d}"]
        for name in ("fp32", "tf32x3", "tf32_bf16c", "tf32"):
            y = dense.conv_rows(xc, wc, tbl, B * Ho * Wo, precision=ops.PRECISION_NAMES[name]).double().cpu()
            e = y - ref
            line.append(f"{name}: max {float(e.abs().max() / ref.abs().max()):.1e} rms {float(e.pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()):.1e} "
                        f"bias {float((e * ref.sign()).mean() / ref.abs().mean()):+.1e}")
        yt = torch.zeros_like(ref)
        print("  ".join(line))


if __name__ == "__main__":
    main()
