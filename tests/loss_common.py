"""Shared inputs of the loss tests (the same generator produced tests/golden/losses.npz)."""
import numpy as np


def loss_inputs(seed, B=2, C=3, H=36, W=40, M=60):
    rng = np.random.default_rng(seed)
    f = np.float32
    d = dict(hm_logits=rng.normal(-2.5, 1.5, (B, C, H, W)).astype(f), t_logits=rng.normal(-2.5, 1.5, (B, C, H, W)).astype(f),
             gt_hm=np.clip(rng.normal(0.05, 0.2, (B, C, H, W)), 0, 1).astype(f),
             box=rng.normal(0, 1, (B, 8, H, W)).astype(f), t_box=rng.normal(0, 1, (B, 8, H, W)).astype(f),
             anno=rng.normal(0, 1, (B, M, 8)).astype(f), ind=rng.integers(0, H * W, (B, M)).astype(np.int64),
             mask=(rng.uniform(size=(B, M)) < 0.6).astype(np.uint8), cat=rng.integers(0, C, (B, M)).astype(np.int64))
    d["mask"][1] = 0 if seed % 2 else d["mask"][1]
    return d
