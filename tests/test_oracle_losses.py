"""CPU: loss oracle (oracle/ref_ops.py) against values produced by the reference's own FastFocalLoss / RegLoss classes and
the trainer's distillation expressions (tests/golden/make_golden.py loss)."""
import os

import numpy as np

from loss_common import loss_inputs
from oracle import ref_ops as R

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "losses.npz")


def oracle_losses(d):
    out = np.clip(1 / (1 + np.exp(-d["hm_logits"])), 1e-4, 1 - 1e-4).astype(np.float32)
    t = (1 / (1 + np.exp(-d["t_logits"]))).astype(np.float32)
    return dict(hm=R.fast_focal_loss(out, d["gt_hm"], d["ind"], d["mask"], d["cat"]),
                kd=R.fast_focal_loss(out, t, d["ind"], d["mask"], d["cat"]),
                reg=R.reg_loss(d["box"], d["mask"], d["ind"], d["anno"]),
                dl=R.reg_loss(d["box"], d["mask"], d["ind"], d["t_box"], squared=True),
                s2d=R.sparse2dense_loss(d["box"], d["t_box"], d["hm_logits"], d["t_logits"]))


def test_loss_oracle_reproduces_reference_values():
    g = np.load(G)
    for seed in (40, 41):
        got = oracle_losses(loss_inputs(seed))
        for k, v in got.items():
            np.testing.assert_allclose(np.asarray(v), g[f"{seed}_{k}"], rtol=2e-5, atol=1e-7)


def test_focal_loss_without_positives_is_minus_neg_sum():
    d = loss_inputs(41)
    d["mask"][:] = 0
    out = np.clip(1 / (1 + np.exp(-d["hm_logits"])), 1e-4, 1 - 1e-4).astype(np.float32)
    v = R.fast_focal_loss(out, d["gt_hm"], d["ind"], d["mask"], d["cat"])
    ref = -(np.log(1 - out) * out ** 2 * (1 - d["gt_hm"]) ** 4).astype(np.float64).sum()
    assert abs(v - ref) < 1e-4 * abs(ref)
