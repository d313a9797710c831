"""GPU: s2d_assign_label through sparse2dense_b200.pipeline.AssignLabel against the golden outputs of the reference's own
AssignLabel (tests/golden/assign_label.npz) and the numpy oracle.  Bar: ind / mask / cat / gt_boxes_and_cls identical; heat
map identical up to the rounding of one float64 exp (<= 1 ulp of float32 on a vanishing fraction of cells); regression
targets within 1 ulp (float32 log / sin / cos of numpy vs correctly rounded here)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN
from sparse2dense_b200 import synth
from sparse2dense_b200.pipeline import AssignLabel

pytestmark = pytest.mark.gpu
SEEDS = (50, 51, 52, 53)
CFG = dict(out_size_factor=8, target_assigner=dict(tasks=[dict(num_class=3, class_names=["VEHICLE", "PEDESTRIAN", "CYCLIST"])]),
           gaussian_overlap=0.1, max_objs=500, min_radius=2)


def test_assign_label_batch_matches_reference_golden():
    g = np.load(os.path.join(GOLDEN, "assign_label.npz"))
    out = AssignLabel(cfg=CFG)([g[f"{s}_boxes"] for s in SEEDS], [g[f"{s}_classes"] for s in SEEDS], (1504, 1504, 40),
                               synth.WAYMO_RANGE, synth.WAYMO_VOXEL)
    for b, s in enumerate(SEEDS):
        for k in ("ind", "mask", "cat"):
            assert np.array_equal(out[k][0][b].cpu().numpy(), g[f"{s}_{k}"]), (s, k)
        bc = out["gt_boxes_and_cls"][b].cpu().numpy()
        assert np.array_equal(bc, g[f"{s}_gt_boxes_and_cls"]), s
        hm, ref = out["hm"][0][b].cpu().numpy(), g[f"{s}_hm"]
        assert np.array_equal(hm > 0, ref > 0)
        diff = hm != ref
        assert diff.mean() < 1e-4 and np.abs(hm - ref).max() <= 6e-8, (s, diff.sum())
        anno, ra = out["anno_box"][0][b].cpu().numpy(), g[f"{s}_anno_box"]
        assert np.abs(anno - ra).max() <= 1.2e-7 * max(1.0, np.abs(ra).max()), s
        assert np.array_equal(anno[:, [0, 1, 2, 6, 7]], ra[:, [0, 1, 2, 6, 7]])      # offsets, z, velocity: exact


def test_assign_label_feeds_the_training_step():
    """Targets made on the device drive CenterHead.loss (shapes / dtypes of the example dict)."""
    from sparse2dense_b200 import losses as L
    g = np.load(os.path.join(GOLDEN, "assign_label.npz"))
    ex = AssignLabel(cfg=CFG)([g["50_boxes"], g["51_boxes"]], [g["50_classes"], g["51_classes"]], (1504, 1504, 40),
                              synth.WAYMO_RANGE, synth.WAYMO_VOXEL)
    assert tuple(ex["hm"][0].shape) == (2, 3, 188, 188) and ex["ind"][0].dtype == torch.int64
    logits = torch.randn(2, 3, 188, 188, device="cuda")
    loss = L.fastfocalloss(logits, ex["hm"][0], ex["ind"][0], ex["mask"][0], ex["cat"][0], out_is_logits=True)
    assert np.isfinite(float(loss))
