"""CPU: the numpy restatement of AssignLabel (oracle/assign_label.py) against the outputs of the reference's own
AssignLabel code (tests/golden/assign_label.npz, produced by ``make_golden.py assign`` from the reference source)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import assign_label as OA
from sparse2dense_b200 import synth

SEEDS = (50, 51, 52, 53)


@pytest.mark.parametrize("seed", SEEDS)
def test_assign_label_oracle_matches_reference_golden(seed):
    g = np.load(os.path.join(GOLDEN, "assign_label.npz"))
    out = OA.assign_label(g[f"{seed}_boxes"], g[f"{seed}_classes"], [3], (1504, 1504), synth.WAYMO_RANGE, synth.WAYMO_VOXEL, 8,
                          0.1, 500, 2)
    for k in ("hm", "ind", "mask", "cat"):
        assert np.array_equal(out[k][0], g[f"{seed}_{k}"]), k            # heat map bit-exact
    assert np.allclose(out["anno_box"][0], g[f"{seed}_anno_box"], rtol=0, atol=1e-6)
    assert np.array_equal(out["gt_boxes_and_cls"], g[f"{seed}_gt_boxes_and_cls"])
