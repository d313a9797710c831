"""CPU: the PointPillars + S2D oracle (oracle/pillars.py) against the fixture generated from the reference's own
PillarFeatureNet / PointPillarsScatter_S2D / RPN modules (tests/golden/make_golden.py pp)."""
import numpy as np
import torch

from oracle import pillars as OP
from pillar_common import G, PP_RANGE, PP_RPN, PP_VOXEL, check, module_states, pillar_inputs


def test_pillar_oracle_reproduces_reference_fixture():
    d = np.load(G)
    v, c, n = pillar_inputs(int(d["scene_seed"]))
    assert len(v) == int(d["n_pillars"]) and int(c.astype(np.int64).sum()) == int(d["coors_checksum"])
    (_, rs), (_, bs), (_, ns) = module_states(d)
    with torch.no_grad():
        f = OP.pfn_forward(rs, v, n, c, PP_VOXEL, PP_RANGE)
        fa, fb = OP.scatter_s2d_forward(bs, f, c, 1, 468, 468)
        x = OP.rpn_forward(ns, fa, PP_RPN["layer_nums"], PP_RPN["ds_layer_strides"], PP_RPN["us_layer_strides"])
    for name, got in (("pfn", f), ("F_S_a", fa), ("F_S_b", fb), ("x", x)):
        check(d, name, got.numpy(), 1e-5)
