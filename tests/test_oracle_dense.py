"""CPU: the dense-stage oracle (torch functional restatement) against the committed samples of the reference's
own S2D_RPN / CenterHead modules (tests/golden/neck_head_s2d.npz, generated through the import shim)."""
import logging
import os

import numpy as np
import torch

from oracle import neck_head as NH
from sparse2dense_b200 import registry, synth

from conftest import GOLDEN

NECK_CFG = dict(layer_nums=[5, 5], ds_layer_strides=[1, 2], ds_num_filters=[128, 256], us_layer_strides=[1, 2],
                us_num_filters=[256, 256], num_input_features=256)
HEAD_CFG = dict(in_channels=512, tasks=[dict(num_class=3, class_names=["VEHICLE", "PEDESTRIAN", "CYCLIST"])],
                dataset="waymo", weight=2, code_weights=[1.0] * 8,
                common_heads={"reg": (2, 2), "height": (1, 2), "dim": (3, 2), "rot": (2, 2)})


def bev_input(seed, batch=1, occupancy=0.25):
    rng = np.random.default_rng(seed)
    occ = rng.uniform(size=(batch, 1, 188, 188)) < occupancy
    return (np.abs(rng.normal(0, 1.0, size=(batch, 256, 188, 188))) * occ).astype(np.float32)


def build_modules():
    neck = registry.build_neck(dict(type="S2D_RPN", logger=logging.getLogger("t"), **NECK_CFG))
    head = registry.build_head(dict(type="CenterHead", **HEAD_CFG))
    return neck, head


def test_dense_oracle_matches_reference_samples():
    g = np.load(os.path.join(GOLDEN, "neck_head_s2d.npz"))
    neck, head = build_modules()
    ns = synth.random_module_state(neck, int(g["neck_seed"]))
    hs = synth.random_module_state(head, int(g["head_seed"]))
    x = torch.from_numpy(bev_input(int(g["input_seed"])))
    with torch.no_grad():
        ox, fa, fb = NH.s2d_rpn_forward(ns, x)
        oh = NH.center_head_forward(hs, ox)[0]
    outs = dict(x=ox, F_S_a=fa, F_S_b=fb, **oh)
    for name, t in outs.items():
        got = t.numpy().reshape(-1)[g[name + "_idx"]]
        assert np.abs(got - g[name + "_val"]).max() <= 1e-5 * float(g[name + "_absmax"]), name


def test_dense_modules_have_reference_state_dict_layout():
    neck, head = build_modules()
    nk, hk = neck.state_dict(), head.state_dict()
    assert sum(p.numel() for p in neck.parameters()) == 15270669          # S2D_RPN 15.27 M
    assert sum(p.numel() for p in head.parameters()) == 486731            # CenterHead 0.49 M
    for k, shape in {"encoder_1.0.weight": (256, 256, 2, 2), "convnext_block_2.1.weight": (256, 47, 47),
                     "decoder_1.0.weight": (256, 256, 4, 4), "blocks.0.1.weight": (128, 256, 3, 3),
                     "blocks.1.1.weight": (256, 128, 3, 3), "deblocks.0.0.weight": (256, 128, 1, 1),
                     "deblocks.1.0.weight": (256, 256, 2, 2), "generator_1.3.weight": (32, 32, 4, 4, 4),
                     "out_conv.0.weight": (640, 256, 1, 1)}.items():
        assert tuple(nk[k].shape) == shape, k
    for k, shape in {"shared_conv.0.weight": (64, 512, 3, 3), "tasks.0.hm.3.bias": (3,),
                     "tasks.0.reg.0.weight": (64, 64, 3, 3), "tasks.0.dim.3.weight": (3, 64, 3, 3)}.items():
        assert tuple(hk[k].shape) == shape, k
    assert list(head.tasks[0].heads) == ["reg", "height", "dim", "rot", "hm"]
    assert float(hk["tasks.0.hm.3.bias"][0]) == np.float32(-2.19)        # center_head.py:96-97
