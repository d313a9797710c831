"""Shared helpers of the pillar tests (inputs, module states, fixture comparison)."""
import logging
import os

import numpy as np
import torch

from oracle import pillars as OP
from oracle import ref_ops as R
from sparse2dense_b200 import registry, synth

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pillars_s2d.npz")
PP_VOXEL, PP_RANGE = (0.32, 0.32, 6.0), (-74.88, -74.88, -2, 74.88, 74.88, 4.0)
PP_RPN = dict(layer_nums=[3, 5, 5], ds_layer_strides=[1, 2, 2], ds_num_filters=[64, 128, 256], us_layer_strides=[1, 2, 4],
              us_num_filters=[128, 128, 128], num_input_features=64)


def pillar_inputs(seed):
    cloud = synth.lidar_scene(seed)
    v, c, n = R.points_to_voxel(cloud, PP_VOXEL, PP_RANGE, 20, True, 32000)
    return v, np.concatenate([np.zeros((len(c), 1), np.int32), c], 1), n


def module_states(d):
    reader = registry.build_reader(dict(type="PillarFeatureNet", num_filters=[64, 64], num_input_features=5,
                                        with_distance=False, voxel_size=PP_VOXEL, pc_range=PP_RANGE))
    bb = registry.build_backbone(dict(type="PointPillarsScatter_S2D", ds_factor=1))
    neck = registry.build_neck(dict(type="RPN", logger=logging.getLogger("x"), **PP_RPN))
    return (reader, synth.random_module_state(reader, int(d["reader_seed"]))), \
           (bb, synth.random_module_state(bb, int(d["backbone_seed"]))), \
           (neck, synth.random_module_state(neck, int(d["neck_seed"])))


def check(d, name, got, tol):
    got = np.asarray(got).reshape(-1)[d[name + "_idx"]]
    err = np.abs(got - d[name + "_val"]).max() / float(d[name + "_absmax"])
    assert err < tol, (name, err)
    return err


