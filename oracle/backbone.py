"""CPU restatement of SpMiddleResNetFHD.forward (det3d/models/backbones/scn.py:88-185).
TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.

``state`` is a reference-format state dict (numpy arrays): ``conv_input.0.weight`` [3,3,3,5,16],
``conv_input.1.{weight,bias,running_mean,running_var}``, ``conv1.0.conv1.{weight,bias}``,
``conv1.0.bn1.*`` ...  BatchNorm runs in eval mode (running statistics), eps from norm_cfg
(scn.py:100-101: 1e-3).
"""
import numpy as np

from . import ref_ops as R

BN_EPS = 1e-3


def _bn(state, prefix, conv_bias=None):
    g, b = state[prefix + ".weight"], state[prefix + ".bias"]
    m, v = state[prefix + ".running_mean"], state[prefix + ".running_var"]
    scale = (g / np.sqrt(v.astype(np.float32) + np.float32(BN_EPS))).astype(np.float32)
    shift = (b - m * scale).astype(np.float32)
    if conv_bias is not None:
        shift = (shift + conv_bias * scale).astype(np.float32)
    return scale, shift


def _basic_block(state, prefix, feats, tbl, wide):
    """SparseBasicBlock.forward (scn.py:69-85)."""
    s1, h1 = _bn(state, prefix + ".bn1", state.get(prefix + ".conv1.bias"))
    out = R.bn_act(R.spconv_fwd(feats, state[prefix + ".conv1.weight"], tbl, wide), s1, h1, None, True)
    s2, h2 = _bn(state, prefix + ".bn2", state.get(prefix + ".conv2.bias"))
    return R.bn_act(R.spconv_fwd(out, state[prefix + ".conv2.weight"], tbl, wide), s2, h2, feats, True)


def backbone_forward(state, voxel_features, coors, batch_size, input_shape, wide=False, return_stats=False):
    """-> (bev [B,256,H,W], {'conv1'..'conv4': (features, indices, spatial_shape)}, stats)."""
    shape = tuple((np.array(input_shape[::-1]) + [1, 0, 0]).tolist())          # scn.py:159
    coors = np.ascontiguousarray(coors, np.int32)
    stats = {"N": [len(coors)], "pairs": {}, "shapes": [shape]}

    tbl0, p0 = R.rulebook_subm(coors, shape, 3)                                # indice_key res0
    stats["pairs"]["res0"] = p0
    s, h = _bn(state, "conv_input.1")
    x = R.bn_act(R.spconv_fwd(voxel_features, state["conv_input.0.weight"], tbl0, wide), s, h, None, True)
    x = _basic_block(state, "conv1.0", x, tbl0, wide)
    x = _basic_block(state, "conv1.1", x, tbl0, wide)
    multi = {"conv1": (x, coors, shape)}

    stages = [("conv2", 3, 2, 1, "res1"), ("conv3", 3, 2, 1, "res2"), ("conv4", 3, 2, (0, 1, 1), "res3")]
    for name, ks, st, pd, key in stages:
        oc, tbl_d, shape_o, pd_pairs = R.rulebook_sparse(coors, shape, ks, st, pd)
        stats["pairs"][name + ".0"] = pd_pairs
        s, h = _bn(state, name + ".1")
        x = R.bn_act(R.spconv_fwd(x, state[name + ".0.weight"], tbl_d, wide), s, h, None, True)
        coors, shape = oc, tuple(shape_o.tolist())
        tbl, pr = R.rulebook_subm(coors, shape, 3)
        stats["pairs"][key] = pr
        x = _basic_block(state, name + ".3", x, tbl, wide)
        x = _basic_block(state, name + ".4", x, tbl, wide)
        multi[name] = (x, coors, shape)
        stats["N"].append(len(coors))
        stats["shapes"].append(shape)

    oc, tbl_e, shape_e, pe = R.rulebook_sparse(coors, shape, (3, 1, 1), (2, 1, 1), 0)
    stats["pairs"]["extra_conv.0"] = pe
    s, h = _bn(state, "extra_conv.1")
    x = R.bn_act(R.spconv_fwd(x, state["extra_conv.0.weight"], tbl_e, wide), s, h, None, True)
    stats["N"].append(len(oc))
    stats["shapes"].append(tuple(shape_e.tolist()))
    bev = R.dense_bev(x, oc, batch_size, shape_e)                               # scn.py:173-176
    multi["extra"] = (x, oc, tuple(shape_e.tolist()))
    return (bev, multi, stats) if return_stats else (bev, multi)


FHD_LAYERS = [("s", 0), ("s", 0), ("d", (3, 2, 1)), ("s", 1), ("s", 1), ("d", (3, 2, 1)), ("s", 2), ("s", 2), ("s", 2),
              ("d", (3, 2, (0, 1, 1))), ("s", 3), ("s", 3), ("s", 3)]


def fhd_forward(state, voxel_features, coors, batch_size, input_shape, wide=False):
    """SpMiddleFHD.forward (det3d/models/backbones/scn.py:187-289): conv / BN(eval) / ReLU groups ``middle_conv.{3i}``,
    ``middle_conv.{3i+1}`` and ``extra_conv.{0,1}`` -> (bev [B,128,H,W], (conv_4 features, indices, spatial_shape))."""
    shape = tuple((np.array(input_shape[::-1]) + [1, 0, 0]).tolist())
    coors = np.ascontiguousarray(coors, np.int32)
    x, tables = voxel_features, {}
    for i, (kind, arg) in enumerate(FHD_LAYERS):
        w = state[f"middle_conv.{3 * i}.weight"]
        s, h = _bn(state, f"middle_conv.{3 * i + 1}")
        if kind == "s":
            key = (arg, len(coors))
            if key not in tables:
                tables[key] = R.rulebook_subm(coors, shape, 3)[0]
            tbl = tables[key]
        else:
            ks, st, pd = arg
            coors, tbl, shape_o, _ = R.rulebook_sparse(coors, shape, ks, st, pd)
            shape = tuple(shape_o.tolist())
        x = R.bn_act(R.spconv_fwd(x, w, tbl, wide), s, h, None, True)
    conv_4 = (x, coors, shape)
    oc, tbl_e, shape_e, _ = R.rulebook_sparse(coors, shape, (3, 1, 1), (2, 1, 1), 0)
    s, h = _bn(state, "extra_conv.1")
    x = R.bn_act(R.spconv_fwd(x, state["extra_conv.0.weight"], tbl_e, wide), s, h, None, True)
    return R.dense_bev(x, oc, batch_size, shape_e), conv_4


def random_state(seed=0, num_input_features=5):
    """Seeded random reference-format weights (shared with bench.py so both arms use the same numbers)."""
    from sparse2dense_b200.synth import backbone_state
    return backbone_state(seed, num_input_features)
