"""ctypes binding of ``libs2d_b200.so`` (the C ABI in ``include/s2d_b200.h``).

The library is the product: there is no CPU or PyTorch fallback.  A missing or unloadable
library raises at first use, and every non-zero status becomes a ``RuntimeError`` carrying
``s2d_last_error()``.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libs2d_b200.so")

_c_int_p = ctypes.POINTER(ctypes.c_int)
_c_float_p = ctypes.POINTER(ctypes.c_float)
_vp = ctypes.c_void_p
_i = ctypes.c_int
_sz = ctypes.c_size_t

# name -> (restype, argtypes); must list every symbol include/s2d_b200.h declares
SIGNATURES = {
    "s2d_version": (_i, []),
    "s2d_last_error": (ctypes.c_char_p, []),
    "s2d_kernel_launches": (ctypes.c_ulonglong, []),
    "s2d_export_i32": (_i, [_vp, _i, _vp, _vp]),
    "s2d_voxelize_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "s2d_voxelize": (_i, [_vp, _c_int_p, _i, _i, _i, _c_float_p, _c_float_p, _i, _i, _vp, _vp, _vp, _vp, _i, _vp,
                          _vp, _sz, _vp]),
    "s2d_voxel_mean": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp]),
    "s2d_grid_index_bytes": (_sz, [_i, _c_int_p, _i]),
    "s2d_grid_index_build": (_i, [_vp, _i, _vp, _i, _c_int_p, _vp, _sz, _vp]),
    "s2d_rulebook_subm": (_i, [_vp, _i, _i, _c_int_p, _c_int_p, _c_int_p, _vp, _vp, _i, _vp, _vp]),
    "s2d_conv_out_shape": (_i, [_c_int_p, _c_int_p, _c_int_p, _c_int_p, _c_int_p, _c_int_p]),
    "s2d_sparse_out_coords": (_i, [_vp, _i, _vp, _i, _c_int_p, _c_int_p, _c_int_p, _c_int_p, _c_int_p, _vp, _sz,
                                   _vp, _i, _vp, _vp]),
    "s2d_rulebook_sparse": (_i, [_vp, _i, _i, _c_int_p, _c_int_p, _c_int_p, _c_int_p, _c_int_p, _vp, _vp, _i, _vp,
                                 _vp]),
    "s2d_spconv_fwd": (_i, [_vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _i, _vp]),
    "s2d_spconv_tf32_supported": (_i, [_i, _i]),
    "s2d_spconv_packed_bytes": (_sz, [_i, _i, _i]),
    "s2d_spconv_pack_weights": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "s2d_conv_fwd": (_i, [_vp, _vp]),
    "s2d_rows_split": (_i, [_vp, ctypes.c_longlong, _i, _i, _vp, _i, _vp]),
    "s2d_grid2d_tile_rows_count": (_i, [_i, _i, _i]),
    "s2d_grid2d_tile_rows": (_i, [_i, _i, _i, _vp, _vp]),
    "s2d_conv_fwd_grid": (_i, [_vp, _i, _i, _i, _i, _i, _vp]),
    "s2d_table_tile_masks": (_i, [_vp, _i, _i, _i, _vp, _vp]),
    "s2d_grid2d_table": (_i, [_i, _i, _i, _i, _i, _i, _i, _vp, _i, _vp]),
    "s2d_grid2d_tconv_table": (_i, [_i, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp]),
    "s2d_nchw_to_nhwc": (_i, [_vp, _i, _i, _i, _vp, _i, _vp]),
    "s2d_nhwc_to_nchw": (_i, [_vp, _i, _i, _i, _i, _vp, _vp]),
    "s2d_dwconv2d": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "s2d_layernorm_workspace_bytes": (_sz, [_i]),
    "s2d_layernorm_chw": (_i, [_vp, _vp, _vp, _i, _i, _i, ctypes.c_float, _vp, _vp, _sz, _vp]),
    "s2d_dense_bev_nhwc": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _i, _vp]),
    "s2d_dense_bev": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "s2d_dense_bev_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "s2d_dense_bev_tiled": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "s2d_centerhead_decode": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "s2d_centerhead_select_workspace_bytes": (_sz, [_i, _i, _i]),
    "s2d_centerhead_select": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, ctypes.c_float, _i, _vp, _vp, _vp, _vp, _vp, _vp,
                                   _sz, _vp]),
    "s2d_nms_workspace_bytes": (_sz, [_i]),
    "s2d_nms_sorted": (_i, [_vp, _i, ctypes.c_float, _vp, _vp, _vp, _sz, _vp]),
    "s2d_iou_bev": (_i, [_vp, _i, _vp, _i, _vp, _vp]),
    "s2d_bev_box_features": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _vp, _i, _i, _c_float_p, _c_float_p, ctypes.c_float,
                                  _vp, _vp]),
    "s2d_roi_refine": (_i, [_vp, _vp, _vp, _i, _i, _vp, _i, _vp, _i, _vp, _vp, _vp]),
    "s2d_pfn_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                         _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "s2d_maxpool2d_rows": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _i, _vp]),
    "s2d_gather_rows": (_i, [_vp, _i, _vp, ctypes.c_longlong, _i, _vp, _i, _vp]),
    "s2d_grid2d_tconv_table_s": (_i, [_i, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _vp, _vp]),
    "s2d_loss_workspace_bytes": (_sz, []),
    "s2d_masked_mse": (_i, [_vp, _vp, ctypes.c_longlong, _vp, _vp, _sz, _vp]),
    "s2d_focal_loss": (_i, [_vp, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, _i,
                            _vp, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, _i,
                            _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _sz, _vp]),
    "s2d_gather_reg_loss": (_i, [_vp, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, _vp, _vp,
                                 ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, _i, _i, _i, _i, _vp, _vp, _vp,
                                 _vp, _sz, _vp]),
    "s2d_masked_mse_bwd": (_i, [_vp, _vp, ctypes.c_longlong, _vp, ctypes.c_float, ctypes.c_float, _vp, _vp, _vp]),
    "s2d_focal_loss_bwd": (_i, [_vp, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, _i, _vp, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp,
                                _vp, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, _vp]),
    "s2d_gather_reg_loss_bwd": (_i, [_vp, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, _vp, _vp, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, _i, _i, _i, _i, _vp, _vp, _vp, _vp,
                                     _vp, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, ctypes.c_longlong, _vp]),
    "s2d_pcr_loss": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, ctypes.c_longlong, _c_float_p, _vp, _vp, _sz, _vp]),
    "s2d_pcr_loss_bwd": (_i, [_vp, _vp, _i, _i, _i, _i, _vp, _vp, ctypes.c_longlong, _c_float_p, _vp, _vp, _vp, _vp, _vp]),
    "s2d_assign_label": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, ctypes.c_float, ctypes.c_float, ctypes.c_float, ctypes.c_float, _i, ctypes.c_double, _i, _i, _vp, _vp, _vp, _vp,
                              _vp, _vp, _vp]),
    "s2d_rulebook_subm_grouped_workspace_bytes": (_sz, [_i]),
    "s2d_rulebook_subm_grouped": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _sz, _vp]),
    "s2d_rulebook_sparse_grouped": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _sz, _vp]),
    "s2d_table_group_rows_workspace_bytes": (_sz, [_i]),
    "s2d_table_group_rows": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _sz, _vp]),
    "s2d_table_transpose": (_i, [_vp, _i, _i, _i, _vp, _vp, _i, _i, _vp]),
    "s2d_conv_wgrad_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "s2d_conv_wgrad": (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _vp, _i, _vp, _sz, _vp]),
    "s2d_conv_wgrad_bf2_supported": (_i, [_i, _i]),
    "s2d_conv_wgrad_bf2_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "s2d_conv_wgrad_bf2": (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _vp, _i, _vp, _sz, _vp]),
    "s2d_rows_workspace_bytes": (_sz, [_i]),
    "s2d_bn_train_stats": (_i, [_vp, _i, _i, _i, ctypes.c_float, ctypes.c_float, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                                _vp, _sz, _vp]),
    "s2d_rows_affine_act": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _i, _i, _i, _vp, _i, _vp]),
    "s2d_rows_affine_act_bwd": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _i, _i, _i, _vp, _i, _vp, _i, _vp, _vp, _sz, _vp]),
    "s2d_bn_train_bwd": (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _i, _vp, _vp, _vp, _sz, _vp]),
    "s2d_rows_workspace_sums_offset": (_sz, [_i]),
    "s2d_bn_train_sums": (_i, [_vp, _i, _i, _i, _vp, _sz, _vp]),
    "s2d_bn_train_finalize": (_i, [_vp, _vp, _i, ctypes.c_float, ctypes.c_float, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "s2d_bn_train_bwd_params": (_i, [_vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "s2d_bn_train_bwd_dx": (_i, [_vp, _i, _i, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _sz, _vp]),
    "s2d_layernorm_bwd_workspace_bytes": (_sz, [_i]),
    "s2d_layernorm_chw_bwd": (_i, [_vp, _vp, _i, _i, _i, ctypes.c_float, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "s2d_dwconv2d_wgrad_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "s2d_dwconv2d_wgrad": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _sz, _vp]),
    "s2d_grad_norm_workspace_bytes": (_sz, []),
    "s2d_grad_norm_clip": (_i, [_vp, ctypes.c_longlong, ctypes.c_float, _vp, _vp, _sz, _vp]),
    "s2d_adam_step": (_i, [_vp, _vp, _vp, _vp, ctypes.c_longlong, ctypes.c_float, ctypes.c_float, ctypes.c_float,
                           ctypes.c_float, ctypes.c_float, _i, _vp, _vp]),
}



class ConvParams(ctypes.Structure):
    """struct s2d_conv_params (include/s2d_b200.h)."""
    _fields_ = [("in_", _vp), ("weights", _vp), ("tbl", _vp), ("scale", _vp), ("shift", _vp), ("residual", _vp),
                ("out", _vp), ("out_rows", _vp), ("in_ld", _i), ("out_ld", _i), ("res_ld", _i), ("tbl_stride", _i),
                ("K", _i), ("n_in", _i), ("n_out", _i), ("Cin", _i), ("Cout", _i), ("act", _i),
                ("res_after_act", _i), ("precision", _i), ("in_split", _vp), ("out_split", _vp), ("tile_masks", _vp),
                ("in_split_ld", _i), ("out_split_ld", _i)]


class DecodeParams(ctypes.Structure):
    """struct s2d_decode_params (include/s2d_b200.h)."""
    _fields_ = [("reg", _vp), ("height", _vp), ("dim", _vp), ("rot", _vp), ("hm", _vp),
                ("ld_reg", _i), ("ld_height", _i), ("ld_dim", _i), ("ld_rot", _i), ("ld_hm", _i),
                ("B", _i), ("H", _i), ("W", _i), ("num_cls", _i),
                ("out_size_factor", ctypes.c_float), ("voxel_x", ctypes.c_float), ("voxel_y", ctypes.c_float),
                ("pc_x", ctypes.c_float), ("pc_y", ctypes.c_float), ("score_threshold", ctypes.c_float),
                ("range", ctypes.c_float * 6)]


_lib = None


class S2DError(RuntimeError):
    pass


def load():
    """Load (once) and return the ctypes handle; raises if the CUDA library was not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise S2DError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C sparse2dense_b200/csrc`).  There is no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)          # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status, what=""):
    if status != 0:
        msg = load().s2d_last_error().decode("utf-8", "replace")
        raise S2DError(f"{what or 's2d call'} failed with status {status}: {msg}")


def ints(values):
    values = [int(v) for v in values]
    return (ctypes.c_int * len(values))(*values)


def floats(values):
    values = [float(v) for v in values]
    return (ctypes.c_float * len(values))(*values)
