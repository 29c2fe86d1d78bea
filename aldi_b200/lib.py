"""ctypes binding of the C-ABI library (include/aldi_b200.h).

The library is the product's compute path; there is NO CPU or PyTorch fallback: if the shared object is
missing or a call fails, an exception is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libaldi_b200.so")

F32, BF16 = 0, 1

c_void_p, c_int, c_ll, c_float, c_double, c_size_t = (
    ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_double, ctypes.c_size_t)


class ConvParams(ctypes.Structure):
    _fields_ = [
        ("x", c_void_p), ("x_c", c_int), ("x_w", c_int), ("x_h", c_int), ("x_n", c_int),
        ("x_sw", c_ll), ("x_sh", c_ll), ("x_sn", c_ll),
        ("w", c_void_p), ("cout_p", c_int),
        ("taps_h", c_int), ("taps_w", c_int), ("pad_h", c_int), ("pad_w", c_int), ("stride", c_int),
        ("n", c_int), ("ho", c_int), ("wo", c_int),
        ("scale", c_void_p), ("bias", c_void_p),
        ("residual", c_void_p), ("res_mode", c_int), ("res_sw", c_ll), ("res_sh", c_ll), ("res_sn", c_ll),
        ("mask", c_void_p), ("mask_sw", c_ll), ("mask_sh", c_ll), ("mask_sn", c_ll),
        ("out", c_void_p), ("out_dtype", c_int), ("cout_store", c_int),
        ("out_sw", c_ll), ("out_sh", c_ll), ("out_sn", c_ll),
        ("relu", c_int), ("accumulate", c_int),
    ]


class WgradParams(ctypes.Structure):
    _fields_ = [
        ("x", c_void_p), ("x_c", c_int), ("x_w", c_int), ("x_h", c_int), ("x_n", c_int),
        ("x_sw", c_ll), ("x_sh", c_ll), ("x_sn", c_ll),
        ("dy", c_void_p), ("dy_c", c_int), ("dy_sw", c_ll), ("dy_sh", c_ll), ("dy_sn", c_ll),
        ("n", c_int), ("ho", c_int), ("wo", c_int),
        ("taps_h", c_int), ("taps_w", c_int), ("pad_h", c_int), ("pad_w", c_int), ("stride", c_int),
        ("scale", c_void_p), ("dw", c_void_p), ("cout_store", c_int), ("cin_store", c_int),
        ("dbias", c_void_p),
    ]


class RefreshDesc(ctypes.Structure):
    _fields_ = [
        ("kind", c_int), ("out_dtype", c_int), ("w", c_void_p), ("bn_w", c_void_p), ("bn_b", c_void_p),
        ("bn_mean", c_void_p), ("bn_var", c_void_p), ("out", c_void_p), ("out2", c_void_p),
        ("cout", c_int), ("taps", c_int), ("cin", c_int), ("cout_p", c_int), ("cin_p", c_int), ("eps", c_float),
    ]


class RoiAlignParams(ctypes.Structure):
    _fields_ = [
        ("feat", c_void_p * 4), ("dfeat", c_void_p * 4), ("feat_h", c_int * 4), ("feat_w", c_int * 4),
        ("scale", c_float * 4), ("num_levels", c_int), ("min_level", c_int), ("canonical_box_size", c_float),
        ("canonical_level", c_int), ("channels", c_int), ("pooled", c_int), ("dtype", c_int),
        ("rois", c_void_p), ("roi_batch", c_void_p), ("num_valid", c_void_p), ("num_rois", c_int),
        ("out", c_void_p), ("dout", c_void_p),
    ]


class MsdaParams(ctypes.Structure):
    _fields_ = [
        ("value", c_void_p), ("sampling_loc", c_void_p), ("attn_weight", c_void_p),
        ("spatial_h", ctypes.POINTER(c_int)), ("spatial_w", ctypes.POINTER(c_int)), ("level_start", ctypes.POINTER(c_int)),
        ("n", c_int), ("s", c_int), ("m", c_int), ("d", c_int), ("lq", c_int), ("l", c_int), ("p", c_int),
        ("dtype", c_int),
        ("out", c_void_p), ("grad_out", c_void_p), ("grad_value", c_void_p), ("grad_loc", c_void_p),
        ("grad_attn", c_void_p),
    ]


AUG_MAX_RADIUS = 8


class AugParams(ctypes.Structure):
    _fields_ = [
        ("h", c_int), ("w", c_int),
        ("src_plane", c_ll), ("src_row", c_ll), ("dst_plane", c_ll), ("dst_row", c_ll),
        ("do_color", c_int), ("contrast_w", c_double), ("brightness_w", c_double), ("saturation_w", c_double),
        ("do_gray", c_int), ("blur_radius", c_int), ("blur_taps", c_double * (2 * AUG_MAX_RADIUS + 1)),
        ("num_erase", c_int), ("erase_rect", (c_int * 4) * 3), ("erase_seed", ctypes.c_uint * 3),
        ("mic_mask", c_void_p), ("mic_h", c_int), ("mic_w", c_int),
    ]


class RpnSparseParams(ctypes.Structure):
    _fields_ = [
        ("levels", c_void_p), ("dtype", c_int), ("channels", c_int),
        ("feat", c_void_p * 5), ("feat_sn", c_ll * 5), ("feat_sh", c_ll * 5), ("feat_sw", c_ll * 5),
        ("hidden", c_void_p * 5),
        ("dfeat", c_void_p * 5), ("dfeat_sn", c_ll * 5), ("dfeat_sh", c_ll * 5), ("dfeat_sw", c_ll * 5),
        ("drpn", c_void_p), ("dstride", c_int), ("idx", c_void_p), ("count", c_void_p), ("cap", c_int),
        ("dy_g", c_void_p), ("t_g", c_void_p), ("x_g", c_void_p), ("err_flag", c_void_p),
    ]


class AttnParams(ctypes.Structure):
    _fields_ = [
        ("qkv", c_void_p), ("batch", c_int), ("gh", c_int), ("gw", c_int), ("heads", c_int),
        ("row_stride", c_ll), ("batch_stride", c_ll),
        ("rel_h", c_void_p), ("rel_w", c_void_p), ("scale", c_float), ("dtype", c_int),
        ("out", c_void_p), ("out_stride", c_ll), ("out_batch_stride", c_ll), ("lse", c_void_p),
        ("dout", c_void_p), ("dqkv", c_void_p), ("drel_h", c_void_p), ("drel_w", c_void_p), ("delta", c_void_p), ("impl", c_int),
    ]


class RpnLevels(ctypes.Structure):
    _fields_ = [
        ("num_levels", c_int), ("num_anchors", c_int),
        ("h", c_int * 5), ("w", c_int * 5), ("stride", c_int * 5), ("loc_off", c_int * 5),
        ("total_locs", c_int), ("ch_stride", c_int),
        ("cell", (c_float * 4) * 3 * 5),
        ("scale_clamp", c_float), ("min_box_size", c_float),
    ]


c_uint = ctypes.c_uint
P = c_void_p  # device / host pointers

# name -> (restype, argtypes); every symbol declared in include/aldi_b200.h must appear here
SIGNATURES = {
    "aldi_last_error": (ctypes.c_char_p, []),
    "aldi_abi_version": (c_int, []),
    "aldi_launch_count": (ctypes.c_ulonglong, []),
    "aldi_reset_launch_count": (None, []),
    "aldi_set_pdl": (None, [c_int]),
    "aldi_ema_update": (c_int, [c_void_p, c_void_p, c_size_t, c_double, c_void_p]),
    "aldi_sgd_momentum_step": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_float, c_float, c_float, c_float,
                                       c_void_p, c_double, c_void_p]),
    "aldi_pack_weight": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int,
                                 c_void_p]),
    "aldi_refresh_blocks": (c_int, [ctypes.POINTER(RefreshDesc)]),
    "aldi_refresh_operands": (c_int, [c_void_p, c_void_p, c_int, c_int, c_void_p]),
    "aldi_conv_tc": (c_int, [ctypes.POINTER(ConvParams), c_void_p]),
    "aldi_conv_f32": (c_int, [ctypes.POINTER(ConvParams), c_void_p]),
    "aldi_wgrad_tc": (c_int, [ctypes.POINTER(WgradParams), c_void_p]),
    "aldi_wgrad_f32": (c_int, [ctypes.POINTER(WgradParams), c_void_p]),
    "aldi_preprocess": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, P, P, P]),
    "aldi_stem_im2col": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, P, P, P]),
    "aldi_stem_s2d": (c_int, [P, P, P, c_int, c_int, c_int, P, P, P]),
    "aldi_stem_s2d_f32": (c_int, [P, P, P, c_int, c_int, c_int, P, P, P]),
    "aldi_split_bf16": (c_int, [P, c_int, c_int, c_int, c_int, c_ll, c_ll, c_ll, P, c_ll, c_int, P]),
    "aldi_conv_epilogue_f32": (c_int, [P, ctypes.POINTER(ConvParams), P]),
    "aldi_maxpool3x3s2": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "aldi_sum2x2_accum": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "aldi_add_f32": (c_int, [P, c_int, P, c_size_t, P]),
    "aldi_cast_f32": (c_int, [P, c_int, P, c_size_t, P]),
    "aldi_gap_backward": (c_int, [P, P, c_int, c_int, c_ll, c_int, c_int, c_float, P, P, P]),
    "aldi_domain_head_loss": (c_int, [P, c_int, c_int, c_int, c_ll, P, c_int, P, P, c_float, c_float, c_float, c_int, P, P,
                                      c_ll, P, P, P, P]),
    "aldi_colsum": (c_int, [P, c_int, c_int, c_ll, c_ll, c_ll, c_int, c_float, P, P]),
    "aldi_frozenbn_fold": (c_int, [P, P, P, P, c_float, P, P, c_int, P]),
    "aldi_roi_align_forward": (c_int, [ctypes.POINTER(RoiAlignParams), P]),
    "aldi_roi_align_backward": (c_int, [ctypes.POINTER(RoiAlignParams), P]),
    "aldi_rpn_topk_workspace_bytes": (c_size_t, [c_int, c_int]),
    "aldi_rpn_topk_decode": (c_int, [P, ctypes.POINTER(RpnLevels), c_int, c_int, P, P, P, P, P, P, c_int, P, P, c_size_t,
                                     P]),
    "aldi_nms_workspace_bytes": (c_size_t, [c_int, c_int]),
    "aldi_nms_sorted": (c_int, [P, P, P, P, P, c_int, c_int, c_float, c_int, P, c_size_t, P, P, P, P, P, P]),
    "aldi_nms_segmented_workspace_bytes": (c_size_t, [c_int, c_int, ctypes.POINTER(c_int), c_int]),
    "aldi_nms_segmented": (c_int, [P, P, P, c_int, c_int, c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int), c_float, c_int,
                                   P, c_size_t, P, P, P, P, P, P]),
    "aldi_rpn_label_workspace_bytes": (c_size_t, [c_int, c_int]),
    "aldi_rpn_label_anchors": (c_int, [ctypes.POINTER(RpnLevels), c_int, P, P, c_int, c_float, c_float, c_int, c_float,
                                       P, P, P, c_size_t, P, P, P, P]),
    "aldi_roi_label_sample": (c_int, [P, P, c_int, c_int, P, P, P, c_int, c_float, c_int, c_int, c_float, P, P,
                                      c_int, P, P, P, P, P, P, P, P]),
    "aldi_roi_inference_candidates": (c_int, [P, c_int, P, P, c_int, c_int, c_int, P, c_float, P, c_float, P, P, P, P,
                                              P, c_int, P]),
    "aldi_pseudo_label_threshold": (c_int, [P, P, P, P, c_int, c_int, c_float, P, P, P, P, c_int, P]),
    "aldi_rpn_sparse_compact": (c_int, [P, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P]),
    "aldi_rpn_sparse_gather": (c_int, [ctypes.POINTER(RpnSparseParams), P]),
    "aldi_rpn_sparse_scatter": (c_int, [ctypes.POINTER(RpnSparseParams), P, P]),
    "aldi_rpn_loss": (c_int, [P, ctypes.POINTER(RpnLevels), c_int, P, P, P, P, c_int, c_int, c_float, c_float, c_float,
                              P, c_int, c_int, c_int, P, P]),
    "aldi_roi_loss": (c_int, [P, c_int, c_int, c_int, P, P, P, P, c_int, P, c_float, c_float, c_float, P, c_int, c_int,
                              P, P]),
    "aldi_distill_rpn_loss": (c_int, [P, P, ctypes.POINTER(RpnLevels), c_int, P, P, c_float, c_float, c_float, c_float,
                                      P, c_int, c_int, c_int, P, P]),
    "aldi_distill_roi_loss": (c_int, [P, P, c_int, c_int, c_int, P, P, c_int, c_float, c_int, c_float, c_float,
                                      c_float, P, c_int, c_int, c_int, P, P]),
    "aldi_domain_bce_loss": (c_int, [P, c_int, c_int, c_float, c_float, c_float, P, c_int, c_int, P, P]),
    "aldi_strong_augment_workspace_bytes": (c_size_t, [c_int, c_int]),
    "aldi_strong_augment": (c_int, [P, P, ctypes.POINTER(AugParams), P, c_size_t, P]),
    "aldi_layernorm_forward": (c_int, [P, P, P, c_float, c_ll, c_int, c_int, c_int, P, P, P]),
    "aldi_layernorm_backward": (c_int, [P, P, P, P, c_ll, c_int, c_int, c_int, P, c_int, P, P, P]),
    "aldi_dwconv7": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, c_int, P]),
    "aldi_dwconv7_wgrad": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, P, P]),
    "aldi_gelu": (c_int, [P, P, P, c_size_t, c_int, P]),
    "aldi_layerscale_forward": (c_int, [P, P, P, P, c_ll, c_ll, c_int, c_int, c_int, P, P]),
    "aldi_layerscale_backward": (c_int, [P, P, P, P, c_ll, c_ll, c_int, c_int, c_int, P, P, P]),
    "aldi_space_to_depth": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "aldi_patchify_image": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P]),
    "aldi_adamw_step": (c_int, [P, P, P, P, c_size_t, c_float, c_float, c_float, c_float, c_float, c_int, c_float, P]),
    "aldi_msda_forward": (c_int, [ctypes.POINTER(MsdaParams), P]),
    "aldi_msda_backward": (c_int, [ctypes.POINTER(MsdaParams), P]),
    "aldi_window_partition": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P]),
    "aldi_add_rows_bcast": (c_int, [P, P, c_int, c_ll, c_int, c_int, P]),
    "aldi_sum_over_batch": (c_int, [P, c_int, c_ll, c_int, c_int, P, P]),
    "aldi_bicubic_resize": (c_int, [P, c_int, c_int, P, c_int, c_int, c_int, c_int, c_int, P]),
    "aldi_linear_resize_rows": (c_int, [P, c_int, P, c_int, c_int, c_int, P]),
    "aldi_maxpool2x2": (c_int, [P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "aldi_maxpool2x2_backward": (c_int, [P, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "aldi_relu": (c_int, [P, P, P, c_size_t, c_int, P]),
    "aldi_relpos_transpose": (c_int, [P, c_int, P, P, c_int, c_int, c_int, c_int, c_int, P]),
    "aldi_attention_forward": (c_int, [ctypes.POINTER(AttnParams), P]),
    "aldi_attention_backward": (c_int, [ctypes.POINTER(AttnParams), P]),
}


class AldiError(RuntimeError):
    pass


_lib = None


def load():
    """Load libaldi_b200.so; raise loudly if it has not been built (python -m aldi_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AldiError(
            "libaldi_b200.so not found at %s: build it with `python -m aldi_b200.build` "
            "(there is no CPU/PyTorch fallback for the hot path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != 0:
        msg = load().aldi_last_error().decode("utf-8", "replace")
        raise AldiError("%s failed (rc=%d): %s" % (what or "aldi call", rc, msg))


def launch_count():
    return int(load().aldi_launch_count())


def reset_launch_count():
    load().aldi_reset_launch_count()
