"""ctypes binding of libtdr_sm100.so (the C ABI declared in include/tdr_sm100.h).

There is deliberately NO fallback: if the shared library is missing or a call fails, a
``TdrError`` is raised.  The product path never imports ``oracle``.
"""
import ctypes as C
import os

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TDR_LIB_PATH") or os.path.join(_PKG, "libtdr_sm100.so")


class TdrError(RuntimeError):
    pass


class ConvGemmDesc(C.Structure):
    """Mirror of ``tdr_conv_gemm_desc`` (include/tdr_sm100.h) -- field order and types must match."""
    _fields_ = [
        ("in_", C.c_void_p), ("in_ld", C.c_longlong),
        ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("Ci", C.c_int),
        ("weight", C.c_void_p), ("w_ld", C.c_longlong),
        ("Co", C.c_int), ("KH", C.c_int), ("KW", C.c_int), ("stride", C.c_int), ("pad", C.c_int), ("dil", C.c_int),
        ("w_batched", C.c_int),
        ("w_batch_stride", C.c_longlong),
        ("origin", C.c_void_p),
        ("n_images", C.c_int), ("img_h", C.c_int), ("img_w", C.c_int),
        ("bias", C.c_void_p), ("rowscale", C.c_void_p),
        ("alpha", C.c_float),
        ("scale_ptr", C.c_void_p),
        ("act", C.c_int),
        ("res1", C.c_void_p), ("res1_ld", C.c_longlong), ("res1_scale", C.c_float),
        ("res2", C.c_void_p), ("res2_ld", C.c_longlong), ("res2_bf16", C.c_int),
        ("out_f32", C.c_void_p), ("out_f32_ld", C.c_longlong),
        ("out_bf16", C.c_void_p), ("out_bf16_ld", C.c_longlong),
        ("store_mode", C.c_int), ("impl", C.c_int),
        ("ln_mode", C.c_int), ("ln_eps", C.c_float), ("ln_weight", C.c_void_p), ("ln_bias", C.c_void_p),
        ("ln_out_bf16", C.c_void_p), ("ln_out_ld", C.c_longlong),
        ("in_fp16", C.c_int), ("out_fp16", C.c_int),
    ]


class GdfnTailDesc(C.Structure):
    """Mirror of ``tdr_gdfn_tail_desc`` (include/tdr_sm100.h)."""
    _fields_ = [
        ("hidden", C.c_void_p), ("hidden_ld", C.c_longlong),
        ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("hp", C.c_int), ("C", C.c_int),
        ("dw_weight", C.c_void_p), ("dw_bias", C.c_void_p),
        ("w_out", C.c_void_p), ("w_ld", C.c_longlong),
        ("bias", C.c_void_p), ("alpha", C.c_float), ("scale_ptr", C.c_void_p),
        ("res1", C.c_void_p), ("res1_ld", C.c_longlong), ("res1_scale", C.c_float),
        ("res2", C.c_void_p), ("res2_ld", C.c_longlong),
        ("out", C.c_void_p), ("out_ld", C.c_longlong),
        ("fp16", C.c_int),
    ]


class PatchDesc(C.Structure):
    """Mirror of ``tdr_patch_desc`` (include/tdr_sm100.h)."""
    _fields_ = [("image", C.c_void_p), ("h", C.c_int), ("w", C.c_int), ("top", C.c_int), ("left", C.c_int),
                ("mode", C.c_int), ("reserved", C.c_int)]


_vp, _ll, _i, _f, _sz = C.c_void_p, C.c_longlong, C.c_int, C.c_float, C.c_size_t

# name -> (restype, argtypes); every symbol include/tdr_sm100.h declares
SIGNATURES = {
    "tdr_last_error": (C.c_char_p, []),
    "tdr_version": (_i, []),
    "tdr_check_device": (_i, []),
    "tdr_set_pdl": (_i, [_i]),
    "tdr_conv_gemm": (_i, [C.POINTER(ConvGemmDesc), _vp]),
    "tdr_conv_gemm_ln_supported": (_i, [C.POINTER(ConvGemmDesc)]),
    "tdr_conv_gemm_desc_layout": (None, [C.POINTER(_i)]),
    "tdr_gdfn_tail": (_i, [C.POINTER(GdfnTailDesc), _vp]),
    "tdr_gdfn_tail_supported": (_i, [C.POINTER(GdfnTailDesc)]),
    "tdr_conv3x3_small_ci": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp, _ll, _vp, _ll, _vp]),
    "tdr_conv3x3_small_co": (_i, [_vp, _ll, _i, _i, _i, _i, _vp, _vp, _i, _vp, _vp, _vp]),
    "tdr_rownorm": (_i, [_vp, _ll, _ll, _i, _i, _vp, _vp, _f, _i, _vp, _ll, _vp, _ll, _vp]),
    "tdr_dwconv3x3": (_i, [_vp, _ll, _i, _i, _i, _i, _vp, _vp, _i, _vp, _ll, _vp]),
    "tdr_gate_mul": (_i, [_vp, _ll, _ll, _i, _vp, _ll, _i, _vp]),
    "tdr_naf_sca_workspace_bytes": (_sz, [_i, _ll, _i]),
    "tdr_naf_sca_fold": (_i, [_vp, _ll, _i, _ll, _i, _vp, _vp, _vp, _i, _vp, _vp, _ll, _vp, _vp, _vp, _vp, _ll, _i, _vp]),
    "tdr_mdta_partials_bytes": (_sz, [_i, _ll, _i, _i]),
    "tdr_mdta_gram": (_i, [_vp, _ll, _i, _ll, _i, _i, _vp, _i, _vp]),
    "tdr_mdta_weff": (_i, [_vp, _i, _ll, _i, _i, _vp, _vp, _vp, _ll, _vp, _vp, _vp, _i, _vp, _vp]),
    "tdr_grouped_stencil": (_i, [_vp, _ll, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _ll, _i, _vp]),
    "tdr_vit_patchify": (_i, [_vp, _i, _i, _i, _i, _i, _vp, _ll, _vp]),
    "tdr_vit_assemble_tokens": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "tdr_softmax_rows": (_i, [_vp, _ll, _ll, _i, _f, _vp, _ll, _vp]),
    "tdr_vit_transpose_v": (_i, [_vp, _ll, _i, _i, _i, _i, _i, _vp, _i, _vp]),
    "tdr_vit_attention_supported": (_i, [_i]),
    "tdr_vit_attention": (_i, [_vp, _ll, _i, _i, _i, _i, _f, _vp, _ll, _vp]),
    "tdr_mefc_gate": (_i, [_vp, _ll, _i, _i, _vp, _vp, _i, _vp, _vp, _i, _i, _vp, _vp]),
    "tdr_mefc_mix_weights": (_i, [_vp, _i, _i, _i, _vp, _ll, _i, _vp, _ll, _i, _vp]),
    "tdr_prompt_weights": (_i, [_vp, _ll, _vp, _vp, _i, _i, _i, _vp, _vp]),
    "tdr_prompt_mix_resize": (_i, [_vp, _i, _i, _i, _vp, _i, _i, _i, _vp, _ll, _i, _vp]),
    "tdr_crop_resize": (_i, [_vp, _i, _i, _i, _vp, _i, _i, _i, _i, _i, _vp, _vp]),
    "tdr_cosine_rows": (_i, [_vp, _vp, _i, _i, _ll, _vp, _vp]),
    "tdr_mean_tokens": (_i, [_vp, _ll, _i, _i, _i, _i, _i, _vp, _ll, _i, _vp]),
    "tdr_sumsq_partial_count": (_i, []),
    "tdr_sumsq_partial": (_i, [_vp, _ll, _vp, _vp]),
    "tdr_clip_coef": (_i, [_vp, _i, _f, _f, _vp, _vp]),
    "tdr_adamw_step": (_i, [_vp, _vp, _vp, _vp, _ll, _f, _f, _f, _f, _f, _i, _f, _vp, _vp]),
    "tdr_ema_update": (_i, [_vp, _vp, _ll, _f, _vp]),
    "tdr_l1_loss_grad": (_i, [_vp, _vp, _ll, _f, _vp, _vp, _vp, _vp]),
    "tdr_image_to_rows16": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _i, _vp]),
    "tdr_add_gaussian_noise": (_i, [_vp, _vp, _vp, _i, _ll, C.c_ulonglong, _vp, _vp]),
    "tdr_prepare_patches": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp]),
    "tdr_psnr_u8_sums": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "tdr_nchw_to_nhwc": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _ll, _vp, _ll, _vp]),
    "tdr_nhwc_to_nchw": (_i, [_vp, _ll, _i, _i, _i, _i, _i, _i, _vp, _ll, _vp, _vp]),
    "tdr_copy_rows_f32": (_i, [_vp, _ll, _ll, _i, _vp, _ll, _vp, _ll, _i, _vp]),
    "tdr_sqnorm_rows": (_i, [_vp, _ll, _ll, _i, _vp, _vp]),
    "tdr_masa_split3": (_i, [_vp, _ll, _ll, _i, _vp, _ll, _vp]),
    "tdr_cast_rows": (_i, [_vp, _ll, _ll, _i, _vp, _ll, _vp, _ll, _vp, _vp, _i, _vp]),
    "tdr_masa_level_scale": (_i, [_vp, _ll, _ll, _i, _vp, _vp, _vp]),
    "tdr_scale_vec": (_i, [_vp, _ll, _vp, _vp, _vp]),
    "tdr_cvt_f16_bf16": (_i, [_vp, _ll, _ll, _i, _vp, _ll, _vp]),

    "tdr_masa_ref_invnorm": (_i, [_vp, _i, _i, _i, C.POINTER(_i), _i, _vp, _vp]),
    "tdr_masa_coarse_filters": (_i, [_vp, _i, _i, _i, _i, _i, _i, C.POINTER(_i), _i, _i, _vp, _vp]),
    "tdr_masa_coarse_argmax": (_i, [_vp, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "tdr_masa_fine_filters": (_i, [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp]),
    "tdr_masa_win_invnorm": (_i, [_vp, _i, _i, _vp, _i, _i, _i, _vp, _vp]),
    "tdr_masa_fine_argmax": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp]),
    "tdr_masa_transfer": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _ll, _vp, _ll, _vp]),
}



class WgradDesc(C.Structure):
    """Mirror of ``tdr_wgrad_desc`` (include/tdr_sm100.h)."""
    _fields_ = [
        ("dy", C.c_void_p), ("dy_ld", C.c_longlong),
        ("x", C.c_void_p), ("x_ld", C.c_longlong),
        ("B", C.c_int), ("H", C.c_int), ("W", C.c_int), ("Ci", C.c_int), ("Co", C.c_int), ("KH", C.c_int),
        ("KW", C.c_int), ("stride", C.c_int), ("pad", C.c_int), ("dil", C.c_int),
        ("per_sample", C.c_int),
        ("out", C.c_void_p),
        ("out_stride_b", C.c_longlong), ("out_stride_co", C.c_longlong), ("out_stride_ci", C.c_longlong),
        ("out_stride_tap", C.c_longlong),
        ("co_map", C.c_void_p), ("ci_map", C.c_void_p),
        ("accumulate", C.c_int), ("scale", C.c_float),
        ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("scale_ptr", C.c_void_p),
    ]


SIGNATURES.update({
    "tdr_wgrad_workspace_bytes": (_sz, [C.POINTER(WgradDesc)]),
    "tdr_wgrad": (_i, [C.POINTER(WgradDesc), _vp]),
    "tdr_reduce_workspace_bytes": (_sz, [_i]),
    "tdr_colsum": (_i, [_vp, _ll, _ll, _i, _vp, _ll, _vp, _i, _vp, _vp]),
    "tdr_dwconv3x3_wgrad": (_i, [_vp, _ll, _vp, _ll, _i, _i, _i, _i, _vp, _vp, _vp, _i, _vp, _vp]),
    "tdr_rownorm_bwd": (_i, [_vp, _ll, _vp, _ll, _ll, _i, _i, _vp, _f, _vp, _ll, _vp, _ll, _vp, _ll, _vp, _vp, _i, _vp, _vp]),
    "tdr_gate_bwd": (_i, [_vp, _ll, _vp, _ll, _ll, _i, _i, _vp, _ll, _vp, _ll, _vp]),
    "tdr_dwconv3x3_gated_train": (_i, [_vp, _ll, _i, _i, _i, _i, _vp, _vp, _i, _vp, _ll, _vp, _ll, _vp]),
    "tdr_dwconv3x3_gate_bwd": (_i, [_vp, _ll, _i, _i, _i, _i, _vp, _vp, _i, _vp, _ll, _vp, _vp, _ll, _vp]),
    "tdr_naf_scaled_conv_bwd": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tdr_naf_sca_bwd": (_i, [_vp, _i, _i, _i, _vp, _vp, _vp, _vp, _ll, _vp, _vp, _vp, _vp, _vp]),
    "tdr_mdta_bwd_workspace_bytes": (_sz, [_i, _i, _i]),
    "tdr_mdta_bwd": (_i, [_vp, _vp, _i, _ll, _i, _i, _vp, _vp, _vp, _vp, _ll, _vp, _vp, _i, _vp, _vp]),
    "tdr_scale_add_f32": (_i, [_vp, _ll, _vp, _ll, _ll, _i, _vp, _f, _vp, _ll, _vp]),
    "tdr_dot_f32": (_i, [_vp, _ll, _vp, _ll, _ll, _i, _vp, _i, _vp, _vp]),
    "tdr_pixel_shuffle_nhwc": (_i, [_vp, _ll, _i, _i, _i, _i, _i, _vp, _ll, _vp]),
    "tdr_pack_conv_weight": (_i, [_vp, _i, _i, _i, _i, _vp, _i, _vp, _i, _vp, _vp, _ll, _vp, _ll, _i, _vp]),
    "tdr_pack_dw_weight": (_i, [_vp, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp]),
    "tdr_gather_vec": (_i, [_vp, _vp, _i, _i, _vp, _vp]),
    "tdr_relu_mask": (_i, [_vp, _ll, _vp, _ll, _ll, _i, _vp, _ll, _vp]),
    "tdr_masa_transfer_bwd": (_i, [_vp, _ll, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "tdr_masa_fine_bwd": (_i, [_vp, _i, _i, _i, _vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "tdr_dilate2_nhwc": (_i, [_vp, _ll, _i, _i, _i, _i, _vp, _ll, _i, _i, _vp]),
})

_lib = None


def load():
    """Load the shared library (once).  Raises TdrError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TdrError(
            f"{LIB_PATH} not found: build it with `python -m textualdegremoval_b200.csrc.build` "
            "(or __graft_entry__.build()).  There is no CPU / PyTorch fallback for the B200 hot path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().tdr_last_error()
        raise TdrError(f"{what or 'libtdr_sm100'} failed (code {rc}): {msg.decode() if msg else ''}")


def call(name: str, *args):
    lib = load()
    check(getattr(lib, name)(*args), name)
