"""ctypes binding of ``csrc/libhyres_b200.so`` (the C-ABI declared in include/hyres_b200.h).

The library is the product: if it is missing or the device is not sm_100 the
package raises -- there is no CPU or PyTorch fallback behind these calls.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "csrc", "libhyres_b200.so")

HYRES_CONV, HYRES_DECONV_K5S2 = 0, 1
EPI_LINEAR, EPI_ADD, EPI_GATE, EPI_GDN, EPI_IGDN, EPI_PIXSCALE = range(6)
ACT_NONE, ACT_RELU, ACT_PRELU, ACT_CLAMP01 = range(4)
SPLIT_COPY, SPLIT_ADD, SPLIT_GATE, SPLIT_GDN, SPLIT_IGDN, SPLIT_SQUARE, SPLIT_ROUND_CHAN = range(7)
SPLIT_F16 = 16  # HYRES_SPLIT_F16: format flag of an nsplit code (2 | SPLIT_F16 = two half parts)


class HyresError(RuntimeError):
    pass


class ConvIO(C.Structure):
    _fields_ = [
        ("x0", C.c_void_p), ("x1", C.c_void_p),
        ("B", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("epi", C.c_int), ("act", C.c_int), ("slope", C.c_float),
        ("aux0", C.c_void_p), ("ld_aux0", C.c_int),
        ("aux1", C.c_void_p), ("ld_aux1", C.c_int),
        ("pixscale", C.c_void_p),
        ("out_bf16", C.c_void_p), ("ld_out", C.c_int),
        ("out_sq", C.c_void_p), ("ld_sq", C.c_int),
        ("out_f32", C.c_void_p),
        ("f32_sb", C.c_int64), ("f32_sh", C.c_int64), ("f32_sw", C.c_int64), ("f32_sc", C.c_int64),
        ("mt_hint", C.c_int), ("ld_x0", C.c_int), ("x0_square", C.c_int),
        ("out_pad", C.c_int), ("up_t2", C.c_void_p), ("up_t3", C.c_void_p),
        ("cta_limit", C.c_int),
        ("split_mode", C.c_int), ("aux0_f32", C.c_void_p), ("aux1_f32", C.c_void_p),
        ("out_split", C.c_void_p), ("out_nsplit", C.c_int), ("split_square", C.c_int),
    ]


class RuIO(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("ld_x", C.c_int),
        ("out", C.c_void_p), ("ld_out", C.c_int),
        ("B", C.c_int), ("H", C.c_int), ("W", C.c_int),
        ("final_relu", C.c_int),
    ]


class RansGroup(C.Structure):
    _fields_ = [
        ("symbols", C.c_void_p), ("index", C.c_void_p), ("enc", C.c_void_p), ("rows", C.c_void_p),
        ("scratch", C.c_void_p), ("n", C.c_int64), ("cap_words", C.c_int64), ("n_entries", C.c_int64),
        ("n_rows", C.c_int32), ("count", C.c_int32), ("slots", C.c_int32),
    ]


_lib = None

_vp, _i, _i64, _f, _u64 = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_uint64

# name -> (restype, argtypes); mirrors include/hyres_b200.h one to one.
SIGNATURES = {
    "hyres_version": (_i, []),
    "hyres_device_check": (_i, [_i]),
    "hyres_last_error": (C.c_char_p, []),
    "hyres_launch_count": (C.c_longlong, []),
    "hyres_conv_create": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp]),
    "hyres_conv_create_split": (_i, [C.POINTER(_vp), _i, _i, _i, _i, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp, _i]),
    "hyres_conv_update": (_i, [_vp, _vp, _vp]),
    "hyres_conv_update_device": (_i, [_vp, _vp, _vp, _vp]),
    "hyres_colsum_workspace_bytes": (_i64, [_i64, _i]),
    "hyres_colsum_bf16": (_i, [_vp, _i64, _i, _vp, _vp, _vp]),
    "hyres_wgrad_supported": (_i, [_vp]),
    "hyres_wgrad_workspace_bytes": (_i64, [_vp, _i, _i, _i]),
    "hyres_wgrad_run": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp, _vp, _vp]),
    "hyres_conv_packed_elems": (_i64, [_vp, _i]),
    "hyres_conv_export_packed": (_i, [_vp, _vp, _vp, _vp]),
    "hyres_conv_import_packed": (_i, [_vp, _vp, _vp, _vp]),
    "hyres_conv_destroy": (None, [_vp]),
    "hyres_conv_macs_per_pos": (_i64, [_vp]),
    "hyres_conv_out_size": (_i, [_vp, _i, _i, C.POINTER(_i), C.POINTER(_i)]),
    "hyres_conv_run": (_i, [_vp, C.POINTER(ConvIO), _vp]),
    "hyres_ru_supported": (_i, [_vp, _vp, _vp]),
    "hyres_ru_run": (_i, [_vp, _vp, _vp, C.POINTER(RuIO), _vp]),
    "hyres_conv3ch_run": (_i, [_vp, _i, _i, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _f, _vp]),
    "hyres_final_clamp": (_i, [_vp, _vp, _vp, _i64, _vp]),
    "hyres_gc_quant_pass": (_i, [_vp, _vp, _i, _i, _u64, _vp, _vp, _i, _i, _i, _i, _vp]),
    "hyres_gc_merge_likelihood": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _u64, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "hyres_gc_symbols": (_i, [_vp, _vp, _i, _vp, _i, _f, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp]),
    "hyres_gc_codes": (_i, [_vp, _i, _vp, _i, _f, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "hyres_gc_indexes": (_i, [_vp, _vp, _i, _f, _vp, _i, _i, _i, _i, _vp]),
    "hyres_gc_dequant": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "hyres_add_to_bf16": (_i, [_vp, _vp, _vp, _i64, _vp]),
    "hyres_split_f32": (_i, [_vp, _i64, _i, _i, _vp, _vp, _vp, _i, _vp, _vp, _i, _vp]),
    "hyres_residual_im2col5s2_split": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "hyres_symbols_to_nhwc_f32": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "hyres_eb_forward": (_i, [_vp, _vp, _vp, _i, _u64, _f, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "hyres_eb_dequant": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "hyres_refine_se_pool": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _vp]),
    "hyres_refine_se_scale_down": (_i, [_vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "hyres_refine_stats3_tc": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "hyres_replicate_border": (_i, [_vp, _i, _i, _i, _i, _vp]),
    "hyres_refine_spatial_att": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "hyres_jpeg_workspace_bytes": (_i64, [_i, _i, _i]),
    "hyres_jpeg_scan_words": (_i64, [_i, _i]),
    "hyres_jpeg_header_bytes": (_i, []),
    "hyres_jpeg_forward": (_i, [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _vp, _vp]),
    "hyres_jpeg_bpp": (_i, [_vp, _i, _i64, _vp, _vp]),
    "hyres_jpeg_assemble": (_i, [_vp, _i64, _i, _i, _i, _vp, _i64, C.POINTER(_i64)]),
    "hyres_nchw_f32_to_nhwc_bf16": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "hyres_nhwc_to_nchw_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "hyres_nhwc_bf16_to_nchw_f32": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "hyres_reduce_sqdiff": (_i, [_vp, _vp, _i64, _vp, _vp]),
    "hyres_reduce_log2": (_i, [_vp, _i64, _vp, _vp]),
    "hyres_rd_loss_finalize": (_i, [_vp, _vp, _vp, _vp, C.c_double, C.c_double, _f, _vp, _vp]),
    "hyres_pmf_to_quantized_cdf": (_i, [_vp, _i, _i, _vp]),
    "hyres_rans_encode_bound": (_i64, [_i64]),
    "hyres_rans_encode": (_i, [_vp, _vp, _i64, _vp, _i, _i, _vp, _vp, _vp, _i64, C.POINTER(_i64)]),
    "hyres_rans_decode": (_i, [_vp, _i64, _vp, _i64, _vp, _i, _i, _vp, _vp, _vp]),
    "hyres_rans_encode_batch": (_i, [_i, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _i]),
    "hyres_rans_decode_batch": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _i]),
    "hyres_rans_table_layout": (_i, [_vp, _i, _i, _vp, _vp, _vp]),
    "hyres_rans_encode_slots_batch": (_i, [_i, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _vp, _vp, _i]),
    "hyres_rans_decode_codes_batch": (_i, [_i, _vp, _vp, _vp, _vp, _vp, _i, _i, _vp, _vp, _vp, _i]),
    "hyres_rans_table_entries": (_i64, [_vp, _i, _i, _vp, _vp]),
    "hyres_rans_table_export": (_i, [_vp, _i, _i, _vp, _vp, _vp, _vp, _vp]),
    "hyres_rans_dev_encode": (_i, [_i, _vp, _vp, _i64, _vp, _i, _vp]),
    "hyres_set_reserved_sms": (_i, [_i]),
    "hyres_rans_dev_decode": (_i, [_vp, _vp, _vp, _vp, _i, _i64, _i, _vp, _vp, _i, _i64, _vp, _vp, _vp]),
}


def lib():
    """Load the shared library (once) and declare every prototype."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise HyresError(
                f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
                "(the hot path has no fallback implementation)"
            )
        handle = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(handle, name, None)
            if fn is None:  # reported by missing_symbols(); calling it raises AttributeError
                continue
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def missing_symbols():
    """Names declared in include/hyres_b200.h that the built library does not export."""
    handle = lib()
    return [n for n in SIGNATURES if not hasattr(handle, n)]


def check(rc, what=""):
    if rc != 0:
        msg = lib().hyres_last_error().decode(errors="replace")
        raise HyresError(f"{what} failed with code {rc}: {msg}")
