"""ctypes binding of libeqxv_b200.so (the C ABI declared in include/eqxv_b200.h).

There is deliberately no fallback: if the shared library is missing or no sm_100 device is
visible, every compute entry raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libeqxv_b200.so")

# eqxv_act
ACT_NONE, ACT_RELU, ACT_SILU, ACT_GELU_TANH, ACT_HARDSWISH, ACT_SIGMOID, ACT_HARDSIGMOID, ACT_RELU6 = range(8)
FLAG_OUT_F32 = 1
FLAG_RES_AFTER_ACT = 2
FLAG_GROUPED_BLOCK64 = 4
FLAG_K_TAIL_SHIFT = 8

ACT_BY_NAME = {
    None: ACT_NONE, "none": ACT_NONE, "identity": ACT_NONE, "relu": ACT_RELU, "silu": ACT_SILU,
    "gelu": ACT_GELU_TANH, "hard_swish": ACT_HARDSWISH, "sigmoid": ACT_SIGMOID,
    "hard_sigmoid": ACT_HARDSIGMOID, "relu6": ACT_RELU6,
}


class EqxvError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [
        ("x", C.c_void_p), ("wgt", C.c_void_p), ("bias", C.c_void_p), ("residual", C.c_void_p),
        ("y", C.c_void_p),
        ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("cin", C.c_int32), ("cout", C.c_int32),
        ("kh", C.c_int32), ("kw", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32), ("dil", C.c_int32),
        ("x_pitch", C.c_int32), ("y_pitch", C.c_int32), ("res_pitch", C.c_int32),
        ("act", C.c_int32), ("flags", C.c_int32),
    ]


class Bottleneck64Desc(C.Structure):
    _fields_ = [
        ("t1", C.c_void_p), ("w2", C.c_void_p), ("b2", C.c_void_p), ("w3", C.c_void_p), ("b3", C.c_void_p),
        ("residual", C.c_void_p), ("x0", C.c_void_p), ("y", C.c_void_p), ("w1n", C.c_void_p), ("b1n", C.c_void_p),
        ("next", C.c_void_p),
        ("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32),
        ("t1_pitch", C.c_int32), ("res_pitch", C.c_int32), ("x0_pitch", C.c_int32), ("y_pitch", C.c_int32),
        ("next_pitch", C.c_int32), ("next_channels", C.c_int32),
    ]


_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float

# name -> argtypes; every function returns int status unless listed in _NON_STATUS
SIGNATURES = {
    "eqxv_init": [C.c_int],
    "eqxv_conv2d_igemm_bf16": [C.POINTER(ConvDesc), _vp],
    "eqxv_bottleneck64_fused_bf16": [C.POINTER(Bottleneck64Desc), _vp],
    "eqxv_gemm_bias_act_res_bf16": [_vp, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _i32, _i32, _i32, _i32, _vp],
    "eqxv_gemm_res_rowstats_bf16": [_vp, _i64, _vp, _vp, _vp, _i64, _vp, _i64, _vp, _i64, _i32, _i32, _vp],
    "eqxv_gemm_ln_act_bf16": [_vp, _i64, _vp, _vp, _vp, _vp, _i32, _f32, _vp, _i64, _i64, _i32, _i32, _i32, _vp],
    "eqxv_gemm_gated_bf16": [_vp, _i64, _vp, _i64, _i32, _vp, _vp, _vp, _i64, _vp, _i64, _i64, _i32, _i32, _i32, _vp],
    "eqxv_conv_stem_bf16": [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp],
    "eqxv_conv_stem_maxpool_bf16": [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp],
    "eqxv_pack_stem_input": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp],
    "eqxv_pack_stem_input_c4": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp],
    "eqxv_conv_stem_c4_bf16": [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp],
    "eqxv_nchw_f32_to_nhwc_bf16": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp],
    "eqxv_nhwc_bf16_to_nchw_f32": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp],
    "eqxv_maxpool2d_nhwc_bf16": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp],
    "eqxv_maxpool2d_ceil_nhwc_bf16": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp],
    "eqxv_avgpool2d_nhwc_bf16": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp],
    "eqxv_adaptive_avgpool_nhwc_bf16": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp],
    "eqxv_layernorm_bf16": [_vp, _i64, _vp, _vp, _vp, _i64, _i64, _i32, _f32, _vp],
    "eqxv_attention_fwd_bf16": [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _f32, _vp],
    "eqxv_patchify_nchw_f32_bf16": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp],
    "eqxv_vit_assemble_tokens_bf16": [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _vp],
    "eqxv_gather_rows_bf16": [_vp, _i64, _vp, _i64, _i32, _i32, _i32, _i32, _vp],
    "eqxv_dwconv_bn_act_bf16": [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp],
    "eqxv_dwconv_pool_workspace_bytes": [_i32, _i32, _i32, _i32, _i32, _i32, _i32, C.POINTER(_i64)],
    "eqxv_dwconv_bn_act_pool_bf16": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32,
                                     _i32, _i32, _i32, _vp],
    "eqxv_dwconv_tile_bf16": [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp],
    "eqxv_eltwise_bf16": [_vp, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp],
    "eqxv_resize_bilinear_nhwc_bf16_to_nchw_f32": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp],
    "eqxv_resize_bilinear_nhwc_bf16": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp],
    "eqxv_copy2d_async": [_vp, _i64, _vp, _i64, _i64, _i64, _vp],
    "eqxv_window_attention_bf16": [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _f32, _vp],
    "eqxv_swin_v2_qk_normalize_bf16": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp],
    "eqxv_patch_merge_bf16": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp],
    "eqxv_debug_attention_timeline": [_vp],
    "eqxv_debug_bottleneck_timeline": [_vp],
    "eqxv_debug_stem_timeline": [_vp],
    "eqxv_u8hwc_to_nchw_f32": [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp],
    "eqxv_u8hwc_pack_stem_input": [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp],
    "eqxv_u8hwc_pack_stem_input_c4": [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp],
    "eqxv_u8hwc_to_nhwc_bf16": [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp],
    "eqxv_u8hwc_patchify_bf16": [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp],
    "eqxv_u8hwc_resize_bilinear": [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp],
    "eqxv_p2p_window_bytes": [_i64, C.POINTER(_i64)],
    "eqxv_p2p_alloc": [C.POINTER(_vp), _i64],
    "eqxv_p2p_free": [_vp],
    "eqxv_ipc_get_handle": [_vp, C.c_char_p],
    "eqxv_ipc_open_handle": [C.c_char_p, C.POINTER(_vp)],
    "eqxv_ipc_close_handle": [_vp],
    "eqxv_allgather_push": [_vp, _i64, C.POINTER(_vp), _i32, _i32, _i64, _i64, _vp],
    "eqxv_p2p_buffer_offset": [_i32, _i64, C.POINTER(_i64)],
    "eqxv_stream_create": [C.POINTER(_vp)],
    "eqxv_stream_destroy": [_vp],
    "eqxv_stream_sync": [_vp],
    "eqxv_graph_begin": [_vp],
    "eqxv_graph_end": [_vp, C.POINTER(_vp)],
    "eqxv_graph_launch": [_vp, _vp],
    "eqxv_graph_destroy": [_vp],
    "eqxv_event_create": [C.POINTER(_vp)],
    "eqxv_event_destroy": [_vp],
    "eqxv_event_record": [_vp, _vp],
    "eqxv_event_sync": [_vp],
    "eqxv_stream_wait_event": [_vp, _vp],
    "eqxv_event_elapsed_ms": [_vp, _vp, C.POINTER(_f32)],
    "eqxv_memcpy_h2d_async": [_vp, _vp, _i64, _vp],
    "eqxv_memcpy_d2h_async": [_vp, _vp, _i64, _vp],
    "eqxv_memcpy_async": [_vp, _vp, _i64, _vp],
    "eqxv_memset_async": [_vp, C.c_int, _i64, _vp],
}
_NON_STATUS = {"eqxv_version": C.c_char_p, "eqxv_last_error": C.c_char_p, "eqxv_sm_count": C.c_int}

_lib = None
_lock = threading.Lock()
_initialised_device = None
launch_count = 0  # number of kernel-launching C-ABI calls made by this process (bench bookkeeping)

_LAUNCHING = {
    "eqxv_conv2d_igemm_bf16", "eqxv_bottleneck64_fused_bf16", "eqxv_conv_stem_maxpool_bf16", "eqxv_conv_stem_c4_bf16", "eqxv_pack_stem_input_c4", "eqxv_u8hwc_pack_stem_input_c4", "eqxv_gemm_bias_act_res_bf16", "eqxv_conv_stem_bf16",
    "eqxv_pack_stem_input", "eqxv_nchw_f32_to_nhwc_bf16", "eqxv_nhwc_bf16_to_nchw_f32",
    "eqxv_maxpool2d_nhwc_bf16", "eqxv_maxpool2d_ceil_nhwc_bf16", "eqxv_avgpool2d_nhwc_bf16", "eqxv_adaptive_avgpool_nhwc_bf16",
    "eqxv_layernorm_bf16", "eqxv_attention_fwd_bf16", "eqxv_patchify_nchw_f32_bf16",
    "eqxv_vit_assemble_tokens_bf16", "eqxv_gather_rows_bf16", "eqxv_dwconv_bn_act_bf16", "eqxv_dwconv_tile_bf16", "eqxv_eltwise_bf16",
    "eqxv_resize_bilinear_nhwc_bf16_to_nchw_f32", "eqxv_resize_bilinear_nhwc_bf16", "eqxv_copy2d_async",
    "eqxv_window_attention_bf16", "eqxv_patch_merge_bf16",
    "eqxv_u8hwc_to_nchw_f32", "eqxv_u8hwc_pack_stem_input", "eqxv_u8hwc_to_nhwc_bf16", "eqxv_u8hwc_patchify_bf16",
    "eqxv_u8hwc_resize_bilinear", "eqxv_allgather_push", "eqxv_swin_v2_qk_normalize_bf16", "eqxv_dwconv_bn_act_pool_bf16", "eqxv_gemm_res_rowstats_bf16",
    "eqxv_gemm_ln_act_bf16", "eqxv_gemm_gated_bf16",
}


def load() -> C.CDLL:
    """dlopen the library and declare prototypes (no device needed)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise EqxvError(
                f"{LIB_PATH} is missing: build it with `python -m eqxvision_b200.csrc.build` "
                "(there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, restype in _NON_STATUS.items():
            fn = getattr(lib, name)
            fn.argtypes = []
            fn.restype = restype
        for name, argtypes in SIGNATURES.items():
            if not hasattr(lib, name):
                continue  # declared lazily: optional entry points are checked by tests
            fn = getattr(lib, name)
            fn.argtypes = argtypes
            fn.restype = C.c_int
        _lib = lib
        return lib


def last_error() -> str:
    return load().eqxv_last_error().decode()


def call(name: str, *args) -> None:
    """Invoke a status-returning entry point; raise EqxvError with the library's message."""
    global launch_count
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise EqxvError(f"{name} failed ({rc}): {last_error()}")
    if name in _LAUNCHING:
        launch_count += 1


def init(device: int = 0) -> None:
    global _initialised_device
    if _initialised_device == device:
        return
    call("eqxv_init", device)
    _initialised_device = device


def ptr(t) -> int:
    """device pointer of a torch tensor (or None)"""
    return None if t is None else t.data_ptr()
