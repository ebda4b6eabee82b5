"""ctypes binding of libnanovs.so (the C ABI declared in include/nanovs.h).

The library is built in-tree by ``nano_vs_slam_b200.build``; there is no fallback path: if the shared
object is missing or a call returns an error code, the caller gets an exception.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# NVS_LIB_PATH: load another build of the same library (A/B timing of kernel variants, debug builds)
LIB_PATH = os.environ.get("NVS_LIB_PATH") or os.path.join(HERE, "lib", "libnanovs.so")

NVS_OK = 0
ACT_NONE, ACT_LRELU, ACT_RELU, ACT_SIGMOID, ACT_TANH, ACT_SIGMOID_TANH, ACT_GELU = range(7)
OUT_PLAIN, OUT_POOL, OUT_BOTH, OUT_SHUFFLE = range(4)
IN_PLAIN, IN_S2D, IN_U8_HWC, IN_UNIT = range(4)

_vp, _i32, _f32, _f64, _sz, _i64 = C.c_void_p, C.c_int32, C.c_float, C.c_double, C.c_size_t, C.c_int64
_u64 = C.c_uint64


class NvsConvArgs(C.Structure):
    _fields_ = [
        ("src0", _vp), ("src1", _vp), ("weight", _vp), ("bias", _vp), ("dst", _vp), ("dst2", _vp),
        ("c0_total", _i32), ("c0_off", _i32), ("c0", _i32),
        ("c1_total", _i32), ("c1_off", _i32), ("c1", _i32),
        ("dst_c_total", _i32), ("dst_c_off", _i32),
        ("dst2_c_total", _i32), ("dst2_c_off", _i32),
        ("B", _i32), ("H", _i32), ("W", _i32),
        ("in_H", _i32), ("in_W", _i32),
        ("cout", _i32), ("ksize", _i32), ("act", _i32), ("out_mode", _i32), ("in_mode", _i32),
        ("dst_nhwc", _i32), ("dst2_nhwc", _i32),
    ]


class NvsConvTcArgs(C.Structure):
    _fields_ = [
        ("src0", _vp), ("src1", _vp), ("w_hi", _vp), ("w_lo", _vp), ("bias", _vp), ("dst", _vp), ("dst_pool", _vp),
        ("c0_total", _i32), ("c0_off", _i32), ("c0", _i32),
        ("c1_total", _i32), ("c1_off", _i32), ("c1", _i32),
        ("dst_c_total", _i32), ("dst_c_off", _i32), ("dst_layout", _i32), ("dst_mode", _i32),
        ("pool_c_total", _i32), ("pool_c_off", _i32),
        ("B", _i32), ("H", _i32), ("W", _i32), ("cout", _i32), ("act", _i32), ("flags", _i32),
        ("c0_real", _i32), ("c1_real", _i32), ("w_scale", C.c_float),
    ]


# name -> (restype, argtypes); this table is also what tests/test_cabi_symbols.py checks against the header
SIGNATURES = {
    "nvs_last_error": (C.c_char_p, []),
    "nvs_abi_version": (_i32, []),
    "nvs_device_ok": (_i32, []),
    "nvs_conv_cout_tile": (_i32, [_i32]),
    "nvs_conv_cin_chunk": (_i32, [_i32]),
    "nvs_conv": (_i32, [C.POINTER(NvsConvArgs), _vp]),
    "nvs_conv_tc_cout_pad": (_i32, [_i32]),
    "nvs_conv_tc_supported": (_i32, [_i32, _i32, _i32]),
    "nvs_conv_tc_plan_bytes": (_sz, []),
    "nvs_conv_tc_plan_init": (_i32, [_vp, C.POINTER(NvsConvTcArgs)]),
    "nvs_conv_tc_run": (_i32, [_vp, _vp, _vp, _vp]),
    "nvs_conv_rs_range_flag": (_i32, [_i32]),
    "nvs_conv_rs_debug_buffer": (None, [_vp]),
    "nvs_flat_debug_buffer": (None, [_vp]),
    "nvs_split16": (_i32, [_vp, _vp, C.c_int64, _i32, _vp]),
    "nvs_unsplit16": (_i32, [_vp, _vp, C.c_int64, _i32, _vp]),
    "nvs_conv_small": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "nvs_dwconv3x3": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "nvs_channel_layernorm": (_i32, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _f32, _vp]),
    "nvs_softmax_channels": (_i32, [_vp, _vp, _i32, _i32, _i32, _vp]),
    "nvs_attention": (_i32, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "nvs_netvlad_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "nvs_netvlad": (_i32, [_vp, _vp, _vp, _vp, _vp, _sz, _i32, _i32, _i32, _i32, _vp]),
    "nvs_preprocess_u8": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "nvs_gem": (_i32, [_vp, _vp, _i32, _i32, _i32, _i32, _f32, _f32, _vp]),
    "nvs_convap_workspace_bytes": (_sz, [_i32, _i32, _i32, _i32]),
    "nvs_convap": (_i32, [_vp, _vp, _vp, _vp, _vp, _sz, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "nvs_l2norm_channels": (_i32, [_vp, _vp, _i32, _i32, _i32, _vp]),
    "nvs_decode": (_i32, [_vp, _vp, _vp, _vp, _vp, _vp] + [_i32] * 9 + [_f32, _vp]),
    "nvs_seg_argmax": (_i32, [_vp, _vp, _vp] + [_i32] * 8 + [_vp]),
    "nvs_select_keypoints": (_i32, [_vp, _vp, _vp, _vp, _vp, _i32, _f32, _i32, _vp, _vp, _vp, _vp, _vp, _vp,
                                    _i32, _i32, _i32, _vp]),
    "nvs_match_workspace_bytes": (_sz, [_i32, _i32]),
    "nvs_match": (_i32, [_vp, _vp, _i32, _i32, _i32, _f64, _i32, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "nvs_match_batch_workspace_bytes": (_sz, [_i32, _i32]),
    "nvs_match_batch": (_i32, [_vp, _vp, _i32, _i32, _i32, _vp, _vp, _i32, _f64, _i32, _vp, _vp, _vp, _vp, _vp, _sz,
                                _vp]),
    "nvs_pose_workspace_bytes": (_sz, [_i32, _i32, _i32]),
    "nvs_pose_batch": (_i32, [_vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _f32, _f32, _f32, _f32, _f32, _i32, _u64,
                               _i32, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "nvs_pose_batch_adaptive": (_i32, [_vp, _i32, _i32, _vp, _vp, _vp, _vp, _vp, _i32, _f32, _f32, _f32, _f32, _f32, _i32,
                                        _u64, _i32, _f32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _sz, _vp]),
    "nvs_flat_padded_dim": (_i32, [_i32]),
    "nvs_flat_max_k": (_i32, []),
    "nvs_flat_list_slots": (_i32, [C.c_int64, _i32, _i32, _i32, _vp]),
    "nvs_flat_prepare": (_i32, [_vp, _i64, _i32, _vp, _vp, _vp, _vp]),
    "nvs_flat_search_workspace_bytes": (_sz, [_i64, _i32, _i32, _i32]),
    "nvs_flat_search": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _i32, _i32, _i32, _i64, _vp, _vp, _vp, _sz, _vp, _vp,
                                _vp]),
    "nvs_flat_search_begin": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _i32, _i32, _i32, _vp, _vp, _sz, _vp, _vp, _vp]),
    "nvs_flat_bound_merge": (_i32, [_vp, _i32, _i32, _i32, _vp, _vp]),
    "nvs_flat_search_end": (_i32, [_vp, _vp, _vp, _vp, _i64, _vp, _i32, _i32, _i32, _i64, _vp, _vp, _vp, _vp, _sz, _vp]),
    "nvs_topk_merge": (_i32, [_vp, _vp, _i32, _i32, _i32, _i64, _i64, _vp, _vp, _vp]),
}

_lib = None


class NanovsError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load libnanovs.so (once).  Raises if it has not been built -- there is no CPU fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NanovsError(
                f"{LIB_PATH} not found: run `python -m nano_vs_slam_b200.build` "
                "(the CUDA extension is the only implementation; there is no fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def check(rc: int, what: str = "") -> None:
    if rc != NVS_OK:
        msg = lib().nvs_last_error()
        codes = {-1: "NVS_ERR_ARG", -2: "NVS_ERR_CUDA", -3: "NVS_ERR_UNSUPPORTED", -4: "NVS_ERR_NO_DEVICE"}
        raise NanovsError(f"{what} failed: {codes.get(rc, rc)} {msg.decode() if msg else ''}")
