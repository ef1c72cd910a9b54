"""ctypes binding of the C-ABI CUDA library (include/founddiff_b200.h).

There is NO CPU or PyTorch fallback: if the shared library is missing it is built in-tree with nvcc
(founddiff_b200/build.py); if that fails, importing the ops raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_long, c_void_p

from . import build as _build

FD_F32, FD_BF16, FD_F16 = 0, 1, 2

EXPORTS = [
    "fd_version", "fd_program_arena_bytes", "fd_program_load", "fd_program_buffer", "fd_program_num_launches", "fd_unet_step", "fd_sample_step", "fd_program_destroy", "fd_ln_fold", "fd_gram_ws_floats", "fd_conv_gn_ws_floats", "fd_gn_stats_ws_floats", "fd_dwconv3x3_silu_tm", "fd_x_proj_tm", "fd_scan_tm_segments", "fd_scan_tm_plan", "fd_selective_scan_tm", "fd_scan_tm_chain_plan", "fd_selective_scan_tm_chained", "fd_selective_scan_fwd", "fd_selective_scan_fwd_merge", "fd_selective_scan_fwd_merge_xdbl", "fd_selective_scan_fwd_merge_cl", "fd_x_proj_tc", "fd_avgpool2x2_nhwc", "fd_slice_metrics", "fd_init_conv7x7_tc", "fd_ln_modulate_io", "fd_row_rstd", "fd_ln_gate", "fd_ln_gate_out_proj", "fd_ln_gate_out_proj_supported", "fd_conv2d_simt", "fd_conv2d_tc_supported", "fd_conv2d_tc_plan_create",
    "fd_conv2d_tc_run", "fd_conv2d_tc_plan_destroy", "fd_init_conv7x7", "fd_ln_modulate", "fd_dwconv3x3_silu_scan",
    "fd_xdt_proj", "fd_xdt_proj_tc", "fd_merge_ln_gate", "fd_dwconv3x3_qkv_gram", "fd_dwconv3x3_nhwc", "fd_gram_qk", "fd_attn_weff", "fd_gn_stats", "fd_gn_silu_add", "fd_gn_scale_shift_silu", "fd_flash_attn_d32", "fd_flash_attn_d32_tc", "fd_linattn_context", "fd_linattn_weff", "fd_softmax_d32",
    "fd_linear_small", "fd_time_sinusoid", "fd_sampler_init", "fd_final_conv_update", "fd_final_conv_update_obj", "fd_unnormalize", "fd_ddpm_update",
]


class ConvParams(Structure):
    """Mirror of fd_conv_params."""
    _fields_ = [
        ("src0", c_void_p), ("src1", c_void_p), ("weight", c_void_p), ("bias", c_void_p), ("gate", c_void_p),
        ("addend", c_void_p), ("out", c_void_p), ("gn_sums", c_void_p), ("weight_up4", c_void_p), ("gn_ws", c_void_p),
        ("ln_v", c_void_p), ("ln_rstd", c_void_p),
        ("c0", c_int), ("c1", c_int), ("ld0", c_int), ("B", c_int), ("Hin", c_int), ("Win", c_int), ("Cout", c_int),
        ("KH", c_int), ("KW", c_int), ("stride", c_int), ("pad", c_int), ("upsample", c_int),
        ("silu_from", c_int), ("gate_stride", c_int), ("gn_groups", c_int), ("per_batch_weight", c_int),
        ("dtype", c_int), ("relu_out", c_int), ("ln_eps", c_float), ("ab_dtype_p1", c_int),
    ]


class FdError(RuntimeError):
    pass


_lib = None


def lib_path() -> str:
    return _build.LIB


PROXY = None          # set by program.Recorder: a wrapper around the loaded library that records every launch it forwards


def load():
    """Load (building first if needed) the CUDA library.  Raises if it cannot be had — no fallback."""
    global _lib
    if _lib is not None:
        return PROXY if PROXY is not None else _lib
    path = _build.LIB
    if _build.needs_build():
        try:
            _build.build()
        except Exception as e:  # stale-but-present library is still usable (e.g. no nvcc on this box)
            if not os.path.exists(path):
                raise FdError(f"founddiff_b200: CUDA library missing and nvcc build failed: {e}") from e
    lib = ctypes.CDLL(path)
    V, I, F, L = c_void_p, c_int, c_float, c_long
    sig = {
        "fd_selective_scan_fwd": [V] * 8 + [I] * 7 + [V],
        "fd_selective_scan_fwd_merge": [V] * 8 + [I] * 7 + [V],
        "fd_selective_scan_fwd_merge_xdbl": [V] * 7 + [I] * 8 + [V],
        "fd_selective_scan_fwd_merge_cl": [V] * 8 + [I] * 7 + [V],
        "fd_x_proj_tc": [V] * 3 + [I] * 6 + [V],
        "fd_dwconv3x3_silu_tm": [V, I, V, V, V, I, I, I, I, I, V],
        "fd_x_proj_tm": [V] * 6 + [I] * 8 + [V],
        "fd_scan_tm_segments": [I, I, I, I],
        "fd_scan_tm_plan": [I] * 6,
        "fd_selective_scan_tm": [V] * 8 + [L, V] + [I] * 8 + [V],
        "fd_scan_tm_chain_plan": [I] * 6 + [POINTER(c_int)],
        "fd_selective_scan_tm_chained": [V] * 7 + [L, V] + [I] * 7 + [V],
        "fd_avgpool2x2_nhwc": [V, V, I, I, I, I, I, V],
        "fd_slice_metrics": [V, V, V, I, I, I, F, V],
        "fd_ln_gate": [V, V, I, I, V, V, V, V, I, I, I, F, I, V],
        "fd_row_rstd": [V, V, L, I, F, I, V],
        "fd_ln_gate_out_proj": [V, V, I, I, V, V, V, V, V, I, V, V, I, I, I, I, F, I, I, V],
        "fd_ln_gate_out_proj_supported": [I] * 7,
        "fd_conv2d_simt": [POINTER(ConvParams), V],
        "fd_conv2d_tc_supported": [POINTER(ConvParams)],
        "fd_conv2d_tc_plan_create": [POINTER(ConvParams), POINTER(c_void_p)],
        "fd_conv2d_tc_run": [V, V],
        "fd_init_conv7x7": [V] * 5 + [I] * 5 + [V],
        "fd_init_conv7x7_tc": [V] * 5 + [I] * 5 + [V],
        "fd_ln_modulate": [V] * 6 + [I] * 4 + [F, I, V],
        "fd_ln_fold": [V] * 5 + [I] + [V] * 2 + [I] * 4 + [V],
        "fd_ln_modulate_io": [V] * 6 + [I] * 4 + [F, I, I, V],
        "fd_dwconv3x3_silu_scan": [V, I, V, V, V, I, I, I, I, I, V],
        "fd_xdt_proj": [V] * 6 + [I] * 6 + [V],
        "fd_xdt_proj_tc": [V] * 6 + [I] * 7 + [V, I, I, V],
        "fd_merge_ln_gate": [V, V, I, I, V, V, V, V, V, I, I, I, I, F, I, V],
        "fd_dwconv3x3_qkv_gram": [V] * 6 + [I] * 5 + [V],
        "fd_attn_weff": [V] * 5 + [I] * 3 + [V],
        "fd_dwconv3x3_nhwc": [V] * 4 + [I] * 6 + [V],
        "fd_gram_qk": [V, I, V, V, V, I, I, I, I, V],
        "fd_gn_stats": [V, V, V, I, I, I, I, I, V],
        "fd_gn_silu_add": [V] * 6 + [I] * 4 + [F, I, V],
        "fd_linear_small": [V] * 5 + [I] * 5 + [V],
        "fd_gn_scale_shift_silu": [V] * 6 + [I, V, V] + [I] * 4 + [F, I, V],
        "fd_flash_attn_d32": [V, V, I, I, I, F, I, V],
        "fd_flash_attn_d32_tc": [V, V, I, I, I, F, I, V],
        "fd_linattn_context": [V, V, V, V, I, I, I, I, V],
        "fd_linattn_weff": [V, V, V, V, I, I, I, I, F, I, V],
        "fd_softmax_d32": [V, V, I, I, I, I, V],
        "fd_time_sinusoid": [V, V, I, I, V],
        "fd_sampler_init": [V, V, F, V, V, V, L, V],
        "fd_final_conv_update": [V] * 11 + [L, I, I, V],
        "fd_final_conv_update_obj": [V] * 14 + [L, I, I, I, V],
        "fd_unnormalize": [V, V, L, V],
        "fd_ddpm_update": [V] * 6 + [L, V],
    }
    for name, argtypes in sig.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = c_int
    lib.fd_conv_gn_ws_floats.argtypes = [c_int]
    lib.fd_conv_gn_ws_floats.restype = c_long
    lib.fd_program_arena_bytes.argtypes = [c_char_p]
    lib.fd_program_arena_bytes.restype = c_long
    lib.fd_program_load.argtypes = [c_char_p, c_void_p, c_long, POINTER(c_void_p)]
    lib.fd_program_load.restype = c_int
    lib.fd_program_buffer.argtypes = [c_void_p, c_char_p, POINTER(c_long)]
    lib.fd_program_buffer.restype = c_void_p
    lib.fd_program_num_launches.argtypes = [c_void_p]
    lib.fd_program_num_launches.restype = c_int
    lib.fd_unet_step.argtypes = [c_void_p, c_void_p]
    lib.fd_unet_step.restype = c_int
    lib.fd_sample_step.argtypes = [c_void_p, c_void_p]
    lib.fd_sample_step.restype = c_int
    lib.fd_program_destroy.argtypes = [c_void_p]
    lib.fd_program_destroy.restype = None
    lib.fd_gram_ws_floats.argtypes = [c_int] * 5
    lib.fd_gram_ws_floats.restype = c_long
    lib.fd_gn_stats_ws_floats.argtypes = [c_int, c_int, c_int]
    lib.fd_gn_stats_ws_floats.restype = c_long
    lib.fd_version.restype = c_char_p
    lib.fd_version.argtypes = []
    lib.fd_conv2d_tc_plan_destroy.argtypes = [c_void_p]
    lib.fd_conv2d_tc_plan_destroy.restype = None
    _lib = lib
    return PROXY if PROXY is not None else lib


def check(rc: int, what: str):
    if rc != 0:
        kind = {-1: "bad argument", -2: "unsupported shape/dtype", -3: "CUDA driver entry point unavailable"}.get(
            rc, f"CUDA error {rc}")
        raise FdError(f"{what}: {kind}")
