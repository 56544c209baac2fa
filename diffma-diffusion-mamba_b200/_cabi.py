"""ctypes binding of ``libdiffma_b200.so`` -- the structs and prototypes of ``include/diffma_b200.h``.

This is the binding a reference-side maintainer would write (INTEGRATION.md shows it used from
``block/mamba.py``).  No torch types cross this boundary: raw device pointers, sizes, strides and the
stream handle.  There is NO fallback: if the library is missing and cannot be built, ``lib()`` raises.
"""
from __future__ import annotations

import ctypes as C
import os

DM_ABI_VERSION = 6
DM_OK, DM_ERR_INVALID_ARG, DM_ERR_UNSUPPORTED, DM_ERR_CUDA = 0, -1, -2, -4
DM_F32, DM_BF16 = 0, 1
DM_MAX_GROUPS = 4
DM_OUT_SCAN_ORDER, DM_OUT_TOKEN_ORDER = 0, 1

EXPORTS = ("dm_mamba1_scan_fwd", "dm_mamba1_scan_phase", "dm_mamba1_sched_workspace_bytes", "dm_mamba1_scan_bwd", "dm_mamba1_bwd_chunk_tokens", "dm_mamba2_ssd_fwd", "dm_mamba2_ssd_bwd", "dm_merge_directions_multi", "dm_spiral_pre", "dm_spiral_post_ln",
           "dm_spiral_post_mix", "dm_spiral_post_mix_pre", "dm_spiral_pre_bwd", "dm_spiral_post_mix_bwd", "dm_spiral_post_ln_bwd",
           "dm_merge_directions", "dm_gemm_bf16_tn", "dm_gemm_bf16_tn_ex", "dm_p_sample_update", "dm_adamw_ema_step", "dm_adamw_ema_step_ex", "dm_spiral_post_mix_fold", "dm_step_head", "dm_final_linear_unpatchify", "dm_version", "dm_status_string", "dm_last_cuda_error",
           "dm_build_info")


class Mamba1Group(C.Structure):
    _fields_ = [
        ("xz", C.c_void_p), ("xz_batch_stride", C.c_int64), ("xz_token_stride", C.c_int64),
        ("out", C.c_void_p), ("out_batch_stride", C.c_int64), ("out_dir_stride", C.c_int64),
        ("out_token_stride", C.c_int64),
        ("u", C.c_void_p), ("x_dbl", C.c_void_p),
        ("conv_weight", C.c_void_p), ("conv_bias", C.c_void_p), ("x_proj_weight", C.c_void_p),
        ("dt_proj_weight", C.c_void_p), ("dt_bias", C.c_void_p), ("A", C.c_void_p), ("D", C.c_void_p),
        ("chunk_states", C.c_void_p), ("delta", C.c_void_p),
    ]


class Mamba1Args(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("n_dir", C.c_int32), ("seqlen", C.c_int32),
        ("d_inner", C.c_int32), ("d_state", C.c_int32), ("dt_rank", C.c_int32), ("d_conv", C.c_int32),
        ("act_dtype", C.c_int32), ("out_order", C.c_int32), ("n_groups", C.c_int32),
        ("order", C.c_void_p),
        ("group", Mamba1Group * DM_MAX_GROUPS),
        ("sched_workspace", C.c_void_p), ("sched_workspace_bytes", C.c_int64),
        ("z_is_gated", C.c_int32), ("reserved_", C.c_int32),
    ]


class GemmArgs(C.Structure):
    _fields_ = [
        ("A", C.c_void_p), ("a_group_stride", C.c_int64), ("a_row_stride", C.c_int64), ("a_sum_stride", C.c_int64),
        ("n_sum", C.c_int32), ("reserved_", C.c_int32),
        ("B", C.c_void_p), ("b_group_stride", C.c_int64), ("b_row_stride", C.c_int64),
        ("C", C.c_void_p), ("c_group_stride", C.c_int64), ("c_row_stride", C.c_int64),
        ("row_scale", C.c_void_p), ("bias", C.c_void_p),
        ("silu_from", C.c_int32), ("groups", C.c_int32), ("M", C.c_int32), ("N", C.c_int32), ("K", C.c_int32),
    ]


class Mamba1BwdGroup(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("dout", "d_xz_scan", "du", "ddelta", "d_x_dbl", "dA", "dD", "d_dt_bias",
                                          "state_workspace", "d_conv_weight", "d_conv_bias")] + [("states_valid", C.c_int64)]


class Mamba2Group(C.Structure):
    _fields_ = [
        ("zxbcdt", C.c_void_p), ("in_batch_stride", C.c_int64), ("in_token_stride", C.c_int64),
        ("out", C.c_void_p), ("out_batch_stride", C.c_int64), ("out_dir_stride", C.c_int64),
        ("out_token_stride", C.c_int64),
        ("sumsq", C.c_void_p), ("sumsq_batch_stride", C.c_int64), ("sumsq_dir_stride", C.c_int64),
        ("conv_weight", C.c_void_p), ("conv_bias", C.c_void_p), ("dt_bias", C.c_void_p), ("A", C.c_void_p),
        ("D", C.c_void_p),
    ]


class Mamba2BwdGroup(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("u", "x_dbl", "d_x_dbl", "d_bc", "d_conv_weight", "d_conv_bias")]


class MergeSegment(C.Structure):
    _fields_ = [("src", C.c_void_p), ("channels", C.c_int32), ("row_stride", C.c_int32)]


class Mamba2Args(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("n_dir", C.c_int32), ("seqlen", C.c_int32),
        ("d_inner", C.c_int32), ("d_state", C.c_int32), ("nheads", C.c_int32), ("d_conv", C.c_int32),
        ("act_dtype", C.c_int32), ("out_order", C.c_int32), ("n_groups", C.c_int32), ("gate", C.c_int32),
        ("order", C.c_void_p),
        ("group", Mamba2Group * DM_MAX_GROUPS),
    ]


class SpiralFoldArgs(C.Structure):
    """dm_spiral_fold_args (include/diffma_b200.h)."""
    _fields_ = [
        ("x", C.c_void_p), ("skip", C.c_void_p), ("ab", C.c_void_p), ("g2", C.c_void_p), ("g2_dtype", C.c_int32),
        ("act_dtype", C.c_int32), ("colsum", C.c_void_p), ("cvec", C.c_void_p), ("w3", C.c_void_p), ("b3", C.c_void_p),
        ("mod", C.c_void_p), ("mod_batch_stride", C.c_int64), ("x_out", C.c_void_p), ("skip_next", C.c_void_p),
        ("ln_weight", C.c_void_p), ("ln_bias", C.c_void_p), ("mod_next", C.c_void_p), ("mod_next_batch_stride", C.c_int64),
        ("w", C.c_void_p), ("out2", C.c_void_p), ("batch", C.c_int32), ("seqlen", C.c_int32), ("d_model", C.c_int32),
        ("eps", C.c_float), ("ln2_eps", C.c_float),
    ]


class AdamwArgs(C.Structure):
    """dm_adamw_args (include/diffma_b200.h)."""
    _fields_ = [
        ("param", C.c_void_p), ("grad", C.c_void_p), ("exp_avg", C.c_void_p), ("exp_avg_sq", C.c_void_p),
        ("ema", C.c_void_p), ("step", C.c_void_p), ("shadow_bf16", C.c_void_p), ("n", C.c_int64),
        ("lr", C.c_double), ("beta1", C.c_double), ("beta2", C.c_double), ("eps", C.c_double),
        ("weight_decay", C.c_double), ("ema_decay", C.c_double), ("grad_scale", C.c_double),
    ]


_LIB = None


def lib_path() -> str:
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libdiffma_b200.so")


def lib() -> C.CDLL:
    """Load (building first if stale and nvcc is present) the C-ABI library.  Raises if unavailable."""
    global _LIB
    if _LIB is not None:
        return _LIB
    from . import build as _build
    path = _build.ensure_built()
    if not os.path.exists(path):
        raise RuntimeError(f"diffma_b200: CUDA library {path} is missing and could not be built; "
                           "there is no CPU fallback")
    L = C.CDLL(path)
    L.dm_version.restype = C.c_int
    L.dm_status_string.restype = C.c_char_p
    L.dm_status_string.argtypes = [C.c_int]
    L.dm_last_cuda_error.restype = C.c_int
    L.dm_build_info.restype = C.c_char_p
    L.dm_mamba1_scan_fwd.restype = C.c_int
    L.dm_mamba1_scan_fwd.argtypes = [C.POINTER(Mamba1Args), C.c_void_p]
    L.dm_mamba1_sched_workspace_bytes.restype = C.c_int64
    L.dm_mamba1_sched_workspace_bytes.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_int32]
    L.dm_mamba1_scan_phase.restype = C.c_int
    L.dm_mamba1_scan_phase.argtypes = [C.POINTER(Mamba1Args), C.c_int, C.c_void_p]
    L.dm_mamba1_bwd_chunk_tokens.restype = C.c_int
    L.dm_mamba1_scan_bwd.restype = C.c_int
    L.dm_mamba1_scan_bwd.argtypes = [C.POINTER(Mamba1Args), C.POINTER(Mamba1BwdGroup), C.c_int, C.c_void_p]
    L.dm_mamba2_ssd_fwd.restype = C.c_int
    L.dm_mamba2_ssd_fwd.argtypes = [C.POINTER(Mamba2Args), C.c_void_p]
    vp, i32, i64, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    L.dm_spiral_pre.restype = C.c_int
    L.dm_spiral_pre.argtypes = [vp, vp, vp, vp, vp, i64, vp, vp, i32, i32, i32, f32, i32, vp]
    L.dm_spiral_post_ln.restype = C.c_int
    L.dm_spiral_post_ln.argtypes = [vp, vp, vp, vp, i32, i32, f32, i32, vp]
    L.dm_spiral_post_mix.restype = C.c_int
    L.dm_spiral_post_mix.argtypes = [vp, vp, vp, vp, vp, vp, vp, i64, vp, i32, i32, i32, i32, vp]
    L.dm_spiral_post_mix_pre.restype = C.c_int
    L.dm_spiral_post_mix_pre.argtypes = [vp, vp, vp, vp, vp, vp, vp, i64, vp, vp, vp, vp, vp, i64, vp, vp, i32, i32, i32, f32,
                                         i32, vp]
    L.dm_spiral_pre_bwd.restype = C.c_int
    L.dm_spiral_pre_bwd.argtypes = [vp, vp, vp, vp, vp, i64, vp, vp, vp, vp, i64, vp, vp, i32, i32, i32, f32, i32, vp]
    L.dm_spiral_post_mix_bwd.restype = C.c_int
    L.dm_spiral_post_mix_bwd.argtypes = [vp, vp, vp, vp, vp, vp, i64, vp, vp, vp, i64, vp, vp, i32, i32, i32, i32, vp]
    L.dm_spiral_post_ln_bwd.restype = C.c_int
    L.dm_spiral_post_ln_bwd.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, f32, i32, vp]
    L.dm_merge_directions.restype = C.c_int
    L.dm_merge_directions.argtypes = [vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]
    L.dm_merge_directions_multi.restype = C.c_int
    L.dm_merge_directions_multi.argtypes = [C.POINTER(MergeSegment), i32, vp, vp, i32, i32, i32, i32, i32, vp]
    L.dm_mamba2_ssd_bwd.restype = C.c_int
    L.dm_mamba2_ssd_bwd.argtypes = [C.POINTER(Mamba2Args), C.POINTER(Mamba2BwdGroup), C.c_int, vp]
    L.dm_gemm_bf16_tn.restype = C.c_int
    L.dm_gemm_bf16_tn.argtypes = [vp, i64, i64, vp, i64, i64, vp, i64, i64, vp, i32, i32, i32, i32, vp]
    L.dm_gemm_bf16_tn_ex.restype = C.c_int
    L.dm_gemm_bf16_tn_ex.argtypes = [C.POINTER(GemmArgs), vp]
    L.dm_p_sample_update.restype = C.c_int
    L.dm_p_sample_update.argtypes = [vp, vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, vp]
    L.dm_adamw_ema_step.restype = C.c_int
    f64 = C.c_double
    L.dm_adamw_ema_step.argtypes = [vp, vp, vp, vp, vp, vp, i64, f64, f64, f64, f64, f64, f64, f64, vp]
    L.dm_final_linear_unpatchify.restype = C.c_int
    L.dm_final_linear_unpatchify.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]
    L.dm_step_head.restype = C.c_int
    L.dm_step_head.argtypes = [vp, vp, vp, vp, i32, i32, i32, i32, vp, vp, i32, vp, vp, i32, vp, i32, i32, vp]
    L.dm_spiral_post_mix_fold.restype = C.c_int
    L.dm_spiral_post_mix_fold.argtypes = [C.POINTER(SpiralFoldArgs), vp]
    L.dm_adamw_ema_step_ex.restype = C.c_int
    L.dm_adamw_ema_step_ex.argtypes = [C.POINTER(AdamwArgs), vp]
    if L.dm_version() != DM_ABI_VERSION:
        raise RuntimeError(f"diffma_b200: library ABI {L.dm_version()} != binding ABI {DM_ABI_VERSION}; rebuild")
    _LIB = L
    return L


def check(status: int, what: str) -> None:
    if status != DM_OK:
        L = lib()
        msg = L.dm_status_string(status).decode()
        extra = f" (cudaError {L.dm_last_cuda_error()})" if status == DM_ERR_CUDA else ""
        raise RuntimeError(f"{what}: {msg}{extra}")
