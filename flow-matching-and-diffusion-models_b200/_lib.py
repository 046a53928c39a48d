"""ctypes binding of the C-ABI library (include/fmdm_b200.h).

The library is the product: there is no eager/CPU fallback.  Importing this module never needs a GPU (so the
symbol/ABI tests run on CPU), but `lib()` raises loudly if the shared object has not been built, and every compute
entry point returns an error on a machine without an sm_100 device, which `check()` turns into RuntimeError.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "csrc" / "libfmdm_b200.so"

FM_CONV_MAX_SEG = 8


class ConvSeg(C.Structure):
    _fields_ = [
        ("src", C.c_void_p),
        ("C", C.c_int32),
        ("ksize", C.c_int32),
        ("upsample", C.c_int32),
        ("norm_act", C.c_int32),
        ("norm_a", C.c_void_p),
        ("norm_b", C.c_void_p),
        ("norm_stride", C.c_int32),
        ("_pad", C.c_int32),
    ]


class ConvParams(C.Structure):
    _fields_ = [
        ("seg", ConvSeg * FM_CONV_MAX_SEG),
        ("nseg", C.c_int32),
        ("B", C.c_int32),
        ("H", C.c_int32),
        ("W", C.c_int32),
        ("stride", C.c_int32),
        ("Cout", C.c_int32),
        ("_pad0", C.c_int32),
        ("weight", C.c_void_p),
        ("bias", C.c_void_p),
        ("addvec", C.c_void_p),
        ("addvec_stride", C.c_int32),
        ("_pad1", C.c_int32),
        ("residual", C.c_void_p),
        ("out", C.c_void_p),
        ("gn_stats", C.c_void_p),
        ("out_upsample", C.c_int32),
        ("_pad2", C.c_int32),
    ]


class PackEntry(C.Structure):
    _fields_ = [
        ("src", C.c_void_p),
        ("dst", C.c_void_p),
        ("dst_row_stride", C.c_int64),
        ("koff", C.c_int64),
        ("Cout", C.c_int32),
        ("Cin_total", C.c_int32),
        ("c_begin", C.c_int32),
        ("Cseg", C.c_int32),
        ("ksize", C.c_int32),
        ("mode", C.c_int32),
    ]


_vp, _i32, _i64, _f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_float

# name -> (restype, argtypes); must list every symbol declared in include/fmdm_b200.h
SIGNATURES = {
    "fm_version": (C.c_int, []),
    "fm_last_error": (C.c_char_p, []),
    "fm_launch_count": (C.c_longlong, []),
    "fm_set_pdl": (C.c_int, [C.c_int]),
    "fm_conv2d_igemm_bf16": (C.c_int, [C.POINTER(ConvParams), _vp]),
    "fm_conv_kernel_kind": (C.c_int, [C.POINTER(ConvParams)]),
    "fm_conv_operand_norm_supported": (C.c_int, [_i32, _i32, _i32, _i32]),
    "fm_groupnorm_affine_f32": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _i32, _i32, _i32, _vp, _vp]),
    "fm_groupnorm_finalize_partials_affine": (
        C.c_int,
        [_vp, _i32, _i32, _vp, _i32, _i32, _i32, _i64, _i32, _f32, _vp, _vp, _vp, _i64, _vp, _vp, _vp],
    ),
    "fm_conv_stats_rows": (C.c_int, [C.POINTER(ConvParams), C.POINTER(C.c_int32)]),
    "fm_groupnorm_finalize_partials": (C.c_int, [_vp, _i32, _i32, _vp, _i32, _i32, _i32, _i64, _i32, _f32, _vp, _vp]),
    "fm_weight_prepack_bf16": (C.c_int, [_vp, _i64, _i64, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "fm_weight_prepack_lo_bf16": (C.c_int, [_vp, _i64, _i64, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "fm_conv_stem_f32_bf16": (
        C.c_int, [_vp, _i32, _vp, _i32, _f32, _f32, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "fm_stem_im2col_bf16": (C.c_int, [_vp, _i32, _vp, _i32, _f32, _f32, _vp, _i32, _i32, _i32, _i32, _vp]),
    "fm_conv_stem_stats_rows": (C.c_int, [_i32, _i32, _i32, _i32]),
    "fm_conv_stem_tc_f32_bf16": (
        C.c_int, [_vp, _i32, _vp, _i32, _f32, _f32, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "fm_conv_stem_tc_stats_rows": (C.c_int, [_i32, _i32, _i32, _i32, _i32]),
    "fm_conv_head_bf16_f32": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp, _i32, _vp]),
    "fm_groupnorm_workspace_elems": (C.c_int64, [_i32, _i64, _i32, _i32]),
    "fm_groupnorm_stats_bf16": (C.c_int, [_vp, _i32, _vp, _i32, _i32, _i64, _i32, _f32, _vp, _vp, _vp]),
    "fm_groupnorm_apply_bf16": (
        C.c_int,
        [_vp, _i32, _vp, _i32, _i32, _i64, _i32, _vp, _vp, _vp, _vp, _i64, _i32, _vp, _vp],
    ),
    "fm_groupnorm_apply_partials_supported": (C.c_int, [_i32, _i32, _i32, _i32, _i32]),
    "fm_groupnorm_apply_partials_bf16": (
        C.c_int,
        [_vp, _i32, _vp, _i32, _i32, _i64, _i32, _vp, _i32, _vp, _i32, _f32, _vp, _vp, _vp, _i64, _i32, _vp, _vp],
    ),
    "fm_memset_f32": (C.c_int, [_vp, _i64, _vp]),
    "fm_upsample_nearest2x_bf16": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "fm_transpose_bf16": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _vp]),
    "fm_attention_bf16": (
        C.c_int,
        [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32] + [_i64] * 9 + [_f32, _vp],
    ),
    "fm_linear_attention_bf16": (
        C.c_int,
        [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32] + [_i64] * 9 + [_f32, _vp],
    ),
    "fm_context_kv_bf16": (C.c_int, [_vp] * 7 + [_i32, _i32, _i32, _i32, _i32, _f32, _i32, _vp]),
    "fm_timestep_embedding_f32": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _f32, _i32, _f32, _vp]),
    "fm_linear_f32": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "fm_sched_flowmatch_f32": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i64, _vp]),
    "fm_sched_ddim_f32": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _f32, _i64, _vp]),
    "fm_sched_ddpm_f32": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _f32, _i64, _vp]),
    "fm_sched_dpmpp2m_f32": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _vp]),
    "fm_sched_unipc_f32": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i64, _vp]),
    "fm_sched_add_noise_f32": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i32, _i64, _vp]),
    "fm_counter_add": (C.c_int, [_vp, _i32, _vp]),
    "fm_clamp_f32": (C.c_int, [_vp, _vp, _f32, _f32, _i64, _vp]),
    # training step (backward / loss / optimiser)
    "fm_weight_prepack_dgrad_bf16": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _i32, _vp]),
    "fm_weight_prepack_batch_block_elems": (C.c_int32, []),
    "fm_weight_prepack_batch_bf16": (C.c_int, [_vp, _vp, _vp, _i32, _vp]),
    "fm_conv_wgrad_workspace_elems": (C.c_int64, [_i32, _i32, _i32, _i32, _i32, _i32]),
    "fm_conv_wgrad_bf16": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _i32, _vp]),
    "fm_colsum_workspace_elems": (C.c_int64, [_i32, _i64, _i32]),
    "fm_colsum_bf16": (C.c_int, [_vp, _vp, _vp, _vp, _i32, _i64, _i32, _vp, _vp]),
    "fm_ticket_ints": (C.c_int32, []),
    "fm_zero_insert2x_bf16": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "fm_sumpool2x2_bf16": (C.c_int, [_vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "fm_groupnorm_bwd_workspace_elems": (C.c_int64, [_i32, _i64, _i32]),
    "fm_groupnorm_bwd_bf16": (
        C.c_int, [_vp, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _i64, _i32, _i32, _i64, _i32, _vp, _vp, _vp, _vp, _vp,
                  _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "fm_groupnorm_bwd_blocks": (C.c_int32, [_i32, _i64]),
    "fm_colsum_finish_f32": (C.c_int, [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp, _vp]),
    "fm_attention_bwd_cross_bf16": (
        C.c_int, [_vp] * 8 + [_i32, _i32, _i32, _i32, _i32] + [_i64] * 9 + [_f32, _vp]),
    "fm_linear_attention_bwd_bf16": (
        C.c_int, [_vp] * 7 + [_i32, _i32, _i32, _i32, _i32] + [_i64] * 9 + [_f32, _vp]),
    "fm_context_kv_bwd_workspace_elems": (C.c_int64, [_i32, _i32, _i32, _i32]),
    "fm_context_kv_bwd_f32": (C.c_int, [_vp] * 11 + [_i32, _i32, _i32, _i32, _i32, _vp]),
    "fm_attention_bwd_bf16": (
        C.c_int, [_vp] * 8 + [_i32, _i32, _i32, _i32] + [_i64] * 6 + [_f32, _vp]),
    "fm_linear_bwd_workspace_elems": (C.c_int64, [_i32, _i32, _i32]),
    "fm_linear_bwd_f32": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "fm_silu_bwd_f32": (C.c_int, [_vp, _vp, _vp, _i64, _vp]),
    "fm_conv_stem_wgrad_workspace_elems": (C.c_int64, [_i32, _i32]),
    "fm_conv_stem_wgrad_f32": (
        C.c_int, [_vp, _i32, _vp, _i32, _f32, _f32, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "fm_conv_head_bwd_workspace_elems": (C.c_int64, [_i32]),
    "fm_conv_head_bwd_f32": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]),
    "fm_sum_f32": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _i32, C.c_double, _vp]),
    "fm_mse_bwd_f32": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _i64, _vp]),
    "fm_adamw_f32": (C.c_int, [_vp, _vp, _vp, _vp, _i64, _f32, _f32, _f32, _f32, _f32, _i64, _f32, _vp]),
}

_LIB = None


def lib() -> C.CDLL:
    """Load (once) and return the shared library; raise if it is not built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = Path(os.environ.get("FMDM_B200_LIB", LIB_PATH))
    if not path.exists():
        raise RuntimeError(
            f"fmdm_b200: CUDA extension not built ({path} missing). Run `python -c 'import __graft_entry__ as g; "
            f"g.build()'` or `make -C {LIB_PATH.parent}`. There is no CPU fallback for the sampling hot path."
        )
    import torch  # noqa: F401  (maps libcudart.so.12, a DT_NEEDED of the library, into the process)

    handle = C.CDLL(str(path))
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(handle, name)  # AttributeError if the symbol is missing
        fn.restype = restype
        fn.argtypes = argtypes
    _LIB = handle
    return handle


def check(code: int, what: str) -> None:
    if code != 0:
        msg = lib().fm_last_error()
        raise RuntimeError(f"fmdm_b200.{what} failed (code {code}): {msg.decode(errors='replace') if msg else ''}")
