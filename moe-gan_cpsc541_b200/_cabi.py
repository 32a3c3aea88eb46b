"""ctypes binding of libmoegan_sm100.so (C ABI declared in include/moegan_sm100.h).

No fallback: if the library is missing or a call fails this raises - the product path never routes
through a CPU or eager-PyTorch substitute.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int32, c_int64, c_void_p, POINTER

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MOEGAN_LIB_PATH") or os.path.join(HERE, "libmoegan_sm100.so")   # override: A/B builds in tools/

ROUTER_HID = 128
MAX_EXPERTS = 32
MAX_TOPK = 8
CHUNK = 32
KL_OUT_FLOATS = 2 + 256
ABI_VERSION = 6


class Shape(ctypes.Structure):
    """moegan_shape (include/moegan_sm100.h)."""
    _fields_ = [(n, c_int32) for n in
                ("T", "d", "Dt", "E", "k", "B", "tokens_per_sample", "precision", "dense", "hard", "m_cap")]


class BlockLayout(ctypes.Structure):
    """moegan_block_layout: byte offsets inside the forward workspace (-1 = absent)."""
    _fields_ = [(n, c_int64) for n in
                ("Wf", "Wt", "Wc", "Wf_op", "Wt_op", "tp", "tl", "fp", "raw", "topk_idx", "gates", "chunk_counts", "chunk_colsum",
                 "plan", "chunk_base", "x_copy", "a_op", "row_token", "row_gate", "token_rows", "pre", "h_op", "y", "eps_f", "eps_t", "eps_c",
                 "total")]


class BlockFwdArgs(ctypes.Structure):
    """moegan_block_fwd_args."""
    _fields_ = ([(n, c_void_p) for n in
                 ("x", "w", "mu_f", "rho_f", "eps_f", "mu_t", "rho_t", "eps_t", "mu_c", "rho_c", "eps_c", "temperature",
                  "W1_op", "b1s", "W2_op", "b2s", "prep_w1", "prep_b1", "prep_w2", "prep_b2")]
                + [("annealing_factor", c_float), ("sampled", c_int32), ("need_backward", c_int32)]
                + [(n, c_void_p) for n in ("out", "probs", "load", "kl_out", "workspace")]
                + [("workspace_bytes", c_int64), ("residual", c_void_p), ("residual_is_bf16", c_int32), ("out_is_bf16", c_int32),
                   ("x_op", c_void_p)])


class BlockBwdArgs(ctypes.Structure):
    """moegan_block_bwd_args."""
    _fields_ = ([(n, c_void_p) for n in
                 ("fwd_workspace", "w", "mu_f", "rho_f", "eps_f", "mu_t", "rho_t", "eps_t", "mu_c", "rho_c", "eps_c", "temperature",
                  "W1_op", "W2_op", "kl_out", "dout", "dkl", "dprobs", "dload")]
                + [("annealing_factor", c_float), ("sampled", c_int32)]
                + [(n, c_void_p) for n in
                   ("dx", "dw", "dmu_f", "drho_f", "dmu_t", "drho_t", "dmu_c", "drho_c", "dtemperature", "dW1", "db1", "dW2", "db2",
                    "workspace")]
                + [("workspace_bytes", c_int64), ("dout_is_bf16", c_int32), ("reserved", c_int32), ("x_op", c_void_p)])


_P = c_void_p
_SP = POINTER(Shape)

# name -> (restype, argtypes); every exported symbol of the header is listed here and checked at load time
SIGNATURES = {
    "moegan_version": (c_int32, []),
    "moegan_last_error_string": (c_char_p, []),
    "moegan_m_cap": (c_int64, [c_int32, c_int32, c_int32, c_int32]),
    "moegan_plan_ints": (c_int64, [c_int32]),
    "moegan_router_bwd_workspace_bytes": (c_int64, [_SP]),
    "moegan_router_sample_kl_fwd": (c_int32, [_SP] + [_P] * 9 + [c_int32] + [_P] * 6 + [_P]),
    "moegan_router_text_fwd": (c_int32, [_SP] + [_P] * 5 + [_P]),
    "moegan_router_fwd": (c_int32, [_SP] + [_P] * 5 + [c_float] + [_P] * 7 + [_P]),
    "moegan_dispatch_plan": (c_int32, [_SP] + [_P] * 5 + [_P]),
    "moegan_dispatch_gather": (c_int32, [_SP] + [_P] * 10 + [_P]),
    "moegan_convert_tokens": (c_int32, [_SP, _P, _P, _P]),
    "moegan_prep_weights": (c_int32, [_SP] + [POINTER(c_void_p)] * 4 + [_P] * 4 + [_P]),
    "moegan_ffn_fwd": (c_int32, [_SP] + [_P] * 9 + [_P]),
    "moegan_combine_fwd": (c_int32, [_SP, _P, _P, _P, _P, c_int32, _P, c_int32, _P]),
    "moegan_combine_bwd": (c_int32, [_SP, _P, _P, c_int32] + [_P] * 7 + [_P]),
    "moegan_ffn_bwd": (c_int32, [_SP] + [_P] * 14 + [_P, c_int64] + [_P]),
    "moegan_ffn_bwd_scratch_bytes": (c_int64, [_SP]),
    "moegan_uncombine_bwd": (c_int32, [_SP] + [_P] * 4 + [_P]),
    "moegan_router_bwd": (c_int32, [_SP] + [_P] * 16 + [c_float, c_int32] + [_P] * 9 + [_P] * 2 + [_P] * 9 + [_P, c_int64] + [_P]),
    "moegan_balance_loss_fwd": (c_int32, [_P, c_int32, c_int32, c_float, _P, _P]),
    "moegan_balance_loss_bwd": (c_int32, [_P, c_int32, c_int32, c_float, _P, _P, _P]),
    "moegan_block_fwd_workspace": (c_int64, [_SP, c_int32, POINTER(BlockLayout)]),
    "moegan_block_bwd_workspace": (c_int64, [_SP]),
    "moegan_block_fwd": (c_int32, [_SP, POINTER(BlockFwdArgs), _P]),
    "moegan_block_bwd": (c_int32, [_SP, POINTER(BlockBwdArgs), _P]),
    "moegan_layernorm_fwd": (c_int32, [_P, c_int32, _P, _P, c_int64, c_int32, c_float, _P, c_int32, _P, _P, _P]),
    "moegan_layernorm_fwd_dual": (c_int32, [_P, c_int32, _P, _P, c_int64, c_int32, c_float, _P, _P, _P, _P, _P]),
    "moegan_layernorm_bwd_scratch_bytes": (c_int64, [c_int64, c_int32]),
    "moegan_layernorm_bwd": (c_int32, [_P, c_int32, _P, c_int32, _P, _P, _P, c_int64, c_int32, _P, _P, _P, c_int64, _P]),
    "moegan_add_layernorm_fwd": (c_int32, [_P, _P, c_int32, _P, _P, c_int64, c_int32, c_float, _P, _P, c_int32, _P, _P, _P, _P]),
    "moegan_layernorm_bwd_add": (c_int32, [_P, c_int32, _P, c_int32, _P, _P, _P, c_int64, c_int32, _P, _P, _P, _P, c_int64, _P]),
    "moegan_channel_scale_fwd": (c_int32, [_P, c_int32, _P, c_int32, c_int32, c_int32, c_int32, c_float, _P, _P]),
    "moegan_channel_scale_bwd_scratch_bytes": (c_int64, [c_int32, c_int32, c_int32]),
    "moegan_channel_scale_bwd": (c_int32, [_P, _P, c_int32, _P, c_int32, c_int32, c_int32, c_int32, c_float, _P, _P, _P, c_int64, _P]),
    "moegan_colsum_scratch_bytes": (c_int64, [c_int64, c_int32]),
    "moegan_colsum": (c_int32, [_P, c_int32, c_int64, c_int32, _P, _P, c_int64, _P]),
    "moegan_demod_coef_fwd": (c_int32, [_P, c_int32, _P, c_int64, c_int64, c_int64, c_int64, c_int32, c_int32, c_int32, c_int32, c_int32, c_float, c_int32, _P, _P, _P]),
    "moegan_demod_coef_bwd": (c_int32, [_P, _P, _P, c_int32, _P, c_int64, c_int64, c_int64, c_int64, _P, c_int32, c_int32, c_int32, c_int32, c_int32, _P, _P, _P]),
    "moegan_offset_sample_fwd": (c_int32, [_P, c_int32, _P, c_int32, c_int32, c_int32, c_int32, _P, _P]),
    "moegan_offset_sample_bwd": (c_int32, [_P, _P, c_int32, _P, c_int32, c_int32, c_int32, c_int32, _P, _P, _P]),
    "moegan_upsample2x_fwd": (c_int32, [_P, c_int32, c_int32, c_int32, c_int32, c_int32, _P, _P]),
    "moegan_upsample2x_bwd": (c_int32, [_P, c_int32, c_int32, c_int32, c_int32, c_int32, _P, _P]),
    "moegan_umma_gemm": (c_int32, [c_int32] * 5 + [_P] * 3 + [_P]),
    "moegan_debug_set_mn_desc": (c_int32, [c_int32] * 3),
    "moegan_debug_launch_count": (c_int64, []),
    "moegan_stream_capture_id": (c_int64, [_P]),
    "moegan_debug_profile_enable": (c_int32, [c_int32]),
    "moegan_debug_profile_read": (c_int32, [_P, _P, c_int32]),
    "moegan_debug_profile_read_ex": (c_int32, [_P, _P, _P, c_int32]),
    "moegan_debug_set_side_streams": (c_int32, [c_int32]),
    "moegan_debug_set_fused_ffn": (c_int32, [c_int32]),
    "moegan_to_operand": (c_int32, [_P, c_int64, c_int64, c_int32, _P, _P]),
}

_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python {os.path.join(HERE, 'build.py')}` "
                "(or __graft_entry__.build()). There is no fallback path.")
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)       # AttributeError if the library does not export a declared symbol
            fn.restype = res
            fn.argtypes = args
        if l.moegan_version() != ABI_VERSION:
            raise RuntimeError(f"libmoegan_sm100.so ABI {l.moegan_version()} != expected {ABI_VERSION}; rebuild")
        _lib = l
    return _lib


def last_error() -> str:
    s = lib().moegan_last_error_string()
    return s.decode("utf-8", "replace") if s else ""


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed (status {rc}): {last_error()}")


def ptr(t) -> c_void_p:
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return c_void_p(0)
    return c_void_p(t.data_ptr())


def ptr_array(tensors) -> "ctypes.Array":
    arr = (c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


def m_cap(T: int, E: int, k: int, dense: bool) -> int:
    return int(lib().moegan_m_cap(T, E, k, 1 if dense else 0))


def make_shape(T, d, Dt, E, k, B, tokens_per_sample, precision, dense, hard) -> Shape:
    return Shape(T, d, Dt, E, k, B, tokens_per_sample, precision, 1 if dense else 0, 1 if hard else 0,
                 m_cap(T, E, k, dense))
