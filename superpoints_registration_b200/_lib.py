"""ctypes binding of libspr_b200.so (include/spr_b200.h).

There is NO fallback: if the library is missing or a call fails, a RuntimeError is raised -- the same
exception type the reference's C++ extensions raise (cpp_neighbors/wrapper.cpp:77,95,133,203).
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libspr_b200.so")

_lib = None

c_fp = c_void_p  # device pointers are passed as integers

_SIGNATURES = {
    # name: (restype, [argtypes])
    "spr_version": (c_int, []),
    "spr_last_error": (ctypes.c_char_p, []),
    "spr_launch_count": (ctypes.c_ulonglong, []),
    "spr_numeric_flags": (ctypes.c_uint, [c_int]),
    "spr_grid_subsample_workspace_bytes": (c_size_t, [c_int, c_int]),
    "spr_grid_subsample_batch": (c_int, [c_fp, c_fp, c_int, c_int, c_float, c_fp, c_fp, c_fp, c_fp, c_size_t, c_void_p]),
    "spr_grid_subsample_batch_ex": (c_int, [c_fp, c_fp, c_int, c_int, c_float, c_int, c_fp, c_fp, c_fp, c_fp, c_size_t,
                                            c_void_p]),
    "spr_cell_grid_workspace_bytes": (c_size_t, [c_int, c_int]),
    "spr_cell_grid_build": (c_int, [c_fp, c_fp, c_int, c_int, c_float, c_fp, c_size_t, c_void_p]),
    "spr_cell_grid_order": (c_int, [c_fp, c_int, c_int, c_fp, c_void_p]),
    "spr_radius_query": (c_int, [c_fp, c_fp, c_int, c_int, c_fp, c_int, c_float, c_int, c_fp, c_int, c_int, c_fp,
                                 c_void_p]),
    "spr_radius_query_ex": (c_int, [c_fp, c_fp, c_int, c_int, c_fp, c_int, c_float, c_int, c_int, c_fp, c_int, c_int, c_fp,
                                    c_void_p]),
    "spr_kpconv_workspace_bytes": (c_size_t, [c_int, c_int, c_int, c_int, c_int]),
    "spr_kpconv_forward": (c_int, [c_fp, c_fp, c_fp, c_int, c_int, c_int, c_fp, c_int, c_fp, c_int, c_fp, c_int,
                                   c_float, c_fp, c_int, c_int, c_int, c_fp, c_size_t, c_fp, c_void_p]),
    "spr_instance_norm_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "spr_instance_norm_lrelu": (c_int, [c_fp, c_fp, c_int, c_int, c_int, c_float, c_float, c_fp, c_fp, c_fp, c_size_t,
                                        c_void_p]),
    "spr_instance_norm_lrelu_ex": (c_int, [c_fp, c_fp, c_int, c_int, c_int, c_float, c_float, c_fp, c_fp, c_fp, c_float,
                                           c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_int, c_fp, c_size_t, c_void_p]),
    "spr_kpconv_weight_image_bytes": (c_size_t, [c_int]),
    "spr_kpconv_prepare_weights": (c_int, [c_fp, c_int, c_fp, c_fp, c_void_p]),
    "spr_kpconv_scratch_bytes": (c_size_t, [c_int, c_int]),
    "spr_kpconv_forward_prepared": (c_int, [c_fp, c_fp, c_int, c_int, c_int, c_fp, c_fp, c_fp, c_int, c_fp, c_fp, c_fp,
                                            c_float, c_fp, c_int, c_int, c_fp, c_fp, c_void_p]),
    "spr_kpconv_gather_supported": (c_int, [c_int, c_int]),
    "spr_kpconv_gather_weight_image_bytes": (c_size_t, [c_int]),
    "spr_kpconv_gather_scratch_bytes": (c_size_t, [c_int]),
    "spr_kpconv_gather_prepare_weights": (c_int, [c_fp, c_int, c_fp, c_fp, c_void_p]),
    "spr_kpconv_forward_gather": (c_int, [c_fp, c_fp, c_int, c_int, c_int, c_fp, c_fp, c_fp, c_int, c_fp, c_fp, c_fp,
                                          c_float, c_fp, c_int, c_int, c_fp, c_fp, c_void_p]),
    "spr_kpconv_staged_supported": (c_int, [c_int, c_int]),
    "spr_kpconv_staged_weight_image_bytes": (c_size_t, [c_int]),
    "spr_kpconv_staged_prepare_weights": (c_int, [c_fp, c_int, c_fp, c_fp, c_void_p]),
    "spr_kpconv_forward_staged": (c_int, [c_fp, c_fp, c_int, c_int, c_int, c_fp, c_fp, c_fp, c_int, c_fp, c_fp, c_fp,
                                          c_float, c_fp, c_int, c_int, c_fp, c_void_p]),
    "spr_max_pool": (c_int, [c_fp, c_fp, c_int, c_int, c_int, c_int, c_int, c_int, c_fp, c_fp, c_void_p]),
    "spr_match_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "spr_dual_softmax_match": (c_int, [c_fp, c_fp, c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int, c_int, c_int,
                                       c_fp, c_fp, c_fp, c_fp, c_fp, c_size_t, c_void_p]),
    "spr_sinkhorn_workspace_bytes": (c_size_t, [c_int, c_int, c_int]),
    "spr_sinkhorn_weighted_targets": (c_int, [c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int, c_int, c_fp, c_float,
                                              c_float, c_int, c_int, c_fp, c_fp, c_fp, c_size_t, c_void_p]),
    "spr_sinkhorn_affinity": (c_int, [c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_fp, c_fp,
                                      c_fp, c_fp, c_fp, c_size_t, c_void_p]),
    "spr_weighted_procrustes": (c_int, [c_fp, c_fp, c_fp, c_fp, c_int, c_fp, c_void_p]),
    "spr_gather_rows3": (c_int, [c_fp, c_int, c_fp, c_fp, c_int, c_fp, c_void_p]),
    "spr_top2_ratio": (c_int, [c_fp, c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_float, c_fp, c_fp, c_void_p]),
    "spr_inlier_reweight": (c_int, [c_fp, c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_float, c_fp, c_void_p]),
    "spr_select_hypothesis": (c_int, [c_fp, c_fp, c_fp, c_int, c_fp, c_int, c_fp, c_fp, c_fp, c_void_p]),
    "spr_split_f16": (c_int, [c_fp, c_int, c_int, c_int, c_fp, c_fp, c_int, c_int, c_float, c_void_p]),
    "spr_attention_varlen": (c_int, [c_fp, c_fp, c_int, c_int, c_int, c_int, c_int, c_int, c_fp, c_int, c_fp, c_int,
                                     c_fp, c_float, c_void_p]),
    "spr_attention_varlen_tc": (c_int, [c_fp, c_fp, c_int, c_int, c_int, c_int, c_int, c_int, c_fp, c_int, c_fp, c_int,
                                        c_fp, c_float, c_void_p]),
    "spr_cross_encoder_forward": (c_int, [c_fp, c_fp, c_int, c_int, c_int, c_int, c_int, c_fp, c_fp, c_fp, c_int, c_fp,
                                          c_int, c_fp, c_fp, c_fp, c_fp, c_float, c_int, c_fp, c_fp, c_float, c_fp,
                                          c_void_p]),
    "spr_gemm_a_image_bytes": (c_size_t, [c_int, c_int]),
    "spr_gemm_w_image_bytes": (c_size_t, [c_int, c_int]),
    "spr_gemm_prepare_weight": (c_int, [c_fp, c_int, c_int, c_float, c_fp, c_void_p]),
    "spr_gemm_prepare_input": (c_int, [c_fp, c_int, c_int, c_int, c_float, c_fp, c_void_p]),
    "spr_layernorm256_prepare": (c_int, [c_fp, c_fp, c_fp, c_fp, c_int, c_float, c_float, c_fp, c_fp, c_void_p]),
    "spr_gemm_tc": (c_int, [c_fp, c_fp, c_fp, c_fp, c_int, c_int, c_int, c_int, c_float, c_int, c_int, c_fp, c_fp,
                            c_int, c_int, c_float, c_float, c_fp, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib() -> ctypes.CDLL:
    """Load libspr_b200.so; raise loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m superpoints_registration_b200.build` "
                "(or __graft_entry__.build()). There is no CPU or PyTorch fallback for the hot path.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)  # AttributeError here means header and library disagree
            fn.restype = res
            fn.argtypes = args
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().spr_last_error()
        raise RuntimeError(f"{what} failed (code {rc}): {msg.decode() if msg else 'unknown error'}")


FLAG_FP16_OVERFLOW = 1


def numeric_flags(reset: bool = True) -> int:
    """Sticky numeric flags of the current device (spr_numeric_flags); synchronises the device."""
    return int(lib().spr_numeric_flags(1 if reset else 0))


def launch_count() -> int:
    return int(lib().spr_launch_count())
