"""ctypes binding of libgrafp_sm100a.so (the C ABI declared in include/grafp.h).

There is no CPU or PyTorch fallback: if the shared library is missing the import of any
product module fails with an explicit error (run ``python -m neuralsampleid_b200.build`` or
``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgrafp_sm100a.so")

ACT_NONE, ACT_RELU, ACT_LEAKY, ACT_GELU, ACT_ELU, ACT_SIGMOID = 0, 1, 2, 3, 4, 5
ENGINE_AUTO, ENGINE_SIMT, ENGINE_TC_3XTF32, ENGINE_TC_TF32, ENGINE_TC_BF16X3, ENGINE_TC_BF16 = 0, 1, 2, 3, 4, 5
ENGINE_TC_F16X3 = 6
ENGINES = {"auto": ENGINE_AUTO, "simt": ENGINE_SIMT, "3xtf32": ENGINE_TC_3XTF32,
           "tf32": ENGINE_TC_TF32, "bf16x3": ENGINE_TC_BF16X3, "bf16": ENGINE_TC_BF16, "f16x3": ENGINE_TC_F16X3}


class GrafpError(RuntimeError):
    pass


class GemmArgs(C.Structure):
    _fields_ = [("a1", C.c_void_p), ("lda1", C.c_int64), ("k1", C.c_int32),
                ("a2", C.c_void_p), ("lda2", C.c_int64), ("k2", C.c_int32),
                ("w", C.c_void_p), ("ldw", C.c_int64), ("w_split", C.c_void_p),
                ("w_split_bf16", C.c_void_p),
                ("scale", C.c_void_p), ("shift", C.c_void_p),
                ("residual", C.c_void_p), ("ldr", C.c_int64),
                ("y", C.c_void_p), ("ldy", C.c_int64), ("row_sumsq", C.c_void_p),
                ("m", C.c_int64), ("n", C.c_int32), ("groups", C.c_int32),
                ("act", C.c_int32), ("act_param", C.c_float),
                ("tap3_nodes", C.c_int32), ("engine", C.c_int32),
                ("y_split", C.c_void_p), ("ldys", C.c_int64),
                ("a1_split", C.c_void_p), ("lda1s", C.c_int64),
                ("a2_gather_idx", C.c_void_p), ("a2_gather_nodes", C.c_int32), ("a2_gather_k", C.c_int32),
                ("w_split_f16", C.c_void_p), ("w_f16_unscale", C.c_float)]


_P, _I, _L, _F = C.c_void_p, C.c_int, C.c_int64, C.c_float
# name -> argtypes (every function returns int except the three noted below)
SIGNATURES = {
    "grafp_nchw_to_nodes": [_P, _P, _I, _I, _I, _P],
    "grafp_nodes_to_nchw": [_P, _P, _I, _I, _I, _P],
    "grafp_knn_fwd": [_P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _P, _P, C.c_size_t, _P],
    "grafp_mr_aggregate_fwd": [_P, _P, _I, _I, _I, _I, _P, _P, _P],
    "grafp_mr_aggregate_bwd": [_P, _P, _P, _I, _I, _I, _I, _P, _P],
    "grafp_index_select": [_P, _P, _I, _I, _I, _I, _P, _P],
    "grafp_nbr_reduce_fwd": [_P, _P, _I, _I, _I, _I, _I, _P, _P, _I, _F, _P, _P, _L, _P],
    "grafp_gemm_fwd": [C.POINTER(GemmArgs), _P],
    "grafp_gemm_tc_supported": [C.POINTER(GemmArgs)],
    "grafp_split_tf32": [_P, _L, _P, _P],
    "grafp_split_bf16": [_P, _L, _P, _P],
    "grafp_split_f16": [_P, _L, _F, _P, _P],
    "grafp_node_mean": [_P, _I, _I, _I, _P, _P],
    "grafp_ffn_fused_supported": [_L, _I, _I],
    "grafp_ffn_fused_fwd": [_P, _L, _L, _I, _I, _P, _L, _F, _P, _P, _I, _F, _P, _L, _F, _P, _P, _P, _L, _P],
    "grafp_mrconv_fc2_fused_supported": [_L, _I],
    "grafp_mrconv_fc2_fused_fwd": [_P, _L, _P, _L, _L, _I, _P, _L, _F, _P, _P, _I, _F, _P, _L, _F, _P, _P, _P, _L, _P,
                                   _L, _P],
    "grafp_frame_window_fwd": [_P, _L, _P, _I, _I, _L, _P, _P],
    "grafp_power_spectrum_fwd": [_P, _L, _L, _I, _I, _P, _L, _P],
    "grafp_amplitude_to_db_fwd": [_P, _L, _F, _F, _F, _P, _P],
    "grafp_unfold_segments_fwd": [_P, _L, _L, _I, _I, _I, _L, _P, _P],
    "grafp_topk_rows_fwd": [_P, _L, _I, _L, _L, _I, _I, _P, _P, _P],
    "grafp_topk_merge_fwd": [_P, _P, _I, _I, _I, _P, _P, _P, _P],
    "grafp_row_sumsq": [_P, _L, _I, _P, _P],
    "grafp_sequence_score_fwd": [_P, _I, _I, _P, _L, _P, _I, _P, _P],
    "grafp_nchw_to_nodes_add": [_P, _P, _P, _I, _I, _I, _P],
    "grafp_mha_pool_fwd": [_P, _L, _P, _L, _P, _L, _I, _I, _I, _I, _I, _F, _P, _L, _P],
    "grafp_stem_fwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _F, _P, _P],
    "grafp_peak_extract_fwd": [_P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P],
    "grafp_l2_normalize_rows": [_P, _L, _I, _F, _P, _P],
    "grafp_ntxent_fwd": [_P, _I, _I, _F, _I, _I, _P, _P, _P],
    "grafp_ntxent_bwd": [_P, _P, _I, _I, _F, _I, _I, _P, _P, _P],
    "grafp_col_stats": [_P, _L, _I, _L, _P, _P, _P],
    "grafp_bn_finalize": [_P, _P, _L, _I, _P, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P, _P],
    "grafp_affine_act": [_P, _L, _I, _L, _P, _P, _I, _F, _P, _L, _P, _L, _P],
    "grafp_bn_bwd_reduce": [_P, _L, _P, _L, _L, _I, _P, _P, _P, _P, _I, _F, _P, _P, _P],
    "grafp_bn_bwd_apply": [_P, _L, _P, _L, _L, _I, _P, _P, _P, _P, _I, _F, _I, _P, _P, _P, _L, _P, _P, _P],
    "grafp_bn_finalize_apply": [_P, _P, _L, _I, _P, _P, _P, _F, _F, _P, _P, _P, _P, _P, _P, _P, _L, _I, _F, _P, _L, _P, _L, _P],
    "grafp_bn_param_grad": [_P, _P, _I, _P, _P, _P],
    "grafp_gemm_wgrad": [_P, _L, _P, _L, _I, _P, _L, _I, _L, _I, _I, _I, _P, _L, _I, _P, C.c_size_t, _P],
    "grafp_tap3_bwd_input": [_P, _L, _I, _I, _P, _P],
    "grafp_node_mean_bwd": [_P, _I, _I, _I, _P, _P],
    "grafp_l2_normalize_rows_bwd": [_P, _P, _L, _I, _F, _P, _P],
    "grafp_peak_extract_bwd": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P, _P],
    "grafp_sq_norm": [_P, _L, _P, _P],
    "grafp_adam_clip_step": [_P, _P, _P, _P, _L, _F, _F, _F, _F, _I, _F, _P, _P],
    "grafp_adam_clip_step_dev": [_P, _P, _P, _P, _L, _P, _F, _F, _F, _P, _F, _P, _P, _P],
    "grafp_add_inplace": [_P, _P, _L, _P],
}
SPECIAL = {"grafp_knn_workspace_bytes": ([_I, _I, _I, _I, _I], C.c_size_t),
           "grafp_gemm_wgrad_workspace_bytes": ([_L, _I, _I, _I, _I, _I], C.c_size_t),
           "grafp_abi_version": ([], C.c_int), "grafp_last_error": ([], C.c_char_p),
           "grafp_launch_count": ([], C.c_int64)}

_lib = None
_cdll = None
_profiler = None      # optional callable(name, fn, args) -> rc, installed by bench.py / tests


def set_profiler(cb) -> None:
    """Installs (or clears, with None) a hook that wraps every C-ABI call: cb(name, fn, args) must
    call fn(*args) and return its result.  Used only for per-kernel timing in bench.py."""
    global _profiler
    _profiler = cb


class _Proxy:
    """Attribute access returns the ctypes function, routed through the profiler hook if set."""

    def __init__(self, cdll):
        self._cdll = cdll

    def __getattr__(self, name):
        fn = getattr(self._cdll, name)
        if _profiler is None:
            return fn
        return lambda *args: _profiler(name, fn, args)


def load():
    """Loads the shared library (once).  Raises GrafpError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GrafpError(
            "libgrafp_sm100a.so is not built (%s). There is no CPU/PyTorch fallback: run "
            "`python -m neuralsampleid_b200.build`." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    for name, (argtypes, restype) in SPECIAL.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = _Proxy(lib)
    return _lib


def exported_symbols():
    return list(SIGNATURES) + list(SPECIAL)


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().grafp_last_error()
        raise GrafpError("%s failed: %s" % (what, msg.decode() if msg else "unknown error"))


def launch_count() -> int:
    return int(load().grafp_launch_count())
