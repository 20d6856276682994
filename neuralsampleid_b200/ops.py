"""Thin functional layer over the C ABI: torch tensors in (device memory + stream plumbing only),
raw pointers out.  All activations here are NODE-MAJOR fp32: (B*N, C) row-major.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import _lib
from ._lib import (ACT_ELU, ACT_GELU, ACT_LEAKY, ACT_NONE, ACT_RELU, ACT_SIGMOID, ENGINES, GemmArgs,
                   GrafpError, check)

_ACTS = {None: ACT_NONE, "none": ACT_NONE, "relu": ACT_RELU, "leakyrelu": ACT_LEAKY,
         "gelu": ACT_GELU, "elu": ACT_ELU, "sigmoid": ACT_SIGMOID}

_engine = ENGINES[os.environ.get("GRAFP_ENGINE", "auto").lower()]
# test/bench hook: a tensor-core engine name applied to every GEMM whose shape the tensor-core
# kernels take (the others keep the exact SIMT engine), e.g. "3xtf32", "bf16"
_engine_override = None


def set_engine(name: str) -> None:
    """Selects the GEMM engine: 'auto' (tcgen05 f16x3 where shapes allow, else fp32 SIMT), 'simt',
    '3xtf32', 'tf32', 'f16x3', 'bf16x3', 'bf16'.  An explicit tensor-core engine applies to every GEMM
    whose shape the tcgen05 kernels take; the others (stem, odd widths) keep the exact SIMT kernel.  The
    kNN Gram tiles always use 3xTF32 on the tensor-core engines.  Prepared weights carry every split."""
    global _engine
    _engine = ENGINES[name.lower()]


_SPLIT16 = (_lib.ENGINE_AUTO, _lib.ENGINE_TC_BF16X3, _lib.ENGINE_TC_BF16, _lib.ENGINE_TC_F16X3)


def _effective_engine() -> int:
    return ENGINES[_engine_override] if _engine_override is not None else _engine


def get_engine() -> int:
    return _engine


def act_code(name) -> int:
    key = name.lower() if isinstance(name, str) else name
    if key not in _ACTS:
        raise NotImplementedError("activation layer [%s] is not found" % name)
    return _ACTS[key]


def _stream(t: torch.Tensor):
    return C.c_void_p(torch.cuda.current_stream(t.device).cuda_stream)


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else None


def _chk(t: torch.Tensor, dtype=torch.float32, name="tensor") -> torch.Tensor:
    if not t.is_cuda:
        raise GrafpError("%s must be a CUDA tensor (there is no CPU path)" % name)
    if t.dtype != dtype:
        raise GrafpError("%s must be %s, got %s" % (name, dtype, t.dtype))
    return t if t.is_contiguous() else t.contiguous()


def nchw_to_nodes(x: torch.Tensor) -> torch.Tensor:
    """(B, C, N[,1]) -> (B*N, C)."""
    x = _chk(x, name="x")
    B, Cc, N = x.shape[0], x.shape[1], x.shape[2]
    out = torch.empty((B * N, Cc), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        check(_lib.load().grafp_nchw_to_nodes(_ptr(x), _ptr(out), B, Cc, N, _stream(x)), "nchw_to_nodes")
    return out


def nodes_to_nchw(x: torch.Tensor, B: int, N: int) -> torch.Tensor:
    """(B*N, C) -> (B, C, N)."""
    x = _chk(x, name="x")
    Cc = x.shape[1]
    out = torch.empty((B, Cc, N), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        check(_lib.load().grafp_nodes_to_nchw(_ptr(x), _ptr(out), B, Cc, N, _stream(x)), "nodes_to_nchw")
    return out


def knn(x: torch.Tensor, B: int, N: int, k: int, dilation: int = 1, normalize: bool = True,
        return_dist: bool = False, engine: Optional[int] = None, row_sumsq: Optional[torch.Tensor] = None):
    """Dense dilated kNN over node-major features -> int32 (B, N, k) [, fp32 (B, N, k)]."""
    x = _chk(x, name="x")
    Cc = x.shape[1]
    idx = torch.empty((B, N, k), device=x.device, dtype=torch.int32)
    dist = torch.empty((B, N, k), device=x.device, dtype=torch.float32) if return_dist else None
    eng = _engine if engine is None else engine
    # the Gram tiles always run 3xTF32: every tensor-core GEMM engine maps onto it (shapes the tcgen05 kNN
    # does not take fall back to the exact SIMT kernel under AUTO)
    if eng in (_lib.ENGINE_TC_TF32, _lib.ENGINE_TC_BF16X3, _lib.ENGINE_TC_BF16, _lib.ENGINE_TC_F16X3):
        eng = _lib.ENGINE_AUTO
    lib = _lib.load()
    ws_bytes = int(lib.grafp_knn_workspace_bytes(B, N, Cc, k, dilation)) if eng != _lib.ENGINE_SIMT else 0
    ws = torch.empty((ws_bytes // 4,), device=x.device, dtype=torch.float32) if ws_bytes else None
    with torch.cuda.device(x.device):
        check(lib.grafp_knn_fwd(_ptr(x), B, N, Cc, k, dilation, int(normalize), eng, _ptr(row_sumsq), _ptr(idx),
                                _ptr(dist), _ptr(ws), ws_bytes, _stream(x)), "knn_fwd")
    return (idx, dist) if return_dist else idx


def knn_engine(B: int, N: int, Cc: int, k: int, dilation: int = 1) -> str:
    """Which kernel family ``knn`` runs for this shape under the default engine: "tcgen05" (knn_tc.cu / knn_big.cu:
    the library asks for a workspace) or "simt" (the exact fp32 kernel)."""
    if _engine == _lib.ENGINE_SIMT:
        return "simt"
    return "tcgen05" if int(_lib.load().grafp_knn_workspace_bytes(B, N, Cc, k, dilation)) > 0 else "simt"


def mr_aggregate(x: torch.Tensor, idx: torch.Tensor, B: int, N: int, want_arg: bool = False):
    """m[n, c] = max_k (x[idx[n, k], c] - x[n, c]);  optional uint8 arg-max ranks."""
    x = _chk(x, name="x")
    idx = _chk(idx, torch.int32, "idx")
    Cc, k = x.shape[1], idx.shape[-1]
    m = torch.empty_like(x)
    arg = torch.empty((B * N, Cc), device=x.device, dtype=torch.uint8) if want_arg else None
    with torch.cuda.device(x.device):
        check(_lib.load().grafp_mr_aggregate_fwd(_ptr(x), _ptr(idx), B, N, Cc, k, _ptr(m), _ptr(arg),
                                                 _stream(x)), "mr_aggregate_fwd")
    return (m, arg) if want_arg else m


NBR_MAX, NBR_SUM_SELF, NBR_EDGE_MAX = 1, 2, 3


def nbr_reduce(x: torch.Tensor, idx: torch.Tensor, B: int, N: int, mode: int, scale=None, shift=None, act=None,
               act_param: float = 0.0, eps: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None):
    """Per-node neighbour reductions of the other GraphConv2d variants (include/grafp.h):
    NBR_MAX max_k x_j; NBR_SUM_SELF (1+eps) x_i + sum_k x_j; NBR_EDGE_MAX max_k act(scale (x_j - x_i) + shift).
    ``out`` may be a column slice of a wider row-major matrix (row stride = out.stride(0))."""
    x = _chk(x, name="x")
    idx = _chk(idx, torch.int32, "idx")
    Cc, k = x.shape[1], idx.shape[-1]
    if out is None:
        out = torch.empty_like(x)
    elif out.shape != x.shape or out.stride(1) != 1 or not out.is_cuda or out.dtype != torch.float32:
        raise GrafpError("nbr_reduce: out must be an fp32 CUDA (M, C) view with unit column stride")
    with torch.cuda.device(x.device):
        check(_lib.load().grafp_nbr_reduce_fwd(_ptr(x), _ptr(idx), B, N, Cc, k, mode, _ptr(scale), _ptr(shift),
                                               act_code(act), act_param, _ptr(eps), _ptr(out), out.stride(0),
                                               _stream(x)), "nbr_reduce_fwd")
    return out


def mr_aggregate_bwd(dm: torch.Tensor, idx: torch.Tensor, arg: torch.Tensor, B: int, N: int,
                     dx: torch.Tensor) -> torch.Tensor:
    """Accumulates the aggregation gradient into dx (in place)."""
    dm = _chk(dm, name="dm")
    Cc, k = dm.shape[1], idx.shape[-1]
    with torch.cuda.device(dm.device):
        check(_lib.load().grafp_mr_aggregate_bwd(_ptr(dm), _ptr(idx), _ptr(arg), B, N, Cc, k, _ptr(dx),
                                                 _stream(dm)), "mr_aggregate_bwd")
    return dx


def index_select(x: torch.Tensor, idx: torch.Tensor, B: int, N: int) -> torch.Tensor:
    """batched_index_select: (B*N, C) + (B, N', k) -> (B, C, N', k)."""
    x = _chk(x, name="x")
    idx = _chk(idx, torch.int32, "idx")
    Cc, k = x.shape[1], idx.shape[-1]
    if idx.shape[1] != N:
        raise GrafpError("index_select: idx must list every node of the graph")
    out = torch.empty((B, Cc, N, k), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        check(_lib.load().grafp_index_select(_ptr(x), _ptr(idx), B, N, Cc, k, _ptr(out), _stream(x)),
              "index_select")
    return out


def split_tf32(w: torch.Tensor) -> torch.Tensor:
    """(n, k) fp32 -> (2n, k) stacked [tf32(w) ; tf32(w - tf32(w))] for the 3xTF32 engine."""
    w = _chk(w, name="w")
    out = torch.empty((2 * w.shape[0], w.shape[1]), device=w.device, dtype=torch.float32)
    with torch.cuda.device(w.device):
        check(_lib.load().grafp_split_tf32(_ptr(w), w.numel(), _ptr(out), _stream(w)), "split_tf32")
    return out


def split_bf16(w: torch.Tensor) -> torch.Tensor:
    """(n, k) fp32 -> bf16 (2n, k) stacked [bf16(w) ; bf16(w - bf16(w))] for the bf16 engines."""
    w = _chk(w, name="w")
    out = torch.empty((2 * w.shape[0], w.shape[1]), device=w.device, dtype=torch.bfloat16)
    with torch.cuda.device(w.device):
        check(_lib.load().grafp_split_bf16(_ptr(w), w.numel(), _ptr(out), _stream(w)), "split_bf16")
    return out


def f16_prescale(w: torch.Tensor) -> float:
    """Power of two 2^s that brings max|w| to [2^9, 2^10): the fp16 lo parts of the f16x3 engine are then normal
    numbers for every weight above 2^-4 of the largest, and hi stays 64x under the half overflow."""
    amax = float(w.abs().max()) if w.numel() else 0.0
    if not (amax > 0.0) or amax != amax or amax == float("inf"):
        return 1.0
    import math
    s = 9 - math.floor(math.log2(amax))
    return float(2.0 ** max(-14, min(24, s)))


def split_f16(w: torch.Tensor, prescale: float) -> torch.Tensor:
    """(n, k) fp32 -> fp16 (2n, k) stacked [f16(w * prescale) ; f16(w * prescale - hi)] for the f16x3 engine."""
    w = _chk(w, name="w")
    out = torch.empty((2 * w.shape[0], w.shape[1]), device=w.device, dtype=torch.float16)
    with torch.cuda.device(w.device):
        check(_lib.load().grafp_split_f16(_ptr(w), w.numel(), float(prescale), _ptr(out), _stream(w)), "split_f16")
    return out


TRAIN_F16_PRESCALE = 256.0


def tc_splits(w: torch.Tensor) -> dict:
    """The split copies of a weight that the engine in use needs, as keyword arguments of ``gemm``.  This is the
    per-step path of the train mode (weights change every step), so the f16x3 pre-scale is the fixed 2^8 instead
    of a data-dependent one (no host synchronisation, CUDA-graph capturable): lo parts stay normal numbers for
    |w| >= 1e-3, weights saturate beyond |w| = 511."""
    e = _effective_engine()
    out = {}
    if e == _lib.ENGINE_TC_3XTF32:
        out["w_split"] = split_tf32(w)
    if e in (_lib.ENGINE_TC_BF16X3, _lib.ENGINE_TC_BF16):
        out["w_split_bf16"] = split_bf16(w)
    if e in (_lib.ENGINE_AUTO, _lib.ENGINE_TC_F16X3):
        out["w_split_f16"] = split_f16(w, TRAIN_F16_PRESCALE)
        out["f16_unscale"] = 1.0 / TRAIN_F16_PRESCALE
    return out


class SplitAct:
    """An activation in the split 16-bit format of include/grafp.h (ABI 2 / 4): a bf16 (bf16 engines) or fp16
    (f16x3 engine) tensor (2, M, C), plane 0 = r(v), plane 1 = r(v - r(v)) -- the operand pair the 3-pass engine
    computes with.  Only a GEMM on the same engine may consume it (``ops.gemm`` / ``ops.linear`` as ``a1``)."""
    __slots__ = ("t",)

    def __init__(self, t: torch.Tensor):
        self.t = t

    @property
    def shape(self):
        return self.t.shape[1:]

    @property
    def device(self):
        return self.t.device

    def float(self) -> torch.Tensor:
        """fp32 value hi (+ lo) (test / debugging helper)."""
        return self.t[0].float() + (self.t[1].float() if self.t.shape[0] > 1 else 0.0)


def split_dtype(engine: Optional[int] = None):
    """Element type of SplitAct planes under ``engine`` (default: the engine in effect)."""
    eng = _effective_engine() if engine is None else engine
    return torch.float16 if eng in (_lib.ENGINE_AUTO, _lib.ENGINE_TC_F16X3) else torch.bfloat16


def split_ok(lin, k_total: int) -> bool:
    """True when the GEMM over ``lin`` (k_total input columns) runs on a 16-bit-operand tensor-core engine, i.e.
    may produce or consume a SplitAct."""
    eng = _effective_engine()
    if eng not in _SPLIT16:
        return False
    have = lin.w_split_f16 if eng in (_lib.ENGINE_AUTO, _lib.ENGINE_TC_F16X3) else lin.w_split_bf16
    if have is None or os.environ.get("GRAFP_NO_SPLIT_ACT"):
        return False
    n = lin.w.shape[0] // lin.groups
    return k_total % (32 * lin.groups) == 0 and n % 32 == 0


def fused_mr_ok(lin, c: int) -> bool:
    """True when MRConv2d's gather + max-relative can run inside the GEMM over ``lin`` (the dual-source
    BasicConv weights, C input channels per source): bf16 tensor-core engine, tcgen05-supported shape."""
    # Opt-in (GRAFP_FUSED_MR=1): measured on B200 the in-kernel gather (four transform warps, dependent
    # idx -> row loads from L2) makes the MRConv GEMMs 3-4x slower (5.6 vs 1.4 ms at stage 3), far more than the
    # 1.3 ms/step of mr_aggregate it removes; a register-blocked variant that batches the loads spills.
    if not os.environ.get("GRAFP_FUSED_MR") or _effective_engine() not in (_lib.ENGINE_TC_BF16X3, _lib.ENGINE_TC_BF16):
        return False
    return lin.groups > 0 and split_ok(lin, 2 * c) and (c // lin.groups) % 32 == 0


def linear(a1, lin, act=None, act_param: float = 0.0, residual=None, a2=None,
           tap3_nodes: int = 0, engine: Optional[int] = None, row_sumsq=None, out=None, out_split: bool = False,
           a2_gather=None):
    """ops.gemm over a prepared ``_prep.Linear``."""
    return gemm(a1, lin.w, lin.scale, lin.shift, act, act_param, residual, a2, lin.groups, tap3_nodes,
                engine, out, lin.w_split, lin.w_split_bf16, row_sumsq, out_split, a2_gather,
                lin.w_split_f16, lin.f16_unscale)


def gemm(a1: torch.Tensor, w: torch.Tensor, scale: Optional[torch.Tensor] = None,
         shift: Optional[torch.Tensor] = None, act=None, act_param: float = 0.0,
         residual: Optional[torch.Tensor] = None, a2: Optional[torch.Tensor] = None,
         groups: int = 1, tap3_nodes: int = 0, engine: Optional[int] = None,
         out: Optional[torch.Tensor] = None, w_split: Optional[torch.Tensor] = None,
         w_split_bf16: Optional[torch.Tensor] = None, row_sumsq: Optional[torch.Tensor] = None,
         out_split: bool = False, a2_gather=None, w_split_f16: Optional[torch.Tensor] = None,
         f16_unscale: float = 0.0):
    """y = act(scale * [a1 | a2] @ w.T + shift) + residual  (per-group, see include/grafp.h).
    ``a2_gather`` = (idx int32 (B, N, k), N): the second source is the max-relative aggregation of a1 over
    those neighbour lists, computed inside the kernel (fused MRConv2d; bf16 tensor-core engines only).

    a1: (M, groups*k1) (or the (2M', Cin) node matrix in tap3 mode), a2: (M, groups*k2) or None,
    w: (groups*n, k1+k2).  ``a1`` may be a SplitAct; ``out_split=True`` returns one, ``out_split="both"`` returns
    (fp32 tensor, SplitAct) written by the same epilogue (bf16 tensor-core engines only: the library refuses
    anything else)."""
    a1s = None
    # the engine this call resolves to (an explicit tensor-core engine only applies where the tcgen05 kernels
    # take the shape: decided below, once the arguments are assembled)
    eng_req = engine if engine is not None else _effective_engine()
    sdt = split_dtype(eng_req)
    if isinstance(a1, SplitAct):
        a1s = _chk(a1.t, sdt, "a1 (split)")
        a1 = None
    else:
        a1 = _chk(a1, name="a1")
    w = _chk(w, name="w")
    n_total, ktot = w.shape
    n = n_total // groups
    if a1s is not None:
        if tap3_nodes > 0 or a2 is not None:
            raise GrafpError("gemm: a SplitAct operand cannot be combined with tap3 / a second source")
        k1, k2, M = a1s.shape[2] // groups, 0, a1s.shape[1]
    elif tap3_nodes > 0:
        cin = a1.shape[1]
        k1, k2 = 3 * cin, 0
        M = a1.shape[0] // 2
    else:
        k1 = a1.shape[1] // groups
        k2 = 0
        M = a1.shape[0]
        if a2 is not None:
            a2 = _chk(a2, name="a2")
            k2 = a2.shape[1] // groups
        elif a2_gather is not None:
            k2 = k1
    if k1 + k2 != ktot:
        raise GrafpError("gemm: weight has %d columns, operands give %d" % (ktot, k1 + k2))
    dev = a1.device if a1 is not None else a1s.device
    # the engine a split GEMM runs on (split_ok() admitted only bf16 tensor-core engines): the 1-pass bf16
    # engine carries the hi plane only
    planes = 1 if eng_req == _lib.ENGINE_TC_BF16 else 2
    if a1s is not None and a1s.shape[0] < planes:
        raise GrafpError("gemm: a hi-plane-only SplitAct can only feed the 1-pass bf16 engine")
    both = out_split == "both"                       # fp32 output AND its split copy, from one epilogue
    out32 = None
    if out_split:
        if out is not None or row_sumsq is not None:
            raise GrafpError("gemm: out_split cannot be combined with out= / row_sumsq")
        out = torch.empty((planes, M, n_total), device=dev, dtype=sdt)
        if both:
            out32 = torch.empty((M, n_total), device=dev, dtype=torch.float32)
    elif out is None:
        out = torch.empty((M, n_total), device=dev, dtype=torch.float32)
    elif out.shape != (M, n_total) or out.stride(1) != 1:
        raise GrafpError("gemm: out must be an (M, groups*n) view with unit column stride")
    args = GemmArgs()
    if a1s is not None:
        args.a1, args.lda1, args.k1 = None, 0, k1
        args.a1_split, args.lda1s = a1s.data_ptr(), a1s.stride(1)
    else:
        args.a1, args.lda1, args.k1 = a1.data_ptr(), a1.stride(0), k1
        args.a1_split, args.lda1s = None, 0
    args.a2, args.lda2, args.k2 = (a2.data_ptr() if a2 is not None else None), \
        (a2.stride(0) if a2 is not None else 0), k2
    if a2_gather is not None:
        gidx = _chk(a2_gather[0], torch.int32, "a2_gather idx")
        args.a2_gather_idx, args.a2_gather_nodes, args.a2_gather_k = gidx.data_ptr(), int(a2_gather[1]), gidx.shape[-1]
    else:
        args.a2_gather_idx, args.a2_gather_nodes, args.a2_gather_k = None, 0, 0
    args.w, args.ldw = w.data_ptr(), w.stride(0)
    args.w_split = w_split.data_ptr() if w_split is not None else None
    args.w_split_bf16 = w_split_bf16.data_ptr() if w_split_bf16 is not None else None
    args.w_split_f16 = w_split_f16.data_ptr() if w_split_f16 is not None else None
    args.w_f16_unscale = float(f16_unscale) if w_split_f16 is not None else 0.0
    args.scale = scale.data_ptr() if scale is not None else None
    args.shift = shift.data_ptr() if shift is not None else None
    if residual is not None:
        residual = _chk(residual, name="residual")
        args.residual, args.ldr = residual.data_ptr(), residual.stride(0)
    else:
        args.residual, args.ldr = None, 0
    if out_split:
        args.y, args.ldy = (out32.data_ptr(), out32.stride(0)) if both else (None, 0)
        args.y_split, args.ldys = out.data_ptr(), out.stride(1)
    else:
        args.y, args.ldy = out.data_ptr(), out.stride(0)
        args.y_split, args.ldys = None, 0
    args.row_sumsq = row_sumsq.data_ptr() if row_sumsq is not None else None
    args.m, args.n, args.groups = M, n, groups
    args.act, args.act_param = act_code(act) if not isinstance(act, int) else act, act_param
    args.tap3_nodes = tap3_nodes
    args.engine = eng_req
    if engine is None and eng_req not in (_lib.ENGINE_AUTO, _lib.ENGINE_SIMT) and a1s is None and not out_split \
            and a2_gather is None:
        # a tensor-core engine selected globally (set_engine / GRAFP_ENGINE): shapes, or unprepared weights, that
        # the tcgen05 kernels do not take run the exact SIMT kernel, as under AUTO.  An engine passed per call
        # (engine=) is taken literally and fails loudly instead.
        need = {_lib.ENGINE_TC_3XTF32: w_split, _lib.ENGINE_TC_TF32: w, _lib.ENGINE_TC_BF16X3: w_split_bf16,
                _lib.ENGINE_TC_BF16: w_split_bf16, _lib.ENGINE_TC_F16X3: w_split_f16}[eng_req]
        if need is None or not _lib.load().grafp_gemm_tc_supported(C.byref(args)):
            args.engine = _lib.ENGINE_SIMT
    with torch.cuda.device(dev):
        check(_lib.load().grafp_gemm_fwd(C.byref(args), _stream(out)), "gemm_fwd")
    if both:
        return out32, SplitAct(out)
    return SplitAct(out) if out_split else out


def ffn_fused_ok(lin1, lin2, x) -> bool:
    """True when FFN fc1 -> act -> fc2 (+ shortcut) can run as ONE kernel with the hidden tile on chip
    (grafp_ffn_fused_fwd): fp32 x of C in {64, 128} channels, f16x3 engine, both layers un-grouped with folded scale
    and shift.  GRAFP_NO_FFN_FUSED=1 keeps the two-GEMM route."""
    if os.environ.get("GRAFP_NO_FFN_FUSED", "0") == "1" or isinstance(x, SplitAct):
        return False
    if _effective_engine() not in (_lib.ENGINE_AUTO, _lib.ENGINE_TC_F16X3):
        return False
    if lin1.groups != 1 or lin2.groups != 1 or lin1.w_split_f16 is None or lin2.w_split_f16 is None:
        return False
    if lin1.scale is None or lin1.shift is None or lin2.scale is None or lin2.shift is None:
        return False
    hid, c = lin1.w.shape
    if lin2.w.shape != (c, hid) or not x.is_cuda or x.dtype != torch.float32 or x.dim() != 2 or x.shape[1] != c:
        return False
    return bool(_lib.load().grafp_ffn_fused_supported(x.shape[0], c, hid)) and x.shape[0] > 0


def ffn_fused(x: torch.Tensor, lin1, lin2, act=None, act_param: float = 0.0) -> torch.Tensor:
    """y = x + scale2 * (act(scale1 * (x W1^T) + shift1) W2^T) + shift2 in one kernel (include/grafp.h)."""
    x = _chk(x, name="x")
    M, c = x.shape
    hid = lin1.w.shape[0]
    y = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(_lib.load().grafp_ffn_fused_fwd(
            _ptr(x), x.stride(0), M, c, hid, _ptr(lin1.w_split_f16), lin1.w_split_f16.stride(0), float(lin1.f16_unscale),
            _ptr(lin1.scale), _ptr(lin1.shift), act_code(act), act_param, _ptr(lin2.w_split_f16),
            lin2.w_split_f16.stride(0), float(lin2.f16_unscale), _ptr(lin2.scale), _ptr(lin2.shift), _ptr(y),
            y.stride(0), _stream(x)), "ffn_fused_fwd")
    return y


def mrconv_fc2_fused_ok(lin_mr, lin_fc2, x, residual) -> bool:
    """True when MRConv2d's grouped conv -> fc2 (+ shortcut) can run as ONE kernel with the 2C-wide MRConv output on
    chip (grafp_mrconv_fc2_fused_fwd): fp32 x of C in {64, 128} channels, f16x3 engine.  GRAFP_NO_MR_FUSED=1 keeps
    the two-GEMM route."""
    if os.environ.get("GRAFP_NO_MR_FUSED", "0") == "1" or isinstance(x, SplitAct) or residual is None:
        return False
    if _effective_engine() not in (_lib.ENGINE_AUTO, _lib.ENGINE_TC_F16X3):
        return False
    if lin_mr.w_mr_chunked is None or lin_fc2.groups != 1 or lin_fc2.w_split_f16 is None:
        return False
    if lin_mr.scale is None or lin_mr.shift is None or lin_fc2.scale is None or lin_fc2.shift is None:
        return False
    if not x.is_cuda or x.dtype != torch.float32 or x.dim() != 2:
        return False
    c = x.shape[1]
    if lin_mr.w_mr_chunked.shape != (2, 2 * c, 64) or lin_fc2.w.shape != (c, 2 * c) or residual.shape != x.shape:
        return False
    return bool(_lib.load().grafp_mrconv_fc2_fused_supported(x.shape[0], c)) and x.shape[0] > 0


def mrconv_fc2_fused(x: torch.Tensor, m: torch.Tensor, lin_mr, act, act_param: float, lin_fc2,
                     residual: torch.Tensor) -> torch.Tensor:
    """y = residual + scale2 * (act(scale1 * ([x | m] Wmr^T) + shift1) Wfc2^T) + shift2 in one kernel
    (include/grafp.h: the tail of Grapher.forward, torch_vertex.py:24-34 + :183-195)."""
    x, m, residual = _chk(x, name="x"), _chk(m, name="m"), _chk(residual, name="residual")
    if m.shape != x.shape:
        raise GrafpError("mrconv_fc2_fused: x and m must have the same shape")
    M, c = x.shape
    y = torch.empty_like(x)
    w1, w2 = lin_mr.w_mr_chunked, lin_fc2.w_split_f16
    with torch.cuda.device(x.device):
        check(_lib.load().grafp_mrconv_fc2_fused_fwd(
            _ptr(x), x.stride(0), _ptr(m), m.stride(0), M, c, _ptr(w1), w1.stride(1), float(lin_mr.f16_unscale),
            _ptr(lin_mr.scale), _ptr(lin_mr.shift), act_code(act), act_param, _ptr(w2), w2.stride(0),
            float(lin_fc2.f16_unscale), _ptr(lin_fc2.scale), _ptr(lin_fc2.shift), _ptr(residual), residual.stride(0),
            _ptr(y), y.stride(0), _stream(x)), "mrconv_fc2_fused_fwd")
    return y


def stem_supported(cin: int, cout: int, N: int) -> bool:
    return cin in (4, 8, 16) and cout % 4 == 0 and 4 <= cout <= 1024 and 256 % (cout // 4) == 0 and \
        (cin * N + cin * cout) * 4 <= 96 * 1024


def stem(x: torch.Tensor, lin, act=None, act_param: float = 0.0, B: int = None, N: int = None) -> torch.Tensor:
    """Stem layer (tiny input width) over a prepared ``_prep.Linear``: x is the reference's (B, Cin, N)
    tensor, or node-major (B*N, Cin) when B and N are given.  -> node-major (B*N, Cout)."""
    x = _chk(x, name="x")
    nchw = x.dim() == 3
    if nchw:
        B, cin, N = x.shape
    else:
        cin = x.shape[1]
    cout = lin.w.shape[0]
    out = torch.empty((B * N, cout), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        check(_lib.load().grafp_stem_fwd(_ptr(x), _ptr(lin.w), _ptr(lin.scale), _ptr(lin.shift), B, cin, N, cout,
                                         int(nchw), act_code(act), act_param, _ptr(out), _stream(x)), "stem_fwd")
    return out


def nchw_to_nodes_add(x: torch.Tensor, pos: Optional[torch.Tensor]) -> torch.Tensor:
    """(B, C, N) -> (B*N, C) with pos (N, C) added to every graph's rows (pos may be None)."""
    x = _chk(x, name="x")
    B, Cc, N = x.shape
    if pos is not None:
        pos = _chk(pos, name="pos")
        if tuple(pos.shape) != (N, Cc):
            raise GrafpError("nchw_to_nodes_add: pos must be (N, C) = (%d, %d), got %s" % (N, Cc, tuple(pos.shape)))
    out = torch.empty((B * N, Cc), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        check(_lib.load().grafp_nchw_to_nodes_add(_ptr(x), _ptr(pos), _ptr(out), B, Cc, N, _stream(x)),
              "nchw_to_nodes_add")
    return out


def mha_pool(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, P: int, Nq: int, Nk: int, heads: int) -> torch.Tensor:
    """mean over the Nq queries of per-head softmax(q k^T / sqrt(Dh)) v:  q (P*Nq, E), k / v (P*Nk, E) row-major
    views with unit column stride (column slices of a fused projection are fine) -> (P, E)."""
    E = q.shape[1]
    for t, nm in ((q, "q"), (k, "k"), (v, "v")):
        if not t.is_cuda or t.dtype != torch.float32 or t.stride(1) != 1 or t.shape[1] != E:
            raise GrafpError("mha_pool: %s must be an fp32 CUDA (rows, E) view with unit column stride" % nm)
    Dh = E // heads
    out = torch.empty((P, E), device=q.device, dtype=torch.float32)
    with torch.cuda.device(q.device):
        check(_lib.load().grafp_mha_pool_fwd(_ptr(q), q.stride(0), _ptr(k), k.stride(0), _ptr(v), v.stride(0), P, Nq, Nk,
                                             heads, Dh, 1.0 / float(Dh) ** 0.5, _ptr(out), out.stride(0), _stream(q)),
              "mha_pool_fwd")
    return out


def node_mean(x: torch.Tensor, B: int, N: int) -> torch.Tensor:
    x = _chk(x, name="x")
    out = torch.empty((B, x.shape[1]), device=x.device, dtype=torch.float32)
    with torch.cuda.device(x.device):
        check(_lib.load().grafp_node_mean(_ptr(x), B, N, x.shape[1], _ptr(out), _stream(x)), "node_mean")
    return out


def peak_extract(spec: torch.Tensor, w: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """(B, n_mels, n_frames) -> node-major (B*N, F)."""
    spec = _chk(spec, name="spec")
    w = _chk(w, name="w")
    bias = _chk(bias, name="bias")
    B, n_mels, n_frames = spec.shape
    F, _, pb, pf = w.shape
    nodes = (n_mels // pb) * (n_frames // pf)
    out = torch.empty((B * nodes, F), device=spec.device, dtype=torch.float32)
    with torch.cuda.device(spec.device):
        check(_lib.load().grafp_peak_extract_fwd(_ptr(spec), _ptr(w), _ptr(bias), B, n_mels, n_frames,
                                                 F, pb, pf, _ptr(out), _stream(spec)), "peak_extract")
    return out


def l2_normalize_rows(z: torch.Tensor, eps: float) -> torch.Tensor:
    z = _chk(z, name="z")
    out = torch.empty_like(z)
    with torch.cuda.device(z.device):
        check(_lib.load().grafp_l2_normalize_rows(_ptr(z), z.shape[0], z.shape[1], eps, _ptr(out),
                                                  _stream(z)), "l2_normalize_rows")
    return out


def ntxent_fwd(z: torch.Tensor, tau: float, row0: int = 0, rows: Optional[int] = None):
    """z: (n, D) interleaved rows of the global batch -> (loss_part (1,), lse (rows,))."""
    z = _chk(z, name="z")
    n, D = z.shape
    rows = n if rows is None else rows
    lse = torch.empty((rows,), device=z.device, dtype=torch.float32)
    loss = torch.zeros((1,), device=z.device, dtype=torch.float32)
    with torch.cuda.device(z.device):
        check(_lib.load().grafp_ntxent_fwd(_ptr(z), n, D, tau, row0, rows, _ptr(lse), _ptr(loss),
                                           _stream(z)), "ntxent_fwd")
    return loss, lse


def ntxent_bwd(z: torch.Tensor, lse_all: torch.Tensor, tau: float, grad_loss: torch.Tensor,
               row0: int = 0, rows: Optional[int] = None) -> torch.Tensor:
    z = _chk(z, name="z")
    n, D = z.shape
    rows = n if rows is None else rows
    dz = torch.empty((rows, D), device=z.device, dtype=torch.float32)
    grad_loss = _chk(grad_loss.reshape(1), name="grad_loss")
    with torch.cuda.device(z.device):
        check(_lib.load().grafp_ntxent_bwd(_ptr(z), _ptr(lse_all), n, D, tau, row0, rows,
                                           _ptr(grad_loss), _ptr(dz), _stream(z)), "ntxent_bwd")
    return dz


# ------------------------------------------------------------------------------------------
# train-step kernels (csrc/train.cu)
# ------------------------------------------------------------------------------------------
def col_stats(x: torch.Tensor) -> torch.Tensor:
    """(M, C) -> fp64 (2, C): column sums and sums of squares."""
    x = _chk(x, name="x")
    M, Cc = x.shape
    out = torch.zeros((2, Cc), device=x.device, dtype=torch.float64)
    with torch.cuda.device(x.device):
        check(_lib.load().grafp_col_stats(_ptr(x), M, Cc, x.stride(0), _ptr(out[0]), _ptr(out[1]), _stream(x)),
              "col_stats")
    return out


def bn_finalize(stats: torch.Tensor, M: int, gamma, beta, conv_bias, eps: float, momentum: float,
                running_mean, running_var):
    """-> fp32 (4, C): scale, shift, mean, invstd; updates the running statistics in place."""
    Cc = stats.shape[1]
    out = torch.empty((4, Cc), device=stats.device, dtype=torch.float32)
    with torch.cuda.device(stats.device):
        check(_lib.load().grafp_bn_finalize(_ptr(stats[0]), _ptr(stats[1]), M, Cc, _ptr(gamma), _ptr(beta),
                                            _ptr(conv_bias), eps, momentum, _ptr(running_mean),
                                            _ptr(running_var), _ptr(out[0]), _ptr(out[1]), _ptr(out[2]),
                                            _ptr(out[3]), _stream(stats)), "bn_finalize")
    return out


def affine_act(x: torch.Tensor, scale, shift, act=None, act_param: float = 0.0, residual=None) -> torch.Tensor:
    x = _chk(x, name="x")
    M, Cc = x.shape
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        check(_lib.load().grafp_affine_act(_ptr(x), M, Cc, x.stride(0), _ptr(scale), _ptr(shift), act_code(act),
                                           act_param, _ptr(residual), residual.stride(0) if residual is not None else 0,
                                           _ptr(out), out.stride(0), _stream(x)), "affine_act")
    return out


def bn_finalize_apply(stats: torch.Tensor, raw: torch.Tensor, gamma, beta, conv_bias, eps: float, momentum: float,
                      running_mean, running_var, act=None, act_param: float = 0.0, residual=None):
    """bn_finalize + affine_act in one launch -> (out (M, C), ssmi fp32 (4, C) = scale, shift, mean, invstd)."""
    raw = _chk(raw, name="raw")
    M, Cc = raw.shape
    ssmi = torch.empty((4, Cc), device=raw.device, dtype=torch.float32)
    out = torch.empty_like(raw)
    with torch.cuda.device(raw.device):
        check(_lib.load().grafp_bn_finalize_apply(
            _ptr(stats[0]), _ptr(stats[1]), M, Cc, _ptr(gamma), _ptr(beta), _ptr(conv_bias), eps, momentum,
            _ptr(running_mean), _ptr(running_var), _ptr(ssmi[0]), _ptr(ssmi[1]), _ptr(ssmi[2]), _ptr(ssmi[3]),
            _ptr(raw), raw.stride(0), act_code(act), act_param, _ptr(residual),
            residual.stride(0) if residual is not None else 0, _ptr(out), out.stride(0), _stream(raw)), "bn_finalize_apply")
    return out, ssmi


def bn_act_bwd(dout: torch.Tensor, raw: torch.Tensor, ssmi: torch.Tensor, act, act_param: float, bn: bool,
               dgamma=None, dbeta=None):
    """Backward of (BatchNorm | bias) + activation.  ssmi: (4, C) scale/shift/mean/invstd.
    Returns (draw (M, C), sums fp64 (2, C) = [sum dz, sum dz*xhat]).  ``dgamma`` / ``dbeta``: (C,) fp32 views that
    accumulate the parameter gradients inside the apply kernel (saves the separate bn_param_grad launch)."""
    dout = _chk(dout, name="dout")
    raw = _chk(raw, name="raw")
    M, Cc = raw.shape
    sums = torch.zeros((2, Cc), device=raw.device, dtype=torch.float64)
    draw = torch.empty_like(raw)
    a = act_code(act)
    lib = _lib.load()
    with torch.cuda.device(raw.device):
        check(lib.grafp_bn_bwd_reduce(_ptr(dout), dout.stride(0), _ptr(raw), raw.stride(0), M, Cc, _ptr(ssmi[0]),
                                      _ptr(ssmi[1]), _ptr(ssmi[2]), _ptr(ssmi[3]), a, act_param, _ptr(sums[0]),
                                      _ptr(sums[1]), _stream(raw)), "bn_bwd_reduce")
        check(lib.grafp_bn_bwd_apply(_ptr(dout), dout.stride(0), _ptr(raw), raw.stride(0), M, Cc, _ptr(ssmi[0]),
                                     _ptr(ssmi[1]), _ptr(ssmi[2]), _ptr(ssmi[3]), a, act_param, int(bn),
                                     _ptr(sums[0]), _ptr(sums[1]), _ptr(draw), draw.stride(0), _ptr(dgamma), _ptr(dbeta),
                                     _stream(raw)), "bn_bwd_apply")
    return draw, sums


def bn_param_grad(sums: torch.Tensor, want_gamma: bool, want_beta: bool, out_gamma=None, out_beta=None):
    """dgamma += sum dz*xhat, dbeta += sum dz.  ``out_gamma`` / ``out_beta``: accumulate into these (C,) fp32 views
    (e.g. slices of the flat gradient buffer) instead of fresh zero tensors."""
    Cc = sums.shape[1]
    dg = (out_gamma if out_gamma is not None else torch.zeros((Cc,), device=sums.device, dtype=torch.float32)) \
        if want_gamma else None
    db = (out_beta if out_beta is not None else torch.zeros((Cc,), device=sums.device, dtype=torch.float32)) \
        if want_beta else None
    with torch.cuda.device(sums.device):
        check(_lib.load().grafp_bn_param_grad(_ptr(sums[0]), _ptr(sums[1]), Cc, _ptr(dg), _ptr(db), _stream(sums)),
              "bn_param_grad")
    return dg, db


def gemm_wgrad(dy: torch.Tensor, a1: torch.Tensor, a2, n_total: int, groups: int = 1,
               tap3_nodes: int = 0, engine: Optional[int] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dw (groups*n, k1+k2) = sum_m dy[m]^T A[m] (per group).  Default engine: the tcgen05 3xTF32 kernel where the
    shape allows (deterministic split reduction), else the fp32 SIMT kernel; ``engine`` forces ENGINE_SIMT /
    ENGINE_TC_3XTF32.  ``out``: accumulate into this (groups*n, k1+k2) fp32 view (row stride multiple of 4) instead
    of a fresh zero tensor -- e.g. a slice of the flat gradient buffer."""
    dy = _chk(dy, name="dy")
    a1 = _chk(a1, name="a1")
    n = n_total // groups
    if tap3_nodes > 0:
        k1, k2, M = 3 * a1.shape[1], 0, a1.shape[0] // 2
    else:
        k1, k2, M = a1.shape[1] // groups, 0, a1.shape[0]
        if a2 is not None:
            a2 = _chk(a2, name="a2")
            k2 = a2.shape[1] // groups
    if out is None:
        dw = torch.zeros((n_total, k1 + k2), device=dy.device, dtype=torch.float32)
    else:
        dw = out
        if dw.shape != (n_total, k1 + k2) or dw.stride(1) != 1 or dw.dtype != torch.float32 or not dw.is_cuda:
            raise GrafpError("gemm_wgrad: out must be an fp32 CUDA (groups*n, k1+k2) view with unit column stride")
    eng = _lib.ENGINE_AUTO if engine is None else engine
    if engine is None and _effective_engine() == _lib.ENGINE_SIMT:
        eng = _lib.ENGINE_SIMT
    lib = _lib.load()
    ws_bytes = int(lib.grafp_gemm_wgrad_workspace_bytes(M, n, k1, k2, groups, tap3_nodes)) if eng != _lib.ENGINE_SIMT else 0
    ws = torch.empty((ws_bytes // 4,), device=dy.device, dtype=torch.float32) if ws_bytes else None
    with torch.cuda.device(dy.device):
        check(lib.grafp_gemm_wgrad(_ptr(dy), dy.stride(0), _ptr(a1), a1.stride(0), k1, _ptr(a2),
                                   a2.stride(0) if a2 is not None else 0, k2, M, n, groups, tap3_nodes,
                                   _ptr(dw), dw.stride(0), eng, _ptr(ws), ws_bytes, _stream(dy)), "gemm_wgrad")
    return dw


def tap3_bwd_input(dA: torch.Tensor, rows_per_graph: int, cin: int) -> torch.Tensor:
    dA = _chk(dA, name="dA")
    rows = dA.shape[0]
    dX = torch.empty((2 * rows, cin), device=dA.device, dtype=torch.float32)
    with torch.cuda.device(dA.device):
        check(_lib.load().grafp_tap3_bwd_input(_ptr(dA), rows, rows_per_graph, cin, _ptr(dX), _stream(dA)),
              "tap3_bwd_input")
    return dX


def node_mean_bwd(dmean: torch.Tensor, B: int, N: int) -> torch.Tensor:
    dmean = _chk(dmean, name="dmean")
    Cc = dmean.shape[1]
    dx = torch.empty((B * N, Cc), device=dmean.device, dtype=torch.float32)
    with torch.cuda.device(dmean.device):
        check(_lib.load().grafp_node_mean_bwd(_ptr(dmean), B, N, Cc, _ptr(dx), _stream(dmean)), "node_mean_bwd")
    return dx


def l2_normalize_rows_bwd(v: torch.Tensor, dz: torch.Tensor, eps: float) -> torch.Tensor:
    v = _chk(v, name="v")
    dz = _chk(dz, name="dz")
    dv = torch.empty_like(v)
    with torch.cuda.device(v.device):
        check(_lib.load().grafp_l2_normalize_rows_bwd(_ptr(v), _ptr(dz), v.shape[0], v.shape[1], eps, _ptr(dv),
                                                      _stream(v)), "l2_normalize_rows_bwd")
    return dv


def peak_extract_bwd(spec: torch.Tensor, w: torch.Tensor, bias: torch.Tensor, dout: torch.Tensor):
    spec = _chk(spec, name="spec")
    dout = _chk(dout, name="dout")
    B, n_mels, n_frames = spec.shape
    F, _, pb, pf = w.shape
    dw = torch.zeros_like(w, dtype=torch.float32).contiguous()
    db = torch.zeros((F,), device=w.device, dtype=torch.float32)
    with torch.cuda.device(spec.device):
        check(_lib.load().grafp_peak_extract_bwd(_ptr(spec), _ptr(_chk(w, name="w")), _ptr(_chk(bias, name="bias")),
                                                 _ptr(dout), B, n_mels, n_frames, F, pb, pf, _ptr(dw), _ptr(db),
                                                 _stream(spec)), "peak_extract_bwd")
    return dw, db


def sq_norm(g: torch.Tensor, out: torch.Tensor) -> None:
    with torch.cuda.device(g.device):
        check(_lib.load().grafp_sq_norm(_ptr(g), g.numel(), _ptr(out), _stream(g)), "sq_norm")


def adam_clip_step(p, g, m, v, lr, beta1, beta2, eps, step, max_norm, sq) -> None:
    with torch.cuda.device(p.device):
        check(_lib.load().grafp_adam_clip_step(_ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel(), lr, beta1, beta2, eps,
                                               step, max_norm, _ptr(sq), _stream(p)), "adam_clip_step")


def adam_clip_step_dev(p, g, m, v, lr_dev, beta1, beta2, eps, step_dev, max_norm, sq, loss_guard) -> None:
    with torch.cuda.device(p.device):
        check(_lib.load().grafp_adam_clip_step_dev(_ptr(p), _ptr(g), _ptr(m), _ptr(v), p.numel(), _ptr(lr_dev),
                                                   beta1, beta2, eps, _ptr(step_dev), max_norm, _ptr(sq),
                                                   _ptr(loss_guard), _stream(p)), "adam_clip_step_dev")


def add_inplace(y: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    with torch.cuda.device(y.device):
        check(_lib.load().grafp_add_inplace(_ptr(y), _ptr(x), y.numel(), _stream(y)), "add_inplace")
    return y
