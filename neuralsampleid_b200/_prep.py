"""Host-side parameter preparation (done once per weight version, never on the hot path):
folding Conv bias + eval-mode BatchNorm into a per-channel (scale, shift) epilogue and reshaping
conv weights into the (n, k) row-major matrices the GEMM kernels read."""
from __future__ import annotations

from typing import Optional, Tuple

import torch


_epoch = 0


def bump_epoch() -> None:
    """Called by code that updates parameters through raw pointers (the fused optimizer kernels, CUDA
    graph replays): torch's version counters do not see those writes."""
    global _epoch
    _epoch += 1


_wepoch = 0


def bump_weights() -> None:
    """Parameters themselves changed through raw pointers (fused optimizer step, CUDA-graph replay of a train step):
    invalidates the eval path's prepared weights AND the train path's per-step operand cache."""
    global _wepoch
    _wepoch += 1
    bump_epoch()


def weights_epoch() -> int:
    return _wepoch


def sig(*tensors) -> tuple:
    """Cheap identity+version signature of parameters, used to invalidate prepared weights after
    optimizer steps / load_state_dict / .to()."""
    return (_epoch,) + tuple((t.data_ptr(), t._version, t.device.index) for t in tensors if t is not None)


@torch.no_grad()
def fold_conv_bn(weight: torch.Tensor, bias: Optional[torch.Tensor], bn) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    """(Cout, Cin/g, 1, 1) conv [+ bias] followed by eval-mode BatchNorm2d ->
    W (Cout, Cin/g) and per-channel scale/shift with  y = scale * (W x) + shift."""
    w = weight.detach().reshape(weight.shape[0], -1).float().contiguous()
    cout = w.shape[0]
    if bn is None:
        scale = torch.ones(cout, device=w.device, dtype=torch.float32)
        shift = bias.detach().float().clone() if bias is not None else torch.zeros_like(scale)
        return w, scale, shift
    inv = bn.weight.detach().double() / torch.sqrt(bn.running_var.detach().double() + bn.eps)
    b = bias.detach().double() if bias is not None else 0.0
    shift = bn.bias.detach().double() + inv * (b - bn.running_mean.detach().double())
    return w, inv.float().contiguous(), shift.float().contiguous()


@torch.no_grad()
def tap3_weight(weight: torch.Tensor) -> torch.Tensor:
    """(Cout, Cin, 3, 3) stride-2 conv on an (N, 1) image only ever sees its centre column
    (SURVEY Q8): returns (Cout, 3*Cin) with column t*Cin + ci = weight[co, ci, t, 1]."""
    return weight.detach()[:, :, :, 1].permute(0, 2, 1).reshape(weight.shape[0], -1).float().contiguous()


class Linear:
    """Prepared operands of one fused GEMM layer: weight (groups*n, k), optional stacked tf32
    [hi ; lo] split for the tcgen05 3xTF32 engine, per-channel scale / shift."""
    __slots__ = ("w", "w_split", "w_split_bf16", "w_split_f16", "f16_unscale", "scale", "shift", "groups",
                 "w_mr_chunked")

    def __init__(self, w, scale, shift, groups=1, w_split=None, w_split_bf16=None, w_split_f16=None,
                 f16_unscale=0.0):
        self.w, self.scale, self.shift, self.groups = w, scale, shift, groups
        self.w_split, self.w_split_bf16 = w_split, w_split_bf16
        self.w_split_f16, self.f16_unscale = w_split_f16, f16_unscale
        # MRConv2d's grouped dual-source weight in the chunk-local form grafp_mrconv_fc2_fused_fwd reads
        # (include/grafp.h): fp16 planes (2, 2C, 64), or None
        self.w_mr_chunked = None


@torch.no_grad()
def make_linear(w: torch.Tensor, scale, shift, groups: int = 1, dual: bool = False) -> Linear:
    """Builds the device operands for ops.gemm.  ``dual``: the k axis of each group is fed from two
    sources of k/2 columns each.  A grouped layer whose per-source k extent is not a multiple of
    32 (the tcgen05 k-block) is densified into one block-diagonal group so it can still run on the
    tensor cores (at most the 2C x 2C first-stage MRConv: +1% of the encoder's flops)."""
    from . import ops
    w = w.float().contiguous()
    n_total, k = w.shape
    parts = 2 if dual else 1
    orig_groups, orig_shape = groups, (n_total, k * groups // parts)      # (2C, C) for MRConv2d's layer
    if groups > 1 and (k // parts) % 32 != 0 and ((k // parts) * groups) % 32 == 0:
        n = n_total // groups
        kp = k // parts
        dense = torch.zeros((n_total, k * groups), device=w.device, dtype=torch.float32)
        for g in range(groups):
            for q in range(parts):
                dense[g * n:(g + 1) * n, q * kp * groups + g * kp: q * kp * groups + (g + 1) * kp] = \
                    w[g * n:(g + 1) * n, q * kp:(q + 1) * kp]
        w, groups = dense, 1
        n_total, k = w.shape
    lin = Linear(w, scale, shift, groups)
    if w.is_cuda and (k // parts) % 32 == 0 and (n_total // groups) % 32 == 0:
        # both operand splits are tiny (weights): keep them so any engine can be selected later
        lin.w_split = ops.split_tf32(w)
        lin.w_split_bf16 = ops.split_bf16(w)
        pre = ops.f16_prescale(w)
        lin.w_split_f16, lin.f16_unscale = ops.split_f16(w, pre), 1.0 / pre
        if dual and orig_groups == 4 and orig_shape[0] == 2 * orig_shape[1]:
            # the (2C, 2C) groups = 4 layer of MRConv2d: chunk j = rows [64 j, 64 j + 64) only reads columns
            # [32 j, 32 j + 32) of either source (one group at C = 128, two block-diagonal groups at C = 64)
            c = orig_shape[1]                       # C = orig k per group (both sources) * 4 / 2 = n_total / 2
            sp = lin.w_split_f16.reshape(2, n_total, -1)
            if groups == 4 and c == 128:
                lin.w_mr_chunked = sp
            elif groups == 1 and c == 64:
                ch = torch.empty((2, n_total, 64), device=w.device, dtype=sp.dtype)
                for j in range(n_total // 64):
                    ch[:, 64 * j:64 * j + 64, :32] = sp[:, 64 * j:64 * j + 64, 32 * j:32 * j + 32]
                    ch[:, 64 * j:64 * j + 64, 32:] = sp[:, 64 * j:64 * j + 64, c + 32 * j:c + 32 * j + 32]
                lin.w_mr_chunked = ch.contiguous()
    return lin
