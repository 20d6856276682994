"""Host-side parameter preparation (done once per weight version, never on the hot path):
folding Conv bias + eval-mode BatchNorm into a per-channel (scale, shift) epilogue and reshaping
conv weights into the (n, k) row-major matrices the GEMM kernels read."""
from __future__ import annotations

from typing import Optional, Tuple

import torch


def sig(*tensors) -> tuple:
    """Cheap identity+version signature of parameters, used to invalidate prepared weights after
    optimizer steps / load_state_dict / .to()."""
    return tuple((t.data_ptr(), t._version, t.device.index) for t in tensors if t is not None)


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
