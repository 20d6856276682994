"""B200 counterparts of the reference's encoder/gcn_lib/torch_nn.py: ``BasicConv`` (grouped 1x1
conv + BN + activation, :52-76), ``batched_index_select`` (:79-98), ``act_layer`` / ``norm_layer``
factories (:9-37).  torch ``nn.Conv2d`` / ``nn.BatchNorm2d`` objects are kept purely as parameter
containers (identical state_dict keys); the arithmetic runs in libgrafp_sm100a kernels."""
from __future__ import annotations

import torch
from torch import nn

from ... import ops
from ..._prep import fold_conv_bn, make_linear, sig

_SUPPORTED_ACTS = ("relu", "leakyrelu", "gelu")


class _Act(nn.Module):
    """Marker module (no parameters): the activation is fused into the producing GEMM epilogue."""

    def __init__(self, name: str, neg_slope: float = 0.2):
        super().__init__()
        self.name = name
        self.neg_slope = neg_slope

    def forward(self, x):   # only reached if a user calls the marker directly
        raise RuntimeError("activation '%s' is fused into the GEMM epilogue; call the owning module"
                           % self.name)


def act_layer(act, inplace=False, neg_slope=0.2, n_prelu=1):
    act = act.lower()
    if act not in _SUPPORTED_ACTS:
        if act in ("prelu", "hswish"):
            raise NotImplementedError("activation layer [%s] has no sm_100a epilogue (the GraphEncoder "
                                      "path uses relu/leakyrelu/gelu)" % act)
        raise NotImplementedError("activation layer [%s] is not found" % act)
    return _Act(act, neg_slope)


def norm_layer(norm, nc):
    norm = norm.lower()
    if norm == "batch":
        return nn.BatchNorm2d(nc, affine=True)
    if norm == "instance":
        raise NotImplementedError("normalization layer [instance] has no sm_100a kernel (the "
                                  "GraphEncoder path uses batch)")
    raise NotImplementedError("normalization layer [%s] is not found" % norm)


class BasicConv(nn.Sequential):
    """Conv2d(1x1, groups=4) [+ BatchNorm2d] [+ activation] per channel pair, as in the reference
    (same child indices, so ``nn.0.weight`` / ``nn.1.running_mean`` ... keys are unchanged)."""

    GROUPS = 4

    def __init__(self, channels, act="relu", norm=None, bias=True, drop=0.0):
        if drop > 0:
            raise NotImplementedError("BasicConv dropout is not used on the GraphEncoder path")
        layers = []
        self._plan = []
        for i in range(1, len(channels)):
            conv = nn.Conv2d(channels[i - 1], channels[i], 1, bias=bias, groups=self.GROUPS)
            entry = {"conv": len(layers), "bn": None, "act": None}
            layers.append(conv)
            if norm is not None and norm.lower() != "none":
                entry["bn"] = len(layers)
                layers.append(norm_layer(norm, channels[-1]))
            if act is not None and act.lower() != "none":
                entry["act"] = len(layers)
                layers.append(act_layer(act))
            self._plan.append(entry)
        super().__init__(*layers)
        self._cache = {}
        self.reset_parameters()

    def reset_parameters(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
            elif isinstance(m, nn.BatchNorm2d):
                m.weight.data.fill_(1)
                m.bias.data.zero_()

    # -- node-major implementation ------------------------------------------------------
    def layer_params(self, i: int, interleaved_sources: bool):
        """Prepared (Linear, act, slope) of layer i (eval mode).  With
        ``interleaved_sources`` the columns of each group are regrouped [even | odd] so the layer
        reads two separate node matrices instead of the reference's channel-interleaved cat."""
        e = self._plan[i]
        conv = self[e["conv"]]
        bn = self[e["bn"]] if e["bn"] is not None else None
        key = (i, interleaved_sources) + sig(conv.weight, conv.bias) + \
            (sig(bn.weight, bn.bias, bn.running_mean, bn.running_var) if bn is not None else ())
        hit = self._cache.get((i, interleaved_sources))
        if hit is None or hit[0] != key:
            w, scale, shift = fold_conv_bn(conv.weight, conv.bias, bn)
            if interleaved_sources:
                w = torch.cat([w[:, 0::2], w[:, 1::2]], dim=1).contiguous()
            hit = (key, make_linear(w, scale, shift, self.GROUPS, dual=interleaved_sources))
            self._cache[(i, interleaved_sources)] = hit
        act = self[e["act"]] if e["act"] is not None else None
        return (hit[1],) + ((act.name, act.neg_slope) if act is not None else (None, 0.0))

    def forward_nodes(self, x: torch.Tensor, x2: torch.Tensor = None, out_split: bool = False, x2_gather=None):
        """x: (M, C) node-major.  If ``x2`` is given, the first layer consumes the virtual
        interleave [x0, x2_0, x1, x2_1, ...] (MRConv2d) without materialising it; ``x2_gather`` = (idx, N)
        instead has the kernel compute x2 = max-relative(x) on the fly (fused MRConv2d).  ``out_split``:
        the caller's consumer is a bf16 tensor-core GEMM, return an ops.SplitAct if this layer can
        produce one."""
        if self.training:
            raise RuntimeError("BasicConv.forward_nodes is the eval path; training goes through "
                               "neuralsampleid_b200.autograd")
        last = len(self._plan) - 1
        for i in range(len(self._plan)):
            dual = i == 0 and (x2 is not None or x2_gather is not None)
            lin, act, slope = self.layer_params(i, interleaved_sources=dual)
            k_total = lin.w.shape[1] * lin.groups
            split = out_split and i == last and not isinstance(x, ops.SplitAct) and ops.split_ok(lin, k_total)
            x = ops.linear(x, lin, act, slope, a2=x2 if i == 0 else None, out_split=split,
                           a2_gather=x2_gather if i == 0 else None)
        return x

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        B, C, N = x.shape[0], x.shape[1], x.shape[2] * (x.shape[3] if x.dim() == 4 else 1)
        shape = x.shape
        y = self.forward_nodes(ops.nchw_to_nodes(x.reshape(B, C, N)))
        return ops.nodes_to_nchw(y, B, N).reshape(B, -1, *shape[2:])


def batched_index_select(x: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    """x (B, C, N, 1), idx (B, N, k) integer -> (B, C, N, k) neighbour features (reference
    torch_nn.py:79-98)."""
    B, C, N = x.shape[:3]
    nodes = ops.nchw_to_nodes(x.reshape(B, C, N))
    return ops.index_select(nodes, idx.to(torch.int32), B, N)
