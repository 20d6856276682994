"""B200 counterparts of the reference's encoder/gcn_lib/torch_vertex.py: ``MRConv2d`` (:11-34),
``GraphConv2d`` (:92-111), ``DyGraphConv2d`` (:114-139) and ``Grapher`` (:142-195), with the
constructor / forward signatures and state_dict keys of the reference.  Every ``forward`` takes
the reference's NCHW tensors; ``forward_nodes`` is the node-major form GraphEncoder chains
without layout round trips."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from ... import ops
from ..._prep import fold_conv_bn, make_linear, sig
from .pos_embed import get_2d_relative_pos_embed
from .torch_edge import DenseDilatedKnnGraph
from .torch_nn import BasicConv


class MRConv2d(nn.Module):
    """Max-relative graph convolution: nn([x, max_j(x_j - x_i)] interleaved)."""

    def __init__(self, in_channels, out_channels, act="relu", norm=None, bias=True):
        super().__init__()
        self.nn = BasicConv([in_channels * 2, out_channels], act, norm, bias)

    def forward_nodes(self, x: torch.Tensor, nn_idx: torch.Tensor, B: int, N: int, out_split: bool = False):
        lin = self.nn.layer_params(0, interleaved_sources=True)[0]
        if ops.fused_mr_ok(lin, x.shape[1]):
            # gather + max-relative computed inside the GEMM: the (M, C) aggregate never reaches HBM
            return self.nn.forward_nodes(x, None, out_split, x2_gather=(nn_idx, N))
        m = ops.mr_aggregate(x, nn_idx, B, N)
        return self.nn.forward_nodes(x, m, out_split)

    def forward(self, x, edge_index, y=None):
        if y is not None:
            raise NotImplementedError("r > 1 (pooled y) graphs are not reached by GraphEncoder (r=1)")
        B, C, N = x.shape[:3]
        nodes = ops.nchw_to_nodes(x.reshape(B, C, N))
        out = self.forward_nodes(nodes, edge_index[0].to(torch.int32), B, N)
        return ops.nodes_to_nchw(out, B, N).unsqueeze(-1)


def _concat_halves(conv_seq, tag_cache):
    """Prepared operands of a BasicConv layer whose input is a PLAIN channel concat [a | b] (EdgeConv2d,
    GraphSAGE.nn2): with groups=4 the first two groups read only `a`, the last two only `b`, so the layer
    is two independent 2-group GEMMs writing the two column halves of the output.
    Returns ((lin_a, lin_b), act_name, slope)."""
    e = conv_seq._plan[0]
    conv = conv_seq[e["conv"]]
    bn = conv_seq[e["bn"]] if e["bn"] is not None else None
    key = sig(conv.weight, conv.bias) + (sig(bn.weight, bn.bias, bn.running_mean, bn.running_var) if bn is not None else ())
    hit = tag_cache.get("halves")
    if hit is None or hit[0] != key:
        w, scale, shift = fold_conv_bn(conv.weight, conv.bias, bn)
        h = w.shape[0] // 2
        sl = lambda t, a, b: t[a:b].contiguous() if t is not None else None
        hit = (key, (make_linear(w[:h].contiguous(), sl(scale, 0, h), sl(shift, 0, h), 2),
                     make_linear(w[h:].contiguous(), sl(scale, h, 2 * h), sl(shift, h, 2 * h), 2)))
        tag_cache["halves"] = hit
    act = conv_seq[e["act"]] if e["act"] is not None else None
    return hit[1], (act.name if act is not None else None), (act.neg_slope if act is not None else 0.0)


class _NodeGraphConv(nn.Module):
    """Shared NCHW wrapper of the node-major graph convolutions below."""

    def forward(self, x, edge_index, y=None):
        if y is not None:
            raise NotImplementedError("r > 1 (pooled y) graphs are not reached by GraphEncoder (r=1)")
        if self.training:
            raise RuntimeError("%s has an eval-mode sm_100a path only (the train path covers conv='mr')"
                               % type(self).__name__)
        B, C, N = x.shape[:3]
        nodes = ops.nchw_to_nodes(x.reshape(B, C, N))
        out = self.forward_nodes(nodes, edge_index[0].to(torch.int32), B, N)
        return ops.nodes_to_nchw(out, B, N).unsqueeze(-1)


class EdgeConv2d(_NodeGraphConv):
    """Edge convolution max_k nn(cat[x_i, x_j - x_i]) (reference torch_vertex.py:37-52), evaluated per node:
    the grouped 1x1 conv commutes with the gather, so groups 0-1 (which only see x_i) are one GEMM on the
    nodes and groups 2-3 are max_k act(scale * (P_j - P_i) + shift) over P = W x -- k times less GEMM work
    than the reference's (B, 2C, N, k) edge tensor, exact for any activation (it is applied per edge)."""

    def __init__(self, in_channels, out_channels, act="relu", norm=None, bias=True):
        super().__init__()
        self.nn = BasicConv([in_channels * 2, out_channels], act, norm, bias)
        self._cache = {}

    def forward_nodes(self, x: torch.Tensor, nn_idx: torch.Tensor, B: int, N: int, out_split: bool = False):
        (lin_a, lin_b), act, slope = _concat_halves(self.nn, self._cache)
        h = lin_a.w.shape[0]
        out = torch.empty((x.shape[0], 2 * h), device=x.device, dtype=torch.float32)
        ops.linear(x, lin_a, act, slope, out=out[:, :h])                       # groups 0-1: functions of x_i only
        raw = _without_epilogue(lin_b)
        p = ops.linear(x, raw)                                                 # P = W x (no bias / BN / act)
        ops.nbr_reduce(p, nn_idx, B, N, ops.NBR_EDGE_MAX, lin_b.scale, lin_b.shift, act, slope, out=out[:, h:])
        return out


class GraphSAGE(_NodeGraphConv):
    """nn2(cat[x, max_k nn1(x_j)]) (reference torch_vertex.py:55-68): nn1 is applied once per node, then
    gathered and max-reduced."""

    def __init__(self, in_channels, out_channels, act="relu", norm=None, bias=True):
        super().__init__()
        self.nn1 = BasicConv([in_channels, in_channels], act, norm, bias)
        self.nn2 = BasicConv([in_channels * 2, out_channels], act, norm, bias)
        self._cache = {}

    def forward_nodes(self, x: torch.Tensor, nn_idx: torch.Tensor, B: int, N: int, out_split: bool = False):
        q = self.nn1.forward_nodes(x)
        xj = ops.nbr_reduce(q, nn_idx, B, N, ops.NBR_MAX)
        (lin_a, lin_b), act, slope = _concat_halves(self.nn2, self._cache)
        h = lin_a.w.shape[0]
        out = torch.empty((x.shape[0], 2 * h), device=x.device, dtype=torch.float32)
        ops.linear(x, lin_a, act, slope, out=out[:, :h])
        ops.linear(xj, lin_b, act, slope, out=out[:, h:])
        return out


class GINConv2d(_NodeGraphConv):
    """nn((1 + eps) x + sum_k x_j) (reference torch_vertex.py:71-88)."""

    def __init__(self, in_channels, out_channels, act="relu", norm=None, bias=True):
        super().__init__()
        self.nn = BasicConv([in_channels, out_channels], act, norm, bias)
        self.eps = nn.Parameter(torch.Tensor([0.0]))

    def forward_nodes(self, x: torch.Tensor, nn_idx: torch.Tensor, B: int, N: int, out_split: bool = False):
        h = ops.nbr_reduce(x, nn_idx, B, N, ops.NBR_SUM_SELF, eps=self.eps.detach())
        return self.nn.forward_nodes(h)


def _without_epilogue(lin):
    """The same prepared weights without the epilogue (scale / shift dropped)."""
    from ..._prep import Linear
    return Linear(lin.w, None, None, lin.groups, lin.w_split, lin.w_split_bf16, lin.w_split_f16, lin.f16_unscale)


class GraphConv2d(nn.Module):
    """Static graph convolution dispatch (reference torch_vertex.py:92-111).  GraphEncoder hard-codes
    conv='mr' (SURVEY Q2); edge / sage / gin are the reference's other variants, eval-mode only here."""

    def __init__(self, in_channels, out_channels, conv="edge", act="relu", norm=None, bias=True):
        super().__init__()
        if conv == "edge":
            self.gconv = EdgeConv2d(in_channels, out_channels, act, norm, bias)
        elif conv == "mr":
            self.gconv = MRConv2d(in_channels, out_channels, act, norm, bias)
        elif conv == "sage":
            self.gconv = GraphSAGE(in_channels, out_channels, act, norm, bias)
        elif conv == "gin":
            self.gconv = GINConv2d(in_channels, out_channels, act, norm, bias)
        else:
            raise NotImplementedError("conv:{} is not supported".format(conv))

    def forward(self, x, edge_index, y=None):
        return self.gconv(x, edge_index, y)


class DyGraphConv2d(GraphConv2d):
    """Dynamic graph convolution: builds the dilated kNN graph of its input on every call."""

    def __init__(self, in_channels, out_channels, kernel_size=9, dilation=1, conv="edge", act="relu",
                 norm=None, bias=True, stochastic=False, epsilon=0.0, r=1):
        super().__init__(in_channels, out_channels, conv, act, norm, bias)
        if r != 1:
            raise NotImplementedError("r > 1 is never used by GraphEncoder (graph_encoder.py:171)")
        self.k, self.d, self.r = kernel_size, dilation, r
        self.dilated_knn_graph = DenseDilatedKnnGraph(kernel_size, dilation, stochastic, epsilon)

    def forward_nodes(self, x: torch.Tensor, B: int, N: int, nn_idx: torch.Tensor = None,
                      out_split: bool = False):
        if nn_idx is None:
            nn_idx = self.dilated_knn_graph.knn_nodes(x, B, N)
        return self.gconv.forward_nodes(x, nn_idx, B, N, out_split)

    def forward(self, x, relative_pos=None):
        if relative_pos is not None:
            raise NotImplementedError("relative_pos is never passed on the GraphEncoder path")
        B, C, H, W = x.shape
        nodes = ops.nchw_to_nodes(x.reshape(B, C, H * W))
        out = self.forward_nodes(nodes, B, H * W)
        return ops.nodes_to_nchw(out, B, H * W).reshape(B, -1, H, W)


class Grapher(nn.Module):
    """x + BN(fc2(graph_conv(BN(fc1(x)))))."""

    def __init__(self, in_channels, kernel_size=9, dilation=1, conv="edge", act="relu", norm=None,
                 bias=True, stochastic=False, epsilon=0.0, r=1, n=196, drop_path=0.0, relative_pos=False):
        super().__init__()
        self.channels, self.n, self.r = in_channels, n, r
        self.fc1 = nn.Sequential(nn.Conv2d(in_channels, in_channels, 1, stride=1, padding=0),
                                 nn.BatchNorm2d(in_channels))
        self.graph_conv = DyGraphConv2d(in_channels, in_channels * 2, kernel_size, dilation, conv, act,
                                        norm, bias, stochastic, epsilon, r)
        self.fc2 = nn.Sequential(nn.Conv2d(in_channels * 2, in_channels, 1, stride=1, padding=0),
                                 nn.BatchNorm2d(in_channels))
        if drop_path > 0.0:
            raise NotImplementedError("DropPath is never instantiated by GraphEncoder (SURVEY Q1)")
        self.drop_path = nn.Identity()
        self.relative_pos = None
        if relative_pos:
            # stored for state_dict compatibility only; never read in forward
            table = torch.from_numpy(np.float32(get_2d_relative_pos_embed(in_channels, int(n ** 0.5))))
            table = F.interpolate(table[None, None], size=(n, n // (r * r)), mode="bicubic",
                                  align_corners=False)
            self.relative_pos = nn.Parameter(-table.squeeze(1), requires_grad=False)
        self._cache = {}

    def _folded(self, name: str):
        seq = getattr(self, name)
        conv, bn = seq[0], seq[1]
        key = sig(conv.weight, conv.bias, bn.weight, bn.bias, bn.running_mean, bn.running_var)
        hit = self._cache.get(name)
        if hit is None or hit[0] != key:
            hit = (key, make_linear(*fold_conv_bn(conv.weight, conv.bias, bn)))
            self._cache[name] = hit
        return hit[1]

    def forward_nodes(self, x: torch.Tensor, B: int, N: int, nn_idx: torch.Tensor = None,
                      taps: dict = None, want_split: bool = False):
        """``want_split``: also return the output as an ops.SplitAct (written by the same fc2 epilogue) for a
        consumer GEMM -- the following FFN's fc1 -- when fc2 runs on a bf16 tensor-core engine: returns
        (out, SplitAct or None)."""
        if self.training:
            raise RuntimeError("Grapher.forward_nodes is the eval path; training goes through "
                               "neuralsampleid_b200.autograd")
        # the fc1 epilogue also accumulates sum_c y^2 per node, which the kNN needs for F.normalize
        rs = torch.zeros((x.shape[0],), device=x.device, dtype=torch.float32) if nn_idx is None else None
        y = ops.linear(x, self._folded("fc1"), row_sumsq=rs)
        if nn_idx is None:
            nn_idx = self.graph_conv.dilated_knn_graph.knn_nodes(y, B, N, rs)
        if taps is not None:
            taps["fc1"] = y
            taps["idx"] = nn_idx
        fc2 = self._folded("fc2")
        gconv = self.graph_conv.gconv
        if isinstance(gconv, MRConv2d) and not want_split:
            lin, act, slope = gconv.nn.layer_params(0, interleaved_sources=True)
            if len(gconv.nn._plan) == 1 and ops.mrconv_fc2_fused_ok(lin, fc2, y, x):
                # C <= 128: MRConv's grouped conv and fc2 are HBM-bound on the 2C-wide tensor between them -- one
                # kernel keeps it on chip (bit-identical to the two GEMMs)
                m = ops.mr_aggregate(y, nn_idx, B, N)
                return ops.mrconv_fc2_fused(y, m, lin, act, slope, fc2, x)
        # the MRConv output feeds only fc2: split-bf16 on the bf16 tensor-core engines (ops.SplitAct)
        g = self.graph_conv.forward_nodes(y, B, N, nn_idx, out_split=ops.split_ok(fc2, 2 * self.channels))
        if want_split:
            if isinstance(g, ops.SplitAct) and ops.split_ok(fc2, 2 * self.channels):
                return ops.linear(g, fc2, residual=x, out_split="both")
            return ops.linear(g, fc2, residual=x), None
        return ops.linear(g, fc2, residual=x)

    def forward(self, x):
        B, C, N = x.shape[:3]
        nodes = ops.nchw_to_nodes(x.reshape(B, C, -1))
        N = nodes.shape[0] // B
        out = self.forward_nodes(nodes, B, N)
        return ops.nodes_to_nchw(out, B, N).reshape(x.shape)
