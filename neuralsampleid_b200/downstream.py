"""B200 counterpart of the reference's re-ranker ``CrossAttentionClassifier`` (downstream.py:30-79; SURVEY
section 8f rank 2): same constructor, ``forward(x_i, x_j)`` signature and state_dict keys
(``positional_embedding``, ``attn.in_proj_weight/bias``, ``attn.out_proj.weight/bias``, ``fc.0.*``, ``fc.3.*``),
eval mode.  It consumes the ``(B, C, N)`` node matrices ``GraphEncoder.forward(x, return_pre_proj=True)``
returns.  The reference calls it once per candidate from a Python loop (eval_hr.py:125-135); here a batch of
pairs is five kernels:

    transpose + positional add (x2)  ->  Q projection, fused K|V projection (tcgen05 GEMMs)
    ->  grafp_mha_pool_fwd: per-(pair, head) softmax attention + mean over the query nodes, taken BEFORE the
        output projection (a linear map commutes with the mean)  ->  out_proj, fc.0 + ReLU, fc.3 + sigmoid
"""
from __future__ import annotations

import torch
from torch import nn

from . import ops
from ._prep import make_linear, sig


class CrossAttentionClassifier(nn.Module):
    def __init__(self, in_dim, num_heads=4, hidden_dim=128, num_nodes=100, pos_embed=True):
        super().__init__()
        self.pos_embed = pos_embed
        if self.pos_embed:
            self.register_buffer("positional_embedding", torch.randn(1, num_nodes, in_dim))
        # parameter container only: its forward is never called
        self.attn = nn.MultiheadAttention(embed_dim=in_dim, num_heads=num_heads, batch_first=True)
        self.fc = nn.Sequential(nn.Linear(in_dim, hidden_dim), nn.ReLU(), nn.Dropout(p=0.3),
                                nn.Linear(hidden_dim, 1), nn.Sigmoid())
        self.num_heads = num_heads
        self._cache = None

    def _prepared(self):
        a = self.attn
        key = sig(a.in_proj_weight, a.in_proj_bias, a.out_proj.weight, a.out_proj.bias, self.fc[0].weight,
                  self.fc[0].bias, self.fc[3].weight, self.fc[3].bias)
        if self._cache is None or self._cache[0] != key:
            E = a.embed_dim
            w, b = a.in_proj_weight.detach().float(), a.in_proj_bias.detach().float()
            lin = lambda wt, bs: make_linear(wt.contiguous(), None, bs.contiguous())
            self._cache = (key, dict(
                q=lin(w[:E], b[:E]), kv=lin(w[E:], b[E:]),                        # K and V: one GEMM, n = 2E
                o=lin(a.out_proj.weight.detach().float(), a.out_proj.bias.detach().float()),
                f0=lin(self.fc[0].weight.detach().float(), self.fc[0].bias.detach().float()),
                f3=lin(self.fc[3].weight.detach().float(), self.fc[3].bias.detach().float())))
        return self._cache[1]

    def forward(self, x_i, x_j):
        """x_i, x_j: (B, C, N) node matrices -> (B, 1) match probability."""
        if self.training:
            raise RuntimeError("CrossAttentionClassifier has an eval-mode sm_100a path only (Dropout / autograd "
                               "are not implemented); call .eval()")
        if x_i.dim() != 3 or x_i.shape != x_j.shape:
            raise ValueError("expected two (B, C, N) tensors of the same shape")
        B, C, N = x_i.shape
        w = self._prepared()
        pos = self.positional_embedding[0, :N, :].contiguous() if self.pos_embed else None
        qi = ops.nchw_to_nodes_add(x_i, pos)                      # (B*N, C) = x_i^T + pos
        kj = ops.nchw_to_nodes_add(x_j, pos)
        q = ops.linear(qi, w["q"])                                # (B*N, E)
        kv = ops.linear(kj, w["kv"])                              # (B*N, 2E): [K | V]
        E = q.shape[1]
        pooled = ops.mha_pool(q, kv[:, :E], kv[:, E:], B, N, N, self.num_heads)   # (B, E), mean over queries
        a = ops.linear(pooled, w["o"])
        h = ops.linear(a, w["f0"], "relu")
        return ops.linear(h, w["f3"], "sigmoid")
