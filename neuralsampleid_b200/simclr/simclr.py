"""B200 counterpart of the reference's simclr/simclr.py (``SimCLR``, :7-47): peak extractor ->
encoder -> Linear/ELU/Linear projector -> L2 normalise, for both views.  Same constructor,
forward signature (returns h_i, h_j, z_i, z_j) and state_dict keys."""
from __future__ import annotations

import torch
from torch import nn

from .. import ops
from .._prep import make_linear, sig
from ..encoder.graph_encoder import GraphEncoder
from ..peak_extractor import GPUPeakExtractorv2


class SimCLR(nn.Module):
    def __init__(self, cfg, encoder):
        super().__init__()
        self.encoder = encoder
        self.cfg = cfg
        d, h, u = cfg["d"], cfg["h"], cfg["u"]
        if cfg["arch"] != "grafp":
            raise NotImplementedError("arch '%s' is outside the accelerated path (grafp only)" % cfg["arch"])
        self.peak_extractor = GPUPeakExtractorv2(cfg)
        self.projector = nn.Sequential(nn.Linear(h, d * u), nn.ELU(), nn.Linear(d * u, d))
        self._cache = {}

    def _proj_weights(self):
        l1, l2 = self.projector[0], self.projector[2]
        key = sig(l1.weight, l1.bias, l2.weight, l2.bias)
        hit = self._cache.get("p")
        if hit is None or hit[0] != key:
            hit = (key, (make_linear(l1.weight.detach(), None, l1.bias.detach().float().contiguous()),
                         make_linear(l2.weight.detach(), None, l2.bias.detach().float().contiguous())))
            self._cache["p"] = hit
        return hit[1]

    def _project(self, h: torch.Tensor) -> torch.Tensor:
        l1, l2 = self._proj_weights()
        z = ops.linear(ops.linear(h, l1, "elu"), l2)
        return ops.l2_normalize_rows(z, 1e-10)

    def _one_view(self, x, forced_idx=None):
        if isinstance(self.encoder, GraphEncoder) and not (self.training and torch.is_grad_enabled()):
            if self.encoder.training:
                # train mode under no_grad: batch-statistics forward (as the reference does), tape discarded
                from ..autograd import view_fwd
                h, z, _ = view_fwd(self, x, forced_idx)
                return h, z
            nodes, N = self.peak_extractor.forward_nodes(x)
            taps = [] if getattr(self, "_taps", None) is not None else None      # parity-test hook
            h = self.encoder.forward_nodes(nodes, x.shape[0], N, forced_idx=forced_idx, taps=taps)
            if taps is not None:
                self._taps.append(taps)
            return h, self._project(h)
        from ..autograd import simclr_view_train
        return simclr_view_train(self, x, forced_idx)

    def forward(self, x_i, x_j):
        forced = getattr(self, "_forced_idx", None)          # parity-test hook: (view i lists, view j lists)
        h_i, z_i = self._one_view(x_i, forced[0] if forced else None)
        h_j, z_j = self._one_view(x_j, forced[1] if forced else None)
        return h_i, h_j, z_i, z_j
