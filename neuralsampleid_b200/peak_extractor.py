"""B200 counterpart of the reference's peak_extractor.py (``GPUPeakExtractorv2``, :6-69): per-segment
min-max normalisation, (time ramp, frequency ramp, spectrogram) stack and the patch convolution +
ReLU, in one kernel that emits node features directly (no pre-built ramp tensors, so any batch size
takes the same path -- SURVEY Q13)."""
from __future__ import annotations

import torch
from torch import nn

from . import ops


class GPUPeakExtractorv2(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.n_filters = cfg["n_filters"]
        self.patch_bins = cfg["patch_bins"]
        self.patch_frames = cfg["patch_frames"]
        self.convs = nn.Sequential(
            nn.Conv2d(3, self.n_filters, kernel_size=(self.patch_bins, self.patch_frames),
                      stride=(self.patch_bins, self.patch_frames)),
            nn.ReLU())
        self.init_weights()

    def init_weights(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode="fan_out", nonlinearity="relu")
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    def forward_nodes(self, spec: torch.Tensor):
        """(B, n_mels, n_frames) -> node-major (B*N, n_filters), N."""
        B, n_mels, n_frames = spec.shape
        conv = self.convs[0]
        if self.training and torch.is_grad_enabled():
            from .autograd import PeakExtractFn
            out = PeakExtractFn.apply(spec, conv.weight, conv.bias)
        else:
            out = ops.peak_extract(spec, conv.weight.detach(), conv.bias.detach())
        return out, (n_mels // self.patch_bins) * (n_frames // self.patch_frames)

    def forward(self, spec_tensor):
        nodes, N = self.forward_nodes(spec_tensor)
        B = spec_tensor.shape[0]
        if nodes.requires_grad:
            return nodes.view(B, N, -1).transpose(1, 2)
        return ops.nodes_to_nchw(nodes, B, N)
