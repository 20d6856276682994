"""CUDA-graph capture of the eval forward.

The eval forward is ~95 kernel launches with static shapes; at the reference's fingerprinting call
shape (chunks of <= 128 segments, generate.py:40-46) the launches themselves dominate.  `GraphedEncoder`
records one forward into a CUDA graph (torch.cuda.graph only provides the capture stream and the
graph-private allocator; every node is a libgrafp_sm100a kernel) and replays it:

    g = GraphedEncoder(enc, batch=128)            # or GraphedSimCLR(model, batch=128)
    emb = g(x)                                    # copies x into the static input, replays, returns g.output
    g.input.copy_(x_host, non_blocking=True); g.replay(); out_host.copy_(g.output, non_blocking=True)

The graph reads the module's prepared (BatchNorm-folded, split) weights: re-capture after the
parameters change (optimizer step / load_state_dict).
"""
from __future__ import annotations

import torch

from . import ops


class GraphedEncoder:
    def __init__(self, enc, batch: int, nodes: int = 256, in_channels: int = None, return_pre_proj: bool = False,
                 warmup: int = 2):
        if enc.training:
            raise RuntimeError("GraphedEncoder captures the eval forward; call enc.eval() first")
        dev = next(enc.parameters()).device
        cin = in_channels if in_channels is not None else enc.stem[0].weight.shape[1]
        self.enc = enc
        self.input = torch.zeros((batch, cin, nodes), device=dev, dtype=torch.float32)
        self._rpp = return_pre_proj
        stream = torch.cuda.Stream(device=dev)
        stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(stream), torch.no_grad():
            for _ in range(warmup):                       # builds the prepared weights outside the capture
                enc(self.input, return_pre_proj=return_pre_proj)
        torch.cuda.current_stream(dev).wait_stream(stream)
        from . import _lib
        self.graph = torch.cuda.CUDAGraph()
        n0 = _lib.launch_count()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.output = enc(self.input, return_pre_proj=return_pre_proj)
        self.kernel_nodes = _lib.launch_count() - n0      # libgrafp kernels recorded in the graph

    def replay(self) -> None:
        self.graph.replay()

    def __call__(self, x: torch.Tensor):
        self.input.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.output


class GraphedSimCLR:
    """Spectrogram segments -> (h, z) of one view, captured (model(x, x) in generate.py only keeps z_i)."""

    def __init__(self, model, batch: int, warmup: int = 2):
        if model.training:
            raise RuntimeError("GraphedSimCLR captures the eval forward; call model.eval() first")
        cfg = model.cfg
        dev = next(model.parameters()).device
        self.model = model
        self.input = torch.zeros((batch, cfg["n_mels"], cfg["n_frames"]), device=dev, dtype=torch.float32)
        stream = torch.cuda.Stream(device=dev)
        stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(stream), torch.no_grad():
            for _ in range(warmup):
                model._one_view(self.input)
        torch.cuda.current_stream(dev).wait_stream(stream)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph), torch.no_grad():
            self.h, self.z = model._one_view(self.input)

    def replay(self) -> None:
        self.graph.replay()

    def __call__(self, x: torch.Tensor):
        self.input.copy_(x, non_blocking=True)
        self.graph.replay()
        return self.h, self.z
