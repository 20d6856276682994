"""Portable synthetic weights / inputs for parity tests (TEST INFRASTRUCTURE, see
grafp_oracle.py header).

torch's CPU RNG streams are build- and ISA-dependent, so golden vectors are not keyed
on ``torch.manual_seed``.  Instead every tensor is derived from numpy's PCG64 *integer*
stream mapped to floats by exact arithmetic -- bit-identical on every machine -- and
loaded into the reference / the oracle / the CUDA modules through ``state_dict``.

``state_spec`` restates the reference's ``state_dict`` layout (names, shapes, order):
encoder/graph_encoder.py:151-179 (stem, backbone, proj), encoder/gcn_lib/torch_vertex.py:
146-172 (Grapher: relative_pos, fc1, graph_conv.gconv.nn, fc2), encoder/graph_encoder.py:
38-89 (Downsample, FFN), simclr/simclr.py:13-28 + peak_extractor.py:14-23 (wrapper).
"""
from __future__ import annotations

import hashlib
from typing import Dict, List, Tuple

import numpy as np
import torch

from .grafp_oracle import SIZES, backbone_layout

Spec = List[Tuple[str, Tuple[int, ...], str]]   # (name, shape, role)


def _bn(prefix: str, c: int) -> Spec:
    return [(prefix + ".weight", (c,), "bn_w"), (prefix + ".bias", (c,), "bn_b"),
            (prefix + ".running_mean", (c,), "bn_rm"), (prefix + ".running_var", (c,), "bn_rv"),
            (prefix + ".num_batches_tracked", (), "count")]


def encoder_state_spec(size: str = "t", in_channels: int = 8, emb_dims: int = 1024,
                       n_nodes: int = 256) -> Spec:
    """Ordered (name, shape, role) list of GraphEncoder.state_dict()."""
    _, channels = SIZES.get(size, SIZES["b"])
    spec: Spec = [("stem.0.weight", (channels[0], in_channels, 1, 1), "w")]
    spec += _bn("stem.1", channels[0])
    n_rel = n_nodes
    for i, (kind, cin, cout) in enumerate(backbone_layout(size)):
        pre = "backbone.%d" % i
        if kind == "down":
            n_rel = n_rel // 4                                   # graph_encoder.py:166
            spec += [(pre + ".conv.0.weight", (cout, cin, 3, 3), "w"),
                     (pre + ".conv.0.bias", (cout,), "b")]
            spec += _bn(pre + ".conv.1", cout)
            continue
        c = cin
        g = pre + ".0"
        spec += [(g + ".relative_pos", (1, n_rel, n_rel), "relpos"),
                 (g + ".fc1.0.weight", (c, c, 1, 1), "w"), (g + ".fc1.0.bias", (c,), "b")]
        spec += _bn(g + ".fc1.1", c)
        spec += [(g + ".graph_conv.gconv.nn.0.weight", (2 * c, (2 * c) // 4, 1, 1), "w"),
                 (g + ".graph_conv.gconv.nn.0.bias", (2 * c,), "b")]
        spec += _bn(g + ".graph_conv.gconv.nn.1", 2 * c)
        spec += [(g + ".fc2.0.weight", (c, 2 * c, 1, 1), "w"), (g + ".fc2.0.bias", (c,), "b")]
        spec += _bn(g + ".fc2.1", c)
        f = pre + ".1"
        spec += [(f + ".fc1.0.weight", (4 * c, c, 1, 1), "w")]
        spec += _bn(f + ".fc1.1", 4 * c)
        spec += [(f + ".fc2.0.weight", (c, 4 * c, 1, 1), "w")]
        spec += _bn(f + ".fc2.1", c)
    spec += [("proj.weight", (emb_dims, channels[-1], 1, 1), "w"), ("proj.bias", (emb_dims,), "b")]
    return spec


def simclr_state_spec(cfg: dict, size: str = "t") -> Spec:
    """Ordered spec of SimCLR(cfg, GraphEncoder(...)).state_dict() for arch 'grafp'.
    Module registration order in simclr/simclr.py:8-28: encoder, peak_extractor, projector."""
    n_nodes = cfg["n_mels"] * cfg["n_frames"] // (cfg["patch_bins"] * cfg["patch_frames"])
    enc = [("encoder." + n, s, r) for n, s, r in
           encoder_state_spec(size, cfg["n_filters"], cfg["h"], n_nodes)]
    pk = [("peak_extractor.convs.0.weight",
           (cfg["n_filters"], 3, cfg["patch_bins"], cfg["patch_frames"]), "w"),
          ("peak_extractor.convs.0.bias", (cfg["n_filters"],), "b")]
    du = cfg["d"] * cfg["u"]
    pr = [("projector.0.weight", (du, cfg["h"]), "w"), ("projector.0.bias", (du,), "b"),
          ("projector.2.weight", (cfg["d"], du), "w"), ("projector.2.bias", (cfg["d"],), "b")]
    return enc + pk + pr


def graphconv_state_spec(conv: str, cin: int, cout: int, signed_bn: bool = True) -> Spec:
    """state_dict of DyGraphConv2d(cin, cout, ..., conv, act, 'batch', True) for the reference's four
    GraphConv2d variants (encoder/gcn_lib/torch_vertex.py:11-111); BasicConv = Conv2d(groups=4)+BN."""
    def bn(prefix, c):
        spec = _bn(prefix, c)
        return [(n, sh, "bn_ws" if (signed_bn and r == "bn_w") else r) for n, sh, r in spec]

    def basic(prefix, ci, co):
        return [(prefix + ".0.weight", (co, ci // 4, 1, 1), "w"), (prefix + ".0.bias", (co,), "b")] + bn(prefix + ".1", co)
    if conv in ("mr", "edge"):
        return basic("gconv.nn", 2 * cin, cout)
    if conv == "sage":
        return basic("gconv.nn1", cin, cin) + basic("gconv.nn2", 2 * cin, cout)
    if conv == "gin":
        return [("gconv.eps", (1,), "eps")] + basic("gconv.nn", cin, cout)
    raise ValueError(conv)


def reranker_state_spec(in_dim: int = 512, hidden: int = 128, num_nodes: int = 100) -> Spec:
    """state_dict of CrossAttentionClassifier(in_dim, 4, hidden, num_nodes, True) (downstream.py:30-56)."""
    return [("positional_embedding", (1, num_nodes, in_dim), "pos"),
            ("attn.in_proj_weight", (3 * in_dim, in_dim), "w"), ("attn.in_proj_bias", (3 * in_dim,), "b"),
            ("attn.out_proj.weight", (in_dim, in_dim), "w"), ("attn.out_proj.bias", (in_dim,), "b"),
            ("fc.0.weight", (hidden, in_dim), "w"), ("fc.0.bias", (hidden,), "b"),
            ("fc.3.weight", (1, hidden), "w"), ("fc.3.bias", (1,), "b")]


def _uniform(rng: np.random.Generator, shape, lo: float, hi: float) -> np.ndarray:
    """Exact-arithmetic uniform floats: 24-bit integers / 2^24 in float64, affine, -> fp32."""
    n = int(np.prod(shape)) if len(shape) else 1
    u = rng.integers(0, 1 << 24, size=n, dtype=np.int64).astype(np.float64) / float(1 << 24)
    return (lo + (hi - lo) * u).astype(np.float32).reshape(shape)


def synth_state(spec: Spec, seed: int = 1234) -> Dict[str, torch.Tensor]:
    """Deterministic, portable state_dict for ``spec``.  BN statistics / affine terms and
    biases are non-trivial so that BN folding and bias paths are really exercised."""
    out: Dict[str, torch.Tensor] = {}
    for i, (name, shape, role) in enumerate(spec):
        rng = np.random.Generator(np.random.PCG64([seed, i]))
        if role == "w":
            fan_in = int(np.prod(shape[1:]))
            b = float(np.sqrt(3.0 / fan_in))              # unit-gain: var = 1/fan_in
            a = _uniform(rng, shape, -b, b)
        elif role == "b":
            a = _uniform(rng, shape, -0.1, 0.1)
        elif role == "bn_w":
            a = _uniform(rng, shape, 0.6, 1.2)
        elif role == "bn_b":
            a = _uniform(rng, shape, -0.2, 0.2)
        elif role == "bn_rm":
            a = _uniform(rng, shape, -0.3, 0.3)
        elif role == "bn_rv":
            a = _uniform(rng, shape, 0.5, 1.5)
        elif role == "bn_ws":                             # signed BatchNorm scale (max over edges then is
            a = _uniform(rng, shape, -1.2, 1.2)           # not a monotone function of the pre-activation)
        elif role == "pos":
            a = _uniform(rng, shape, -1.0, 1.0)
        elif role == "eps":
            a = _uniform(rng, shape, 0.1, 0.4)
        elif role == "count":
            out[name] = torch.zeros((), dtype=torch.int64)
            continue
        elif role == "relpos":
            a = np.zeros(shape, dtype=np.float32)         # never read on the hot path (Q7)
        else:
            raise ValueError(role)
        out[name] = torch.from_numpy(a.copy())
    return out


def synth_uniform(shape, seed: int, lo: float = 0.0, hi: float = 1.0) -> torch.Tensor:
    rng = np.random.Generator(np.random.PCG64([seed, 0xC0FFEE]))
    return torch.from_numpy(_uniform(rng, tuple(shape), lo, hi).copy())


def synth_normal(shape, seed: int) -> torch.Tensor:
    """Portable ~N(0,1): sum of 12 exact uniforms - 6 (Irwin-Hall), float64 -> fp32."""
    rng = np.random.Generator(np.random.PCG64([seed, 0xBEEF]))
    n = int(np.prod(shape))
    u = rng.integers(0, 1 << 24, size=(12, n), dtype=np.int64).astype(np.float64) / float(1 << 24)
    return torch.from_numpy((u.sum(0) - 6.0).astype(np.float32).reshape(tuple(shape)).copy())


def state_sha256(sd: Dict[str, torch.Tensor]) -> str:
    h = hashlib.sha256()
    for name in sd:
        h.update(name.encode())
        h.update(sd[name].detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def reranker_inputs(B: int, N: int, seed: int, in_dim: int = 512):
    """(x_i, x_j) node matrices (B, in_dim, N) for the re-ranker: the first half are perturbed copies (matching
    pairs), the second half unrelated.  Portable (same PCG64 streams everywhere), so fixtures store outputs only."""
    x_i = synth_normal((B, in_dim, N), seed)
    x_j = x_i + 0.5 * synth_normal((B, in_dim, N), seed + 1)
    x_j[B // 2:] = synth_normal((B - B // 2, in_dim, N), seed + 2)
    return x_i, x_j


def synth_wave(n_samples: int, seed: int) -> torch.Tensor:
    """Portable mono test waveform (exactly reproducible everywhere: integer noise stream + integer-pattern square
    waves, no transcendental functions): broadband noise, a 444 Hz and a 2 kHz square tone with a slow gate."""
    n = torch.arange(n_samples)
    sq1 = (((n // 18) % 2) * 2 - 1).float()              # 16000 / 36 = 444.4 Hz
    sq2 = (((n // 4) % 2) * 2 - 1).float()               # 2 kHz
    gate = ((n // 4096) % 3 != 1).float()
    return 0.3 * synth_normal((n_samples,), seed) + 0.4 * sq1 * gate + 0.1 * sq2
