"""CPU oracle for the GraphEncoder hot path -- TEST INFRASTRUCTURE, NOT PRODUCT.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this file.  The product package
(``neuralsampleid_b200``) never imports it and fails loudly when its CUDA library is
missing.

What it is
----------
A functional, state-dict driven restatement (PyTorch CPU, fp32) of the reference's
``GraphEncoder`` forward, the SimCLR wrapper around it and the NT-Xent loss.  The
reference's arithmetic lives in PyTorch ATen (conv2d / batch_norm / matmul / topk),
so the oracle calls the same ATen primitives in the same association order; it
shares no code with the reference modules -- there is no ``nn.Module`` here, only
functions over a ``{name: tensor}`` dict with the reference's ``state_dict`` keys.

Parity pin
----------
The reference ships no tests / golden vectors for this path.
The oracle is pinned instead against outputs of the reference itself:
``tests/golden/make_golden.py`` imports ``/root/reference`` (with stub modules for
the absent ``timm`` / ``torchmetrics``), runs it on seeded inputs and commits the
vectors under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks this file
against those vectors (bit-exact on CPU) and, where ``/root/reference`` exists,
against the live reference on every intermediate tensor.

Every function cites the reference file:line it follows (paths under
``/root/reference``).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
Params = Dict[str, Tensor]

# size -> (blocks, channels)                      encoder/graph_encoder.py:118-129
SIZES = {
    "t": ([2, 2, 6, 2], [64, 128, 256, 512]),
    "s": ([2, 2, 6, 2], [80, 160, 400, 640]),
    "m": ([2, 2, 16, 2], [96, 192, 384, 768]),
    "b": ([2, 2, 18, 2], [128, 256, 512, 1024]),
}
BN_EPS = 1e-5        # nn.BatchNorm2d default
BN_MOMENTUM = 0.1


def backbone_layout(size: str = "t") -> List[Tuple[str, int, int]]:
    """Sequence of backbone entries as built by encoder/graph_encoder.py:160-175.

    Returns [(kind, c_in, c_out)] with kind in {'block', 'down'}; index in the list
    == index in ``backbone`` (state_dict key ``backbone.<i>``)."""
    blocks, channels = SIZES.get(size, SIZES["b"])
    out = []
    for i, nb in enumerate(blocks):
        if i > 0:
            out.append(("down", channels[i - 1], channels[i]))
        for _ in range(nb):
            out.append(("block", channels[i], channels[i]))
    return out


# --------------------------------------------------------------------------- #
# primitive layers
# --------------------------------------------------------------------------- #
def _bn(p: Params, prefix: str, x: Tensor, training: bool, stats: Optional[dict]) -> Tensor:
    """nn.BatchNorm2d forward (eps 1e-5, momentum 0.1).  In training mode uses batch
    statistics and (if ``stats`` is given) records the updated running stats there
    instead of mutating ``p``."""
    w, b = p[prefix + ".weight"], p[prefix + ".bias"]
    rm, rv = p[prefix + ".running_mean"], p[prefix + ".running_var"]
    if not training:
        return F.batch_norm(x, rm, rv, w, b, False, BN_MOMENTUM, BN_EPS)
    rm2, rv2 = rm.detach().clone(), rv.detach().clone()
    y = F.batch_norm(x, rm2, rv2, w, b, True, BN_MOMENTUM, BN_EPS)
    if stats is not None:
        stats[prefix + ".running_mean"] = rm2
        stats[prefix + ".running_var"] = rv2
    return y


def _act(name: str, x: Tensor) -> Tensor:
    """act_layer, encoder/gcn_lib/torch_nn.py:9-25 (relu / leakyrelu(0.2) / gelu)."""
    name = name.lower()
    if name == "relu":
        return F.relu(x)
    if name == "leakyrelu":
        return F.leaky_relu(x, 0.2)
    if name == "gelu":
        return F.gelu(x)
    raise NotImplementedError("activation layer [%s] is not found" % name)


def l2_normalize_nodes(x: Tensor) -> Tensor:
    """F.normalize(x, p=2, dim=1) -- encoder/gcn_lib/torch_edge.py:281.  x: (B,C,N,1)."""
    return F.normalize(x, p=2.0, dim=1)


def pairwise_distance(x: Tensor) -> Tensor:
    """encoder/gcn_lib/torch_edge.py:7-18.  x: (B,N,C) -> (B,N,N);
    association order (sq + (-2 x x^T)) + sq^T."""
    inner = -2 * torch.matmul(x, x.transpose(2, 1))
    sq = torch.sum(torch.mul(x, x), dim=-1, keepdim=True)
    return sq + inner + sq.transpose(2, 1)


def dense_knn(x: Tensor, k: int) -> Tuple[Tensor, Tensor]:
    """dense_knn_matrix, encoder/gcn_lib/torch_edge.py:70-103 (N <= 10000 branch).
    x: (B,C,N,1) already normalised.  Returns (nn_idx (B,N,k) int64 sorted by
    ascending distance, dist (B,N,N))."""
    with torch.no_grad():
        xt = x.transpose(2, 1).squeeze(-1)
        dist = pairwise_distance(xt.detach())
        _, nn_idx = torch.topk(-dist, k=k)
    return nn_idx, dist


def dilated_knn_graph(x: Tensor, k: int, dilation: int = 1) -> Tuple[Tensor, Tensor]:
    """DenseDilatedKnnGraph.forward, encoder/gcn_lib/torch_edge.py:270-284 with
    DenseDilated (:245-255, stochastic=False): top-(k*d) then every d-th rank.
    Returns (edge_index (2,B,N,k) int64, dist of the normalised features)."""
    xn = l2_normalize_nodes(x)
    nn_idx, dist = dense_knn(xn, k * dilation)
    B, N, K = nn_idx.shape
    center = torch.arange(N).view(1, N, 1).expand(B, N, K)
    edge = torch.stack((nn_idx, center), dim=0)
    return edge[:, :, :, ::dilation], dist


def gather_nodes(x: Tensor, idx: Tensor) -> Tensor:
    """batched_index_select, encoder/gcn_lib/torch_nn.py:79-98.
    x: (B,C,N,1), idx: (B,N,k) -> (B,C,N,k)."""
    B, C, N = x.shape[:3]
    k = idx.shape[-1]
    xt = x.squeeze(-1).transpose(1, 2)                       # (B,N,C)
    flat = (idx + torch.arange(B).view(-1, 1, 1) * N).reshape(-1)
    feat = xt.reshape(B * N, C)[flat]
    return feat.view(B, idx.shape[1], k, C).permute(0, 3, 1, 2).contiguous()


def max_relative(x: Tensor, edge_index: Tensor) -> Tensor:
    """MRConv2d aggregation, encoder/gcn_lib/torch_vertex.py:21-29:
    max_k (x_j - x_i) -> (B,C,N,1)."""
    x_i = gather_nodes(x, edge_index[1])
    x_j = gather_nodes(x, edge_index[0])
    m, _ = torch.max(x_j - x_i, -1, keepdim=True)
    return m


def interleave(x: Tensor, m: Tensor) -> Tensor:
    """encoder/gcn_lib/torch_vertex.py:31-32: channels [x0, m0, x1, m1, ...]."""
    b, c, n, w = x.shape
    return torch.cat([x.unsqueeze(2), m.unsqueeze(2)], dim=2).reshape(b, 2 * c, n, w)


def basic_conv(p: Params, prefix: str, x: Tensor, act: str, training: bool,
               stats: Optional[dict]) -> Tensor:
    """BasicConv([2C, 2C], act, 'batch', bias): Conv2d 1x1 groups=4 -> BN -> act.
    encoder/gcn_lib/torch_nn.py:52-64."""
    y = F.conv2d(x, p[prefix + ".0.weight"], p.get(prefix + ".0.bias"), groups=4)
    y = _bn(p, prefix + ".1", y, training, stats)
    return _act(act, y)


def mr_conv(p: Params, prefix: str, x: Tensor, edge_index: Tensor, act: str,
            training: bool, stats: Optional[dict], taps: Optional[dict] = None) -> Tensor:
    """MRConv2d.forward, encoder/gcn_lib/torch_vertex.py:19-34."""
    m = max_relative(x, edge_index)
    if taps is not None:
        taps["max_rel"] = m
    return basic_conv(p, prefix + ".nn", interleave(x, m), act, training, stats)


def edge_conv(p: Params, prefix: str, x: Tensor, edge_index: Tensor, act: str,
              training: bool, stats: Optional[dict]) -> Tensor:
    """EdgeConv2d.forward, encoder/gcn_lib/torch_vertex.py:44-52 (optional path)."""
    x_i = gather_nodes(x, edge_index[1])
    x_j = gather_nodes(x, edge_index[0])
    y = basic_conv(p, prefix + ".nn", torch.cat([x_i, x_j - x_i], dim=1), act, training, stats)
    mx, _ = torch.max(y, -1, keepdim=True)
    return mx


def _basic_conv_any(p: Params, prefix: str, x: Tensor, act: str, training: bool,
                    stats: Optional[dict]) -> Tensor:
    """BasicConv with the norm layer optional (keys decide): Conv2d 1x1 groups=4 [-> BN] -> act,
    encoder/gcn_lib/torch_nn.py:52-64."""
    y = F.conv2d(x, p[prefix + ".0.weight"], p.get(prefix + ".0.bias"), groups=4)
    if prefix + ".1.weight" in p:
        y = _bn(p, prefix + ".1", y, training, stats)
    return _act(act, y)


def sage_conv(p: Params, prefix: str, x: Tensor, edge_index: Tensor, act: str,
              training: bool, stats: Optional[dict]) -> Tensor:
    """GraphSAGE.forward, encoder/gcn_lib/torch_vertex.py:60-68:
    nn2(cat[x, max_k nn1(x_j)])."""
    x_j = gather_nodes(x, edge_index[0])
    x_j, _ = torch.max(_basic_conv_any(p, prefix + ".nn1", x_j, act, training, stats), -1, keepdim=True)
    return _basic_conv_any(p, prefix + ".nn2", torch.cat([x, x_j], dim=1), act, training, stats)


def gin_conv(p: Params, prefix: str, x: Tensor, edge_index: Tensor, act: str,
             training: bool, stats: Optional[dict]) -> Tensor:
    """GINConv2d.forward, encoder/gcn_lib/torch_vertex.py:81-88:
    nn((1 + eps) * x + sum_k x_j)."""
    x_j = torch.sum(gather_nodes(x, edge_index[0]), -1, keepdim=True)
    return _basic_conv_any(p, prefix + ".nn", (1 + p[prefix + ".eps"]) * x + x_j, act, training, stats)


def dy_graph_conv(p: Params, prefix: str, x: Tensor, k: int, dilation: int, conv: str,
                  act: str, training: bool, stats: Optional[dict],
                  taps: Optional[dict] = None) -> Tensor:
    """DyGraphConv2d.forward (r == 1), encoder/gcn_lib/torch_vertex.py:126-139."""
    B, C, H, W = x.shape
    x = x.reshape(B, C, -1, 1).contiguous()
    edge_index, dist = dilated_knn_graph(x, k, dilation)
    if taps is not None:
        taps["knn_in"] = x
        taps["idx"] = edge_index[0]
        taps["dist"] = dist
    if conv == "mr":
        y = mr_conv(p, prefix + ".gconv", x, edge_index, act, training, stats, taps)
    elif conv == "edge":
        y = edge_conv(p, prefix + ".gconv", x, edge_index, act, training, stats)
    elif conv == "sage":
        y = sage_conv(p, prefix + ".gconv", x, edge_index, act, training, stats)
    elif conv == "gin":
        y = gin_conv(p, prefix + ".gconv", x, edge_index, act, training, stats)
    else:
        raise NotImplementedError("conv:{} is not supported".format(conv))
    return y.reshape(B, -1, H, W).contiguous()


def grapher(p: Params, prefix: str, x: Tensor, k: int, dilation: int = 1, conv: str = "mr",
            act: str = "relu", training: bool = False, stats: Optional[dict] = None,
            taps: Optional[dict] = None) -> Tensor:
    """Grapher.forward, encoder/gcn_lib/torch_vertex.py:183-195:
    x + BN(fc2(graph_conv(BN(fc1(x))))) ; drop_path is Identity."""
    y = F.conv2d(x, p[prefix + ".fc1.0.weight"], p[prefix + ".fc1.0.bias"])
    y = _bn(p, prefix + ".fc1.1", y, training, stats)
    if taps is not None:
        taps["fc1"] = y
    y = dy_graph_conv(p, prefix + ".graph_conv", y, k, dilation, conv, act, training, stats, taps)
    if taps is not None:
        taps["graph_conv"] = y
    y = F.conv2d(y, p[prefix + ".fc2.0.weight"], p[prefix + ".fc2.0.bias"])
    y = _bn(p, prefix + ".fc2.1", y, training, stats)
    return y + x


def ffn(p: Params, prefix: str, x: Tensor, act: str = "relu", training: bool = False,
        stats: Optional[dict] = None) -> Tensor:
    """FFN.forward, encoder/graph_encoder.py:82-89."""
    y = F.conv2d(x, p[prefix + ".fc1.0.weight"])
    y = _bn(p, prefix + ".fc1.1", y, training, stats)
    y = _act(act, y)
    y = F.conv2d(y, p[prefix + ".fc2.0.weight"])
    y = _bn(p, prefix + ".fc2.1", y, training, stats)
    return y + x


def downsample(p: Params, prefix: str, x: Tensor, training: bool = False,
               stats: Optional[dict] = None) -> Tensor:
    """Downsample.forward, encoder/graph_encoder.py:48-50 (3x3 stride-2 pad-1 conv + BN
    on a (N,1) image)."""
    y = F.conv2d(x, p[prefix + ".conv.0.weight"], p[prefix + ".conv.0.bias"], stride=2, padding=1)
    return _bn(p, prefix + ".conv.1", y, training, stats)


def stem(p: Params, x: Tensor, training: bool = False, stats: Optional[dict] = None) -> Tensor:
    """encoder/graph_encoder.py:151-153: Conv2d 1x1 (no bias) + BN + LeakyReLU(0.2)."""
    y = F.conv2d(x, p["stem.0.weight"])
    y = _bn(p, "stem.1", y, training, stats)
    return F.leaky_relu(y, 0.2)


# --------------------------------------------------------------------------- #
# encoder / wrapper / loss
# --------------------------------------------------------------------------- #
def encoder_forward(p: Params, x: Tensor, k: int = 3, size: str = "t", act: str = "relu",
                    training: bool = False, stats: Optional[dict] = None,
                    taps: Optional[list] = None, return_pre_proj: bool = False):
    """GraphEncoder.forward, encoder/graph_encoder.py:190-214.  x: (B,C_in,N) ->
    (B, emb_dims).  Every Grapher gets k=num_k[0], dilation 1, conv 'mr' (Q2).

    ``taps`` (a list) receives one dict per backbone entry with the intermediate
    tensors used by the teacher-forced parity tests."""
    x = x.unsqueeze(-1)
    x = stem(p, x, training, stats)
    if taps is not None:
        taps.append({"kind": "stem", "out": x})
    for i, (kind, _, _) in enumerate(backbone_layout(size)):
        pre = "backbone.%d" % i
        t = {"kind": kind, "in": x} if taps is not None else None
        if kind == "down":
            x = downsample(p, pre, x, training, stats)
        else:
            x = grapher(p, pre + ".0", x, k, 1, "mr", act, training, stats, t)
            if t is not None:
                t["grapher"] = x
            x = ffn(p, pre + ".1", x, act, training, stats)
        if t is not None:
            t["out"] = x
            taps.append(t)
    nodes = x
    x = F.conv2d(x, p["proj.weight"], p["proj.bias"])
    x = torch.mean(x, dim=2).squeeze(-1).squeeze(-1)
    if return_pre_proj:
        return nodes.squeeze(-1), x
    return x


def peak_extractor(p: Params, spec: Tensor, prefix: str = "peak_extractor") -> Tensor:
    """GPUPeakExtractorv2.forward, peak_extractor.py:45-69. spec: (B, n_mels, n_frames)."""
    mn = torch.amin(spec, dim=(1, 2), keepdim=True)
    mx = torch.amax(spec, dim=(1, 2), keepdim=True)
    s = (spec - mn) / (mx - mn)
    B, Fm, T = s.shape
    t_ramp = torch.linspace(0, 1, steps=T).view(1, 1, T).repeat(B, Fm, 1)
    f_ramp = torch.linspace(0, 1, steps=Fm).view(1, Fm, 1).repeat(B, 1, T)
    t = torch.cat((t_ramp.unsqueeze(1), f_ramp.unsqueeze(1), s.unsqueeze(1)), dim=1)
    w = p[prefix + ".convs.0.weight"]
    y = F.relu(F.conv2d(t, w, p[prefix + ".convs.0.bias"], stride=(w.shape[2], w.shape[3])))
    return y.reshape(B, y.shape[1], -1)


def projector(p: Params, h: Tensor, prefix: str = "projector") -> Tensor:
    """simclr/simclr.py:25-28,37-38: Linear -> ELU -> Linear -> F.normalize(eps=1e-10)."""
    z = F.linear(h, p[prefix + ".0.weight"], p[prefix + ".0.bias"])
    z = F.elu(z)
    z = F.linear(z, p[prefix + ".2.weight"], p[prefix + ".2.bias"])
    return F.normalize(z, p=2, eps=1e-10)


def simclr_forward(p: Params, x_i: Tensor, x_j: Tensor, k: int = 3, size: str = "t",
                   training: bool = False, stats: Optional[dict] = None):
    """SimCLR.forward (arch 'grafp'), simclr/simclr.py:31-47.  ``p`` holds the SimCLR
    state_dict (keys ``peak_extractor.*``, ``encoder.*``, ``projector.*``)."""
    enc = {n[len("encoder."):]: t for n, t in p.items() if n.startswith("encoder.")}
    outs = []
    for x in (x_i, x_j):
        g = peak_extractor(p, x)
        st = {} if stats is not None else None
        h = encoder_forward(enc, g, k=k, size=size, training=training, stats=st)
        if stats is not None:
            stats.update({"encoder." + n: t for n, t in st.items()})
        outs.append((h, projector(p, h)))
    return outs[0][0], outs[1][0], outs[0][1], outs[1][1]


def cross_attention_classifier(p: Params, x_i: Tensor, x_j: Tensor, num_heads: int = 4,
                               prefix: str = "") -> Tensor:
    """CrossAttentionClassifier.forward in eval mode, downstream.py:58-75: positional embedding, 
    nn.MultiheadAttention(query = x_i, key = value = x_j, batch_first), mean over the nodes, fc (Linear, ReLU,
    Dropout = identity, Linear, Sigmoid).  x_i, x_j: (B, C, N) -> (B, 1)."""
    xi, xj = x_i.permute(0, 2, 1), x_j.permute(0, 2, 1)                    # (B, N, C)
    if prefix + "positional_embedding" in p:
        pos = p[prefix + "positional_embedding"][:, :xi.shape[1], :]
        xi, xj = xi + pos, xj + pos
    E = xi.shape[-1]
    w, b = p[prefix + "attn.in_proj_weight"], p[prefix + "attn.in_proj_bias"]
    q = F.linear(xi, w[:E], b[:E])
    k = F.linear(xj, w[E:2 * E], b[E:2 * E])
    v = F.linear(xj, w[2 * E:], b[2 * E:])
    B, N, _ = q.shape
    dh = E // num_heads
    split = lambda t: t.reshape(B, N, num_heads, dh).transpose(1, 2)       # (B, H, N, dh)
    att = torch.softmax((split(q) * (1.0 / float(dh) ** 0.5)) @ split(k).transpose(-1, -2), dim=-1)
    o = (att @ split(v)).transpose(1, 2).reshape(B, N, E)
    o = F.linear(o, p[prefix + "attn.out_proj.weight"], p[prefix + "attn.out_proj.bias"])
    h = o.mean(dim=1)
    h = torch.relu(F.linear(h, p[prefix + "fc.0.weight"], p[prefix + "fc.0.bias"]))
    return torch.sigmoid(F.linear(h, p[prefix + "fc.3.weight"], p[prefix + "fc.3.bias"]))


def ntxent_loop(z_i: Tensor, z_j: Tensor, tau: float) -> Tensor:
    """ntxent_loss restated row by row exactly as simclr/ntxent.py:18-29 (small cases)."""
    z = torch.stack((z_i, z_j), dim=1).view(2 * z_i.shape[0], z_i.shape[1])
    a = torch.matmul(z, z.T) / tau
    rows = []
    for i in range(z.shape[0]):
        others = torch.cat([a[i, :i], a[i, i + 1:]])
        ls = F.log_softmax(others, dim=0)
        rows.append(ls[i if i % 2 == 0 else i - 1])
    return torch.sum(torch.stack(rows)) / -z.shape[0]


def ntxent(z_i: Tensor, z_j: Tensor, tau: float) -> Tensor:
    """Vectorised equivalent of simclr/ntxent.py:5-30:
    -mean_i( a[i, i^1] - logsumexp_{j != i} a[i, j] ),  a = z z^T / tau, rows of z
    interleaved (z_i[0], z_j[0], z_i[1], ...)."""
    z = torch.stack((z_i, z_j), dim=1).view(2 * z_i.shape[0], z_i.shape[1])
    a = torch.matmul(z, z.T) / tau
    n = a.shape[0]
    eye = torch.eye(n, dtype=torch.bool)
    lse = torch.logsumexp(a.masked_fill(eye, float("-inf")), dim=1)
    pos = a[torch.arange(n), torch.arange(n) ^ 1]
    return -(pos - lse).mean()


# --------------------------------------------------------------------------- #
# tie analysis used by the kNN parity tests
# --------------------------------------------------------------------------- #
def knn_tie_rows(dist: Tensor, kk: int, tol: float) -> Tensor:
    """Rows whose top-kk list is not uniquely determined at tolerance ``tol``: some pair
    of adjacent values among the (kk+1) smallest distances of the row differs by <= tol.
    dist: (B,N,N) reference distances.  Returns a bool (B,N) mask ("documented ties")."""
    kk1 = min(kk + 1, dist.shape[-1])
    vals, _ = torch.topk(-dist, k=kk1)
    vals = -vals
    gaps = vals[..., 1:] - vals[..., :-1]
    return (gaps <= tol).any(dim=-1)


class CascadeTracker:
    """Tie-aware comparison of two runs of the encoder (this is how "bit-exact except on
    documented distance ties" is checked end to end).

    The graph of block l is built from the features of block l-1, so one legitimately flipped
    near-tie neighbour changes everything downstream of it *for that segment*.  Walking the blocks
    in order, a segment stays ``alive`` while every neighbour list so far is identical; at the
    first block where a segment differs, all differing rows must be documented ties (adjacent
    reference distances within ``tol`` among the k*d+1 smallest), after which the segment is
    marked diverged and excluded from later checks.  ``bad`` counts off-tie mismatches."""

    def __init__(self, n_segments: int):
        self.alive = torch.ones(n_segments, dtype=torch.bool)
        self.bad = 0
        self.tie_flips = 0
        self.log = []

    def update(self, layer: int, idx_test: Tensor, idx_ref: Tensor, dist_ref: Tensor, kk: int,
               tol: float) -> None:
        diff = (idx_test.long() != idx_ref.long()).any(-1)            # (B, N)
        tie = knn_tie_rows(dist_ref, kk, tol)
        off = diff & ~tie & self.alive[:, None]
        self.bad += int(off.sum())
        flipped = (diff & self.alive[:, None]).any(-1)
        self.tie_flips += int((diff & tie & self.alive[:, None]).sum())
        if flipped.any():
            self.log.append((layer, [int(i) for i in torch.nonzero(flipped).flatten()], int(off.sum())))
        self.alive &= ~flipped


def clip_grad_norm_(grads: List[Tensor], max_norm: float) -> Tensor:
    """torch.nn.utils.clip_grad_norm_ semantics used at train.py:73 (L2, eps 1e-6)."""
    total = torch.sqrt(sum((g.detach().double() ** 2).sum() for g in grads)).float()
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    for g in grads:
        g.mul_(coef)
    return total


def adam_step(param: Tensor, grad: Tensor, m: Tensor, v: Tensor, step: int, lr: float,
              b1: float = 0.9, b2: float = 0.999, eps: float = 1e-8) -> None:
    """torch.optim.Adam (no weight decay, no amsgrad) single-tensor update, train.py:126."""
    m.mul_(b1).add_(grad, alpha=1 - b1)
    v.mul_(b2).addcmul_(grad, grad, value=1 - b2)
    bc1 = 1 - b1 ** step
    bc2 = 1 - b2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    param.addcdiv_(m, denom, value=-lr / bc1)
