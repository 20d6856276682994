"""Train-mode forward and the hand-written backward of the SimCLR(GraphEncoder) path.

The reference trains through PyTorch autograd over ~40 ATen kernels per block (train.py:61-70);
here each view is ONE autograd node: the forward runs the node-major kernel pipeline in train mode
(raw GEMM -> batch statistics -> fused normalise/activation/shortcut) while recording a tape, and
the backward replays the tape with the backward kernels of csrc/train.cu, returning gradients for
every parameter.  The kNN graph carries no gradient (it is built under no_grad in the reference,
encoder/gcn_lib/torch_edge.py:78,96); max-relative routes the gradient to the arg-max neighbour.

Weight re-layouts between the reference's parameter shapes and the (n, k) GEMM operands (the MRConv
even/odd column regrouping, the Downsample centre-column extraction, transposes for the input
gradient) are O(parameters) index operations done with torch on the parameter side.
"""
from __future__ import annotations

import weakref
from typing import Dict, List, Optional

import torch
from torch import nn

from . import _prep, ops
from ._prep import tap3_weight


class _Layer:
    """Tape entry of one  out = act(BN(A W^T [+ bias])) [+ residual]  layer."""
    __slots__ = ("a1", "a2", "raw", "ssmi", "act", "slope", "bn", "w2d", "groups", "tap3_nodes",
                 "weight", "bias", "kind", "k1", "k2", "has_residual")


def _w_split(w2d: torch.Tensor, groups: int, k_parts: int):
    """Split copies (keyword arguments of ops.gemm) of a GEMM operand when its shape suits the tcgen05 engines."""
    n_total, k = w2d.shape
    if (k // k_parts) % 32 == 0 and (n_total // groups) % 32 == 0:
        return ops.tc_splits(w2d)
    return {}


class GradSink(dict):
    """A gradient dict with in-place destinations: ``views[p]`` is where the gradient of parameter p accumulates (a
    slice of the fused optimizer's flat gradient buffer).  Kernels that accumulate (weight-gradient GEMM, BatchNorm
    parameter gradients) write there directly; everything else is added with one grafp_add_inplace -- no per-parameter
    temporaries and no second accumulation pass."""

    def __init__(self, views):
        super().__init__()
        self.views = views

    def view(self, p):
        return self.views.get(p) if p is not None and p.requires_grad else None


def _sink_view(grads, p):
    return grads.view(p) if isinstance(grads, GradSink) else None


# Per-step operand cache of the train path: derived weight layouts (MRConv even/odd regrouping, Downsample centre
# column, transposes for the input gradient) and their tensor-core splits are functions of the parameters alone, and
# both SimCLR views use the same parameters: built once per weight version instead of once per layer call.
_WCACHE = {"epoch": -1, "items": {}}


def _cached(param: torch.Tensor, tag, make):
    ep = (_prep.weights_epoch(), ops.get_engine(), ops._engine_override)
    if _WCACHE["epoch"] != ep:
        _WCACHE["epoch"], _WCACHE["items"] = ep, {}
    # keyed by the parameter OBJECT (a freed model's storage address can be handed to the next model's parameters),
    # validated by storage address and version; a newer version replaces the entry, so nothing accumulates
    key = (id(param), tag)
    stamp = (param.data_ptr(), param._version)
    hit = _WCACHE["items"].get(key)
    if hit is None or hit[0]() is not param or hit[1] != stamp:
        hit = (weakref.ref(param), stamp, make())
        _WCACHE["items"][key] = hit
        if len(_WCACHE["items"]) > 4096:                # dead models' entries
            _WCACHE["items"] = {k: v for k, v in _WCACHE["items"].items() if v[0]() is not None}
    return hit[2]


def layer_fwd(tape: List[_Layer], a1: torch.Tensor, weight: nn.Parameter, w2d: torch.Tensor, kind: str,
              bias: Optional[nn.Parameter] = None, bn: Optional[nn.BatchNorm2d] = None, act=None,
              slope: float = 0.0, residual: Optional[torch.Tensor] = None, a2: Optional[torch.Tensor] = None,
              groups: int = 1, tap3_nodes: int = 0) -> torch.Tensor:
    w2d = w2d.contiguous()
    sp = _cached(weight, ("fwd", kind), lambda: (w2d, _w_split(w2d, groups, 2 if a2 is not None else 1)))
    w2d = sp[0]
    raw = ops.gemm(a1, w2d, None, None, None, 0.0, None, a2, groups, tap3_nodes, None, None, **sp[1])
    M, C = raw.shape
    if bn is not None:
        stats = ops.col_stats(raw)
        if C % 4 == 0 and C <= 8192:
            out, ssmi = ops.bn_finalize_apply(stats, raw, bn.weight.detach(), bn.bias.detach(),
                                              bias.detach() if bias is not None else None, bn.eps, bn.momentum,
                                              bn.running_mean, bn.running_var, act, slope, residual)
        else:
            ssmi = ops.bn_finalize(stats, M, bn.weight.detach(), bn.bias.detach(),
                                   bias.detach() if bias is not None else None, bn.eps, bn.momentum,
                                   bn.running_mean, bn.running_var)
            out = ops.affine_act(raw, ssmi[0], ssmi[1], act, slope, residual)
        bn.num_batches_tracked += 1
        # running_mean / running_var were just updated through raw pointers (torch's version counters do not see
        # it): invalidate the eval path's folded (scale, shift) caches, which are keyed on _prep.sig()
        _prep.bump_epoch()
    else:
        ssmi = torch.zeros((4, C), device=raw.device, dtype=torch.float32)
        ssmi[0].fill_(1.0)
        ssmi[3].fill_(1.0)
        if bias is not None:
            ssmi[1].copy_(bias.detach())
        out = ops.affine_act(raw, ssmi[0], ssmi[1], act, slope, residual)
    L = _Layer()
    L.a1, L.a2, L.raw, L.ssmi, L.act, L.slope, L.bn = a1, a2, raw, ssmi, act, slope, bn
    L.w2d, L.groups, L.tap3_nodes, L.weight, L.bias, L.kind = w2d, groups, tap3_nodes, weight, bias, kind
    L.k1 = (3 * a1.shape[1]) if tap3_nodes else a1.shape[1] // groups
    L.k2 = a2.shape[1] // groups if a2 is not None else 0
    L.has_residual = residual is not None
    tape.append(L)
    return out


def _acc(grads: Dict, p, g: torch.Tensor) -> None:
    if p is None or not p.requires_grad:
        return
    v = _sink_view(grads, p)
    if v is not None:
        ops.add_inplace(v, g.reshape(v.shape).contiguous())
        return
    g = g.reshape(p.shape)
    if p in grads:
        ops.add_inplace(grads[p], g.contiguous())
    else:
        grads[p] = g.contiguous()


def _weight_grad_to_param(L: _Layer, dw2d: torch.Tensor) -> torch.Tensor:
    """(n, k) GEMM-operand gradient -> the reference parameter's layout."""
    w = L.weight
    if L.kind == "mr":            # columns were regrouped [even | odd] per group
        half = dw2d.shape[1] // 2
        out = torch.empty_like(dw2d)
        out[:, 0::2] = dw2d[:, :half]
        out[:, 1::2] = dw2d[:, half:]
        return out.reshape(w.shape)
    if L.kind == "tap3":          # (Cout, 3*Cin) -> centre column of the (Cout, Cin, 3, 3) kernel
        cout, cin = w.shape[0], w.shape[1]
        g = torch.zeros_like(w)
        g[:, :, :, 1] = dw2d.view(cout, 3, cin).permute(0, 2, 1)
        return g
    return dw2d.reshape(w.shape)


def layer_bwd(L: _Layer, dout: torch.Tensor, grads: Dict, need_input: bool = True,
              add_to: Optional[torch.Tensor] = None):
    """Returns (da1, da2).  The shortcut gradient of a residual layer is `dout` itself (caller's)."""
    bn = L.bn is not None
    # the parameter gradients (dgamma / dbeta, or the bias gradient of a layer without BatchNorm) are accumulated by
    # the apply kernel itself: into the flat gradient buffer when there is one, else into fresh tensors
    pg = L.bn.weight if bn else None
    pb = L.bn.bias if bn else L.bias
    vg, vb = _sink_view(grads, pg), _sink_view(grads, pb)
    tg = None if (pg is None or not pg.requires_grad or vg is not None) else torch.zeros_like(pg)
    tb = None if (pb is None or not pb.requires_grad or vb is not None) else torch.zeros_like(pb)
    draw, sums = ops.bn_act_bwd(dout, L.raw, L.ssmi, L.act, L.slope, bn,
                                vg if vg is not None else tg, vb if vb is not None else tb)
    if tg is not None:
        _acc(grads, pg, tg)
    if tb is not None:
        _acc(grads, pb, tb)
    if bn:
        # a conv bias in front of a train-mode BatchNorm has exactly zero gradient
        if L.bias is not None and L.bias.requires_grad and L.bias not in grads and _sink_view(grads, L.bias) is None:
            grads[L.bias] = torch.zeros_like(L.bias)
    if L.weight.requires_grad:
        vw = _sink_view(grads, L.weight)
        if vw is not None and L.kind == "dense":
            # the GEMM-operand layout IS the parameter's: the kernel accumulates into the flat gradient buffer
            ops.gemm_wgrad(draw, L.a1, L.a2, L.w2d.shape[0], L.groups, L.tap3_nodes, out=vw.view(L.w2d.shape))
        else:
            dw = ops.gemm_wgrad(draw, L.a1, L.a2, L.w2d.shape[0], L.groups, L.tap3_nodes)
            _acc(grads, L.weight, _weight_grad_to_param(L, dw))
    if not need_input:
        return None, None
    n = L.w2d.shape[0] // L.groups
    if L.tap3_nodes:
        cin = L.k1 // 3

        def make_t():
            wT = L.w2d.t().contiguous()                               # (3*Cin, Cout)
            return wT, _w_split(wT, 1, 1)
        wT, sp = _cached(L.weight, ("T", L.kind), make_t)
        dA = ops.gemm(draw, wT, **sp)
        return ops.tap3_bwd_input(dA, L.tap3_nodes, cin), None

    def make_t1():
        wg = L.w2d.view(L.groups, n, L.k1 + L.k2)
        wT1 = wg[:, :, :L.k1].transpose(1, 2).reshape(L.groups * L.k1, n).contiguous()
        return wT1, _w_split(wT1, L.groups, 1)
    wT1, sp = _cached(L.weight, ("T1", L.kind), make_t1)
    da1 = ops.gemm(draw, wT1, residual=add_to, groups=L.groups, **sp)
    da2 = None
    if L.k2:
        def make_t2():
            wg = L.w2d.view(L.groups, n, L.k1 + L.k2)
            wT2 = wg[:, :, L.k1:].transpose(1, 2).reshape(L.groups * L.k2, n).contiguous()
            return wT2, _w_split(wT2, L.groups, 1)
        wT2, sp = _cached(L.weight, ("T2", L.kind), make_t2)
        da2 = ops.gemm(draw, wT2, groups=L.groups, **sp)
    return da1, da2


# ------------------------------------------------------------------------------------------
# GraphEncoder
# ------------------------------------------------------------------------------------------
def _conv2d_w(conv: nn.Conv2d) -> torch.Tensor:
    return conv.weight.detach().reshape(conv.weight.shape[0], -1)


class _EncTape:
    __slots__ = ("layers", "blocks", "B", "N_out", "mean", "enc")


def encoder_train_fwd(enc, x_nodes: torch.Tensor, B: int, N: int, forced_idx=None):
    """Train-mode forward from node-major (B*N, C_in) features.  Returns (emb, nodes, tape).
    ``forced_idx``: optional per-block int32 (B, N, k) neighbour lists (parity-test hook)."""
    from .encoder.graph_encoder import Downsample
    tape = _EncTape()
    tape.layers, tape.blocks, tape.B, tape.enc = [], [], B, enc
    Ls = tape.layers
    stem = enc.stem
    h = layer_fwd(Ls, x_nodes, stem[0].weight, _conv2d_w(stem[0]), "dense", None, stem[1], "leakyrelu",
                  stem[2].negative_slope)
    tape.blocks.append(("stem", 1))
    for entry in enc.backbone:
        if isinstance(entry, Downsample):
            conv, bn = entry.conv[0], entry.conv[1]
            h = layer_fwd(Ls, h, conv.weight, _cached(conv.weight, "tap3w", lambda: tap3_weight(conv.weight)), "tap3",
                          conv.bias, bn, tap3_nodes=N // 2)
            N //= 2
            tape.blocks.append(("down", 1))
            continue
        g, f = entry[0], entry[1]
        y = layer_fwd(Ls, h, g.fc1[0].weight, _conv2d_w(g.fc1[0]), "dense", g.fc1[0].bias, g.fc1[1])
        gc = g.graph_conv
        if forced_idx is not None:
            idx = forced_idx[sum(1 for b in tape.blocks if b[0] == "block")]
        else:
            idx = gc.dilated_knn_graph.knn_nodes(y, B, N)
        m, arg = ops.mr_aggregate(y, idx, B, N, want_arg=True)
        mr = gc.gconv.nn
        conv, bnm = mr[0], mr[1]
        w_mr = _cached(conv.weight, "mrw", lambda: torch.cat([_conv2d_w(conv)[:, 0::2], _conv2d_w(conv)[:, 1::2]], dim=1))
        act = mr[2] if len(mr) > 2 else None
        u = layer_fwd(Ls, y, conv.weight, w_mr, "mr", conv.bias, bnm, act.name if act else None,
                      act.neg_slope if act else 0.0, a2=m, groups=mr.GROUPS)
        h2 = layer_fwd(Ls, u, g.fc2[0].weight, _conv2d_w(g.fc2[0]), "dense", g.fc2[0].bias, g.fc2[1], residual=h)
        t = layer_fwd(Ls, h2, f.fc1[0].weight, _conv2d_w(f.fc1[0]), "dense", None, f.fc1[1], f.act.name,
                      f.act.neg_slope)
        h = layer_fwd(Ls, t, f.fc2[0].weight, _conv2d_w(f.fc2[0]), "dense", None, f.fc2[1], residual=h2)
        tape.blocks.append(("block", (idx, arg, N)))
    mean = ops.node_mean(h, B, N)
    emb = layer_fwd(Ls, mean, enc.proj.weight, _conv2d_w(enc.proj), "dense", enc.proj.bias, None)
    tape.N_out = N
    return emb, h, tape


def encoder_train_bwd_steps(enc, tape: _EncTape, demb: torch.Tensor, grads: Dict,
                            dnodes: Optional[torch.Tensor] = None, need_input: bool = False, out: Optional[list] = None):
    """Backward of encoder_train_fwd as a generator: yields the sub-module (enc.proj, each backbone entry from the
    last to the first, enc.stem) whose parameter gradients have just been completed, so a data-parallel driver can
    start reducing them while the rest of the backward runs.  d(x_nodes) is appended to ``out`` if need_input."""
    Ls = list(tape.layers)
    B = tape.B
    dmean, _ = layer_bwd(Ls.pop(), demb.contiguous(), grads)
    dh = ops.node_mean_bwd(dmean, B, tape.N_out)
    if dnodes is not None:
        ops.add_inplace(dh, dnodes.contiguous())
    yield enc.proj
    entries = list(enc.backbone)
    ei = len(entries)
    for kind, info in reversed(tape.blocks):
        if kind == "block":
            idx, arg, N = info
            L_f2, L_f1, L_fc2, L_mr, L_fc1 = Ls.pop(), Ls.pop(), Ls.pop(), Ls.pop(), Ls.pop()
            dt, _ = layer_bwd(L_f2, dh, grads)                       # shortcut grad to h2: dh
            dh2, _ = layer_bwd(L_f1, dt, grads, add_to=dh)           # dh2 = dA + dh
            du, _ = layer_bwd(L_fc2, dh2, grads)                     # shortcut grad to h: dh2
            dy, dm = layer_bwd(L_mr, du, grads)
            ops.mr_aggregate_bwd(dm, idx, arg, B, N, dy)             # dy += scatter(dm)
            dh, _ = layer_bwd(L_fc1, dy, grads, add_to=dh2)
            ei -= 1
            yield entries[ei]
        elif kind == "down":
            dh, _ = layer_bwd(Ls.pop(), dh, grads)
            ei -= 1
            yield entries[ei]
        else:                                                        # stem
            dx, _ = layer_bwd(Ls.pop(), dh, grads, need_input=need_input)
            if out is not None:
                out.append(dx)
            yield enc.stem
            return


def encoder_train_bwd(tape: _EncTape, demb: torch.Tensor, grads: Dict, dnodes: Optional[torch.Tensor] = None,
                      need_input: bool = False, enc=None):
    """Backward of encoder_train_fwd.  Returns d(x_nodes) if need_input."""
    out: list = []
    for _ in encoder_train_bwd_steps(enc if enc is not None else tape.enc, tape, demb, grads, dnodes, need_input, out):
        pass
    return out[0] if out else None


class _EncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, enc, x, return_nodes, *params):
        B, _, N = x.shape
        emb, nodes, tape = encoder_train_fwd(enc, ops.nchw_to_nodes(x.detach()), B, N)
        ctx.tape, ctx.params, ctx.need_x = tape, params, x.requires_grad
        ctx.B, ctx.N = B, N
        if return_nodes:
            return emb, ops.nodes_to_nchw(nodes, B, tape.N_out)
        return emb

    @staticmethod
    def backward(ctx, demb, dnodes=None):
        grads: Dict = {}
        dn = None
        if dnodes is not None:
            dn = ops.nchw_to_nodes(dnodes.contiguous())
        dx = encoder_train_bwd(ctx.tape, demb, grads, dn, ctx.need_x)
        ctx.tape = None
        gx = ops.nodes_to_nchw(dx, ctx.B, ctx.N) if (ctx.need_x and dx is not None) else None
        return (None, gx, None) + tuple(grads.get(p) for p in ctx.params)


def encoder_forward_train(enc, x: torch.Tensor, return_pre_proj: bool = False):
    params = tuple(enc.parameters())
    out = _EncoderFn.apply(enc, x, return_pre_proj, *params)
    if return_pre_proj:
        emb, nodes = out
        return nodes, emb
    return out


# ------------------------------------------------------------------------------------------
# SimCLR view: peak extractor -> encoder -> projector -> normalise
# ------------------------------------------------------------------------------------------
class _ViewCtx:
    __slots__ = ("tape", "ptape", "z2", "spec")


def view_fwd(model, spec: torch.Tensor, forced_idx=None):
    """One SimCLR view in train mode -> (h, z, ctx).  ``forced_idx``: parity-test hook (per-block
    neighbour lists to use instead of the computed graph)."""
    pe = model.peak_extractor.convs[0]
    spec = spec.detach().contiguous()
    B = spec.shape[0]
    nodes = ops.peak_extract(spec, pe.weight.detach(), pe.bias.detach())
    N = nodes.shape[0] // B
    h, _, tape = encoder_train_fwd(model.encoder, nodes, B, N, forced_idx)
    ptape: List[_Layer] = []
    l0, l2 = model.projector[0], model.projector[2]
    z1 = layer_fwd(ptape, h, l0.weight, l0.weight.detach(), "dense", l0.bias, None, "elu")
    z2 = layer_fwd(ptape, z1, l2.weight, l2.weight.detach(), "dense", l2.bias, None)
    z = ops.l2_normalize_rows(z2, 1e-10)
    c = _ViewCtx()
    c.tape, c.ptape, c.z2, c.spec = tape, ptape, z2, spec
    return h, z, c


def view_bwd_steps(model, c: _ViewCtx, dh: Optional[torch.Tensor], dz: torch.Tensor, grads: Dict):
    """Backward of one SimCLR view as a generator yielding the sub-modules whose parameter gradients are complete:
    model.projector, then encoder_train_bwd_steps' sequence, then model.peak_extractor."""
    dz2 = ops.l2_normalize_rows_bwd(c.z2, dz.contiguous(), 1e-10)
    dz1, _ = layer_bwd(c.ptape[1], dz2, grads)
    dhh, _ = layer_bwd(c.ptape[0], dz1, grads, add_to=dh.contiguous() if dh is not None else None)
    yield model.projector
    out: list = []
    yield from encoder_train_bwd_steps(model.encoder, c.tape, dhh, grads, None, True, out)
    dnodes = out[0]
    pe = model.peak_extractor.convs[0]
    if pe.weight.requires_grad:
        dw, db = ops.peak_extract_bwd(c.spec, pe.weight.detach(), pe.bias.detach(), dnodes)
        _acc(grads, pe.weight, dw)
        _acc(grads, pe.bias, db)
    yield model.peak_extractor


def view_bwd(model, c: _ViewCtx, dh: Optional[torch.Tensor], dz: torch.Tensor, grads: Dict) -> None:
    for _ in view_bwd_steps(model, c, dh, dz, grads):
        pass


class _ViewFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, spec, forced_idx, *params):
        h, z, c = view_fwd(model, spec, forced_idx)
        ctx.c, ctx.params, ctx.model = c, params, model
        return h, z

    @staticmethod
    def backward(ctx, dh, dz):
        grads: Dict = {}
        view_bwd(ctx.model, ctx.c, dh, dz, grads)
        ctx.c = None
        return (None, None, None) + tuple(grads.get(p) for p in ctx.params)


def simclr_view_train(model, x: torch.Tensor, forced_idx=None):
    from .encoder.graph_encoder import GraphEncoder
    if not isinstance(model.encoder, GraphEncoder):
        raise NotImplementedError("the train path is implemented for GraphEncoder")
    params = tuple(model.parameters())
    return _ViewFn.apply(model, x, forced_idx, *params)


class PeakExtractFn(torch.autograd.Function):
    """Stand-alone peak extractor with weight gradients (module-level use)."""

    @staticmethod
    def forward(ctx, spec, w, b):
        ctx.save_for_backward(spec, w, b)
        return ops.peak_extract(spec.detach().contiguous(), w.detach(), b.detach())

    @staticmethod
    def backward(ctx, dout):
        spec, w, b = ctx.saved_tensors
        dw, db = ops.peak_extract_bwd(spec.contiguous(), w.detach(), b.detach(), dout.contiguous())
        return None, dw, db
