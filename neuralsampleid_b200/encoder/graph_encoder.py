"""B200 counterpart of the reference's encoder/graph_encoder.py: ``GraphEncoder`` (:91-214),
``FFN`` (:67-89), ``Downsample`` (:38-50) -- same constructor / forward signatures, attribute
names and state_dict keys, so ``SimCLR(cfg, encoder=GraphEncoder(cfg=cfg, in_channels=
cfg['n_filters'], k=3))`` (generate.py:68) and ``train.py:111`` work unchanged.

The forward chains node-major (B*N, C) activations through libgrafp_sm100a kernels: one layout
conversion at entry, none per layer.  Reference quirks reproduced on purpose:
every Grapher gets k=num_k[0], dilation 1 and no DropPath (Q1); conv is always 'mr' (Q2)."""
from __future__ import annotations

import torch
from torch import nn

from .. import ops
from .._prep import fold_conv_bn, make_linear, sig, tap3_weight
from .gcn_lib.torch_nn import act_layer
from .gcn_lib.torch_vertex import Grapher

SIZES = {"t": ([2, 2, 6, 2], [64, 128, 256, 512]), "s": ([2, 2, 6, 2], [80, 160, 400, 640]),
         "m": ([2, 2, 16, 2], [96, 192, 384, 768])}
SIZE_DEFAULT = ([2, 2, 18, 2], [128, 256, 512, 1024])


def _ffn_slab_rows(hidden: int) -> int:
    """Rows per FFN slab: GRAFP_FFN_SLAB_MB of hidden activations (multiple of 256); 0 = off (default).
    Measured on B200 at B = 4096: off 155 k seg/s, 96/64 MB 135 k, 32 MB 112 k, 16 MB 79 k -- the smaller
    launches lose more to persistent-kernel ramp/tail than L2 residency of the hidden tensor returns."""
    import os
    mb = float(os.environ.get("GRAFP_FFN_SLAB_MB", "0"))
    if mb <= 0:
        return 0
    return max(256, int(mb * (1 << 20) / (hidden * 4)) // 256 * 256)


class _Cached(nn.Module):
    def __init__(self):
        super().__init__()
        self._cache = {}

    def _memo(self, name, params, make):
        key = sig(*params)
        hit = self._cache.get(name)
        if hit is None or hit[0] != key:
            hit = (key, make())
            self._cache[name] = hit
        return hit[1]


def _bn_tensors(bn):
    return (bn.weight, bn.bias, bn.running_mean, bn.running_var)


class Downsample(_Cached):
    """Conv2d(3x3, stride 2, pad 1) + BN on an (N, 1) image == 3-tap stride-2 conv over nodes."""

    def __init__(self, in_dim=3, out_dim=768):
        super().__init__()
        self.conv = nn.Sequential(nn.Conv2d(in_dim, out_dim, 3, stride=2, padding=1),
                                  nn.BatchNorm2d(out_dim))

    def forward_nodes(self, x: torch.Tensor, B: int, N: int) -> torch.Tensor:
        if N % 2:
            raise NotImplementedError("Downsample needs an even node count (got %d)" % N)
        conv, bn = self.conv[0], self.conv[1]

        def make():
            _, s, t = fold_conv_bn(conv.weight[:, :, :1, 1], conv.bias, bn)
            return make_linear(tap3_weight(conv.weight), s, t)
        lin = self._memo("conv", (conv.weight, conv.bias) + _bn_tensors(bn), make)
        return ops.linear(x, lin, tap3_nodes=N // 2)

    def forward(self, x):
        B, C, N, W = x.shape
        if W != 1:
            raise NotImplementedError("Downsample is only defined on (B, C, N, 1) node images")
        out = self.forward_nodes(ops.nchw_to_nodes(x.reshape(B, C, N)), B, N)
        return ops.nodes_to_nchw(out, B, N // 2).unsqueeze(-1)


class FFN(_Cached):
    """x + BN(fc2(act(BN(fc1(x)))))."""

    def __init__(self, in_features, hidden_features=None, out_features=None, act="relu", drop_path=0.0):
        super().__init__()
        out_features = out_features if out_features is not None else in_features
        hidden_features = hidden_features if hidden_features is not None else in_features
        if drop_path > 0.0:
            raise NotImplementedError("DropPath is never instantiated by GraphEncoder (SURVEY Q1)")
        self.drop_path = nn.Identity()
        self.act = act_layer(act)
        self.fc1 = nn.Sequential(nn.Conv2d(in_features, hidden_features, 1, stride=1, bias=False, padding=0),
                                 nn.BatchNorm2d(hidden_features))
        self.fc2 = nn.Sequential(nn.Conv2d(hidden_features, out_features, 1, stride=1, bias=False, padding=0),
                                 nn.BatchNorm2d(out_features))

    def _folded(self, name):
        conv, bn = getattr(self, name)[0], getattr(self, name)[1]
        return self._memo(name, (conv.weight, conv.bias) + _bn_tensors(bn),
                          lambda: make_linear(*fold_conv_bn(conv.weight, conv.bias, bn)))

    def wants_split_input(self, channels: int) -> bool:
        """True when fc1 gains from a pre-split A operand written by the producer (measured on B200: the C >= 256
        stages, whose fc1 is MMA / shared-memory bound: 494 -> 433 us and 826 -> 673 us per layer; at C <= 128 the
        extra 268 MB write of the split copy costs more than the conversion stage it removes)."""
        import os
        if os.environ.get("GRAFP_NO_DUAL_OUT"):
            return False
        min_c = int(os.environ.get("GRAFP_DUAL_MIN_C", "256"))
        return channels >= min_c and ops.split_ok(self._folded("fc1"), channels) and _ffn_slab_rows(1) <= 0

    def forward_nodes(self, x: torch.Tensor, x_split=None) -> torch.Tensor:
        """``x_split``: the same tensor as an ops.SplitAct (from the producing GEMM's dual-output epilogue); fc1 then
        reads it instead of converting x in shared memory.  The shortcut always adds the fp32 x."""
        if self.training:
            raise RuntimeError("FFN.forward_nodes is the eval path; training goes through "
                               "neuralsampleid_b200.autograd")
        fc1, fc2 = self._folded("fc1"), self._folded("fc2")
        M, hid = x.shape[0], fc1.w.shape[0]
        if x_split is None and ops.ffn_fused_ok(fc1, fc2, x):
            # C <= 128: both GEMMs are HBM-bound on the 4C-wide hidden tensor -- one kernel keeps it on chip
            return ops.ffn_fused(x, fc1, fc2, self.act.name, self.act.neg_slope)
        slab = _ffn_slab_rows(hid)
        if slab <= 0 or M <= slab:
            # the hidden tensor feeds only fc2: on the bf16 tensor-core engines it travels as the
            # split-bf16 operand pair (same bytes, bit-identical operands, no conversion stage in fc2)
            split = ops.split_ok(fc1, x.shape[1]) and ops.split_ok(fc2, hid)
            a = x_split if (x_split is not None and split) else x
            h = ops.linear(a, fc1, self.act.name, self.act.neg_slope, out_split=split)
            return ops.linear(h, fc2, residual=x)
        # Optional (off by default, see _ffn_slab_rows): row slabs sized so the hidden activations of
        # one slab stay L2-resident between the two GEMMs.
        out = torch.empty_like(x)
        h = torch.empty((slab, hid), device=x.device, dtype=torch.float32)
        for r0 in range(0, M, slab):
            r1 = min(M, r0 + slab)
            xs = x[r0:r1]
            ops.linear(xs, fc1, self.act.name, self.act.neg_slope, out=h[:r1 - r0])
            ops.linear(h[:r1 - r0], fc2, residual=xs, out=out[r0:r1])
        return out

    def forward(self, x):
        B, C, N = x.shape[:3]
        out = self.forward_nodes(ops.nchw_to_nodes(x.reshape(B, C, -1)))
        return ops.nodes_to_nchw(out, B, out.shape[0] // B).reshape(x.shape)


class GraphEncoder(_Cached):
    def __init__(self, cfg, k=3, conv="mr", act="relu", norm="batch", bias=True, dropout=0.0, dilation=True,
                 epsilon=0.2, drop_path=0.1, size="t", emb_dims=1024, in_channels=3):
        super().__init__()
        self.blocks, self.channels = SIZES.get(size, SIZE_DEFAULT)
        self.k = int(k)
        self.act, self.norm, self.bias = act, norm, bias
        self.drop_path, self.emb_dims, self.epsilon = drop_path, emb_dims, epsilon
        self.dilation, self.dropout = dilation, dropout
        self.num_blocks = sum(self.blocks)
        self.conv = "mr"                                     # reference ignores the ctor arg (Q2)
        N = cfg["n_mels"] * cfg["n_frames"] // (cfg["patch_bins"] * cfg["patch_frames"])
        self.stem = nn.Sequential(nn.Conv2d(in_channels, self.channels[0], kernel_size=1, bias=False),
                                  nn.BatchNorm2d(self.channels[0]), nn.LeakyReLU(negative_slope=0.2))
        entries = []
        for i, nb in enumerate(self.blocks):
            if i > 0:
                entries.append(Downsample(self.channels[i - 1], self.channels[i]))
                N = N // 4                                   # sizes relative_pos only (Q7)
            for _ in range(nb):
                # every block: k = num_k[0], dilation = 1, drop_path = 0 (reference never advances idx, Q1)
                entries.append(nn.Sequential(
                    Grapher(self.channels[i], self.k, 1, self.conv, self.act, self.norm, self.bias, False,
                            epsilon, 1, n=N, drop_path=0.0, relative_pos=True),
                    FFN(in_features=self.channels[i], hidden_features=self.channels[i] * 4,
                        out_features=self.channels[i], act=act, drop_path=0.0)))
        self.backbone = nn.Sequential(*entries)
        self.proj = nn.Conv2d(self.channels[-1], self.emb_dims, 1, bias=True)

    def model_init(self):
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight)
                m.weight.requires_grad = True
                if m.bias is not None:
                    m.bias.data.zero_()
                    m.bias.requires_grad = True

    # ------------------------------------------------------------------------------------
    def _stem_lin(self):
        conv, bn = self.stem[0], self.stem[1]
        return self._memo("stem", (conv.weight,) + _bn_tensors(bn),
                          lambda: make_linear(*fold_conv_bn(conv.weight, None, bn)))

    def _stem_nodes(self, x_nodes, B=None, N=None):
        lin = self._stem_lin()
        cout, cin = lin.w.shape
        if B is not None and lin.groups == 1 and ops.stem_supported(cin, cout, N):
            return ops.stem(x_nodes, lin, "leakyrelu", self.stem[2].negative_slope, B, N)
        return ops.linear(x_nodes, lin, "leakyrelu", self.stem[2].negative_slope)

    def _proj(self, mean):
        def make():
            w, _, b = fold_conv_bn(self.proj.weight, self.proj.bias, None)
            return make_linear(w, None, b)
        return ops.linear(mean, self._memo("proj", (self.proj.weight, self.proj.bias), make))

    def forward(self, x, return_pre_proj=False, forced_idx=None, taps=None):
        """x: (B, in_channels, N) -> (B, emb_dims).

        ``return_pre_proj`` additionally returns the (B, C_last, N_last) node matrix (the call
        shape the reference's evaluation scripts use, encoder/dgl/graph_encoder.py:118-147).
        ``forced_idx`` / ``taps`` are parity-test hooks: a list of per-block int32 (B, N, k)
        neighbour lists to use instead of the computed graph (teacher forcing), and a list that
        receives per-block intermediates."""
        if self.training and torch.is_grad_enabled():
            from ..autograd import encoder_forward_train
            return encoder_forward_train(self, x, return_pre_proj)
        if self.training:
            # train mode under no_grad (e.g. a validation loss computed without .eval()): like the reference, a
            # batch-statistics BatchNorm forward that also moves the running statistics; the tape is discarded
            from ..autograd import encoder_train_fwd
            if x.dim() != 3:
                raise ValueError("expected (B, C, N) input, got %s" % (tuple(x.shape),))
            B, _, N = x.shape
            emb, nodes, tape = encoder_train_fwd(self, ops.nchw_to_nodes(x.detach()), B, N)
            if return_pre_proj:
                return ops.nodes_to_nchw(nodes, B, tape.N_out), emb
            return emb
        if x.dim() != 3:
            raise ValueError("expected (B, C, N) input, got %s" % (tuple(x.shape),))
        B, cin, N = x.shape
        lin = self._stem_lin()
        if lin.groups == 1 and ops.stem_supported(cin, lin.w.shape[0], N):
            # stem fused with the layout change: reads (B, C, N) directly, writes node-major
            h = ops.stem(x, lin, "leakyrelu", self.stem[2].negative_slope)
            return self.forward_nodes(None, B, N, return_pre_proj, forced_idx, taps, stem_out=h)
        return self.forward_nodes(ops.nchw_to_nodes(x), B, N, return_pre_proj, forced_idx, taps)

    def forward_nodes(self, x_nodes, B, N, return_pre_proj=False, forced_idx=None, taps=None, stem_out=None):
        """Eval forward from node-major (B*N, in_channels) features (what the peak extractor
        kernel emits), skipping the NCHW round trip."""
        h = stem_out if stem_out is not None else self._stem_nodes(x_nodes, B, N)
        blk = 0
        for entry in self.backbone:
            if isinstance(entry, Downsample):
                h = entry.forward_nodes(h, B, N)
                N //= 2
                continue
            t = {} if taps is not None else None
            fi = forced_idx[blk] if forced_idx is not None else None
            if t is not None:
                t["in"] = h
            if entry[1].wants_split_input(h.shape[1]):
                h, h_split = entry[0].forward_nodes(h, B, N, fi, t, want_split=True)
            else:
                h, h_split = entry[0].forward_nodes(h, B, N, fi, t), None
            if t is not None:
                t["grapher"] = h
            h = entry[1].forward_nodes(h, h_split)
            if t is not None:
                t["out"] = h
                taps.append(t)
            blk += 1
        emb = self._proj(ops.node_mean(h, B, N))
        if return_pre_proj:
            return ops.nodes_to_nchw(h, B, N), emb
        return emb
