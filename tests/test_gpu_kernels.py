"""Kernel-level parity: every C-ABI entry point against the CPU oracle / a plain fp32 torch
restatement on the same seeded inputs.  Bit-exact for index work and the max-relative aggregate;
stated fp32 tolerances for the floating-point kernels."""
import os

import numpy as np
import pytest
import torch

from oracle import grafp_oracle as O
from oracle import synth

pytestmark = pytest.mark.gpu

DEV = "cuda:0"


def _ops():
    from neuralsampleid_b200 import ops
    return ops


def _nodes(x_bcn1: torch.Tensor) -> torch.Tensor:
    """(B, C, N, 1) -> node-major (B*N, C) on the CPU (test-side layout helper)."""
    B, C, N = x_bcn1.shape[:3]
    return x_bcn1.reshape(B, C, N).transpose(1, 2).reshape(B * N, C).contiguous()


# ------------------------------------------------------------------------------------------
# layout
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,C,N", [(3, 8, 256), (2, 64, 96), (1, 5, 33)])
def test_layout_roundtrip(B, C, N):
    ops = _ops()
    x = synth.synth_normal((B, C, N), 3)
    nodes = ops.nchw_to_nodes(x.to(DEV))
    assert torch.equal(nodes.cpu(), x.transpose(1, 2).reshape(B * N, C))
    assert torch.equal(ops.nodes_to_nchw(nodes, B, N).cpu(), x)


# ------------------------------------------------------------------------------------------
# kNN
# ------------------------------------------------------------------------------------------
def _knn_check(x, k, d, tol=4e-6, normalize=True, engine=None):
    """x: (B, C, N, 1) cpu.  Compares the kernel's ordered neighbour lists with the oracle's,
    except on rows the oracle's own distances mark as ties at `tol`."""
    ops = _ops()
    B, C, N = x.shape[:3]
    if normalize:
        edge, dist = O.dilated_knn_graph(x, k, d)
        ref = edge[0]
    else:
        nn_idx, dist = O.dense_knn(x, k * d)
        ref = nn_idx[:, :, ::d]
    from neuralsampleid_b200 import _lib
    got, gd = ops.knn(_nodes(x).to(DEV), B, N, k, d, normalize=normalize, return_dist=True,
                      engine=None if engine is None else _lib.ENGINES[engine])
    got = got.cpu().long()
    assert got.shape == ref.shape
    assert int(got.min()) >= 0 and int(got.max()) < N
    tie = O.knn_tie_rows(dist, k * d, tol)
    diff = (got != ref).any(-1)
    assert not (diff & ~tie).any(), "%d off-tie rows differ" % int((diff & ~tie).sum())
    # distances of the selected ranks agree with the reference matrix to fp32 round-off
    want = torch.gather(dist, 2, ref)
    ok = ~diff
    # (the tensor-core engine's truncating fp32 accumulation biases distances by up to ~1e-5 at C=512)
    assert torch.allclose(gd.cpu()[ok], want[ok], rtol=0, atol=4e-6 if engine == 'simt' else 3e-5)
    return float(diff.float().mean()), float(tie.float().mean())


SWEEP = [(k, N) for k in (9, 16, 32) for N in (256, 512, 1024, 2048)]        # BASELINE configs[4], dilation on (d = 2)


@pytest.mark.parametrize("B,C,N,k,d", [
    (4, 64, 256, 3, 1), (4, 128, 128, 3, 1), (4, 256, 64, 3, 1), (4, 512, 32, 3, 1),   # size-'t' stages
    (4, 64, 256, 5, 1), (2, 64, 256, 9, 2), (2, 64, 96, 4, 3),                           # train k, dilation
    (2, 64, 512, 16, 2), (1, 64, 1024, 32, 2), (1, 64, 2048, 9, 3), (1, 64, 2048, 32, 2),  # stress sweep corners
    (3, 32, 40, 4, 1), (2, 8, 16, 16, 1), (1, 64, 200, 9, 2),                            # ragged sizes, k*d == N
])
def test_knn_matches_oracle(B, C, N, k, d):
    x = synth.synth_normal((B, C, N, 1), 100 + N + k)
    _knn_check(x, k, d, tol=4e-6)          # default engine: tensor cores where the shape allows


@pytest.mark.parametrize("k,N", SWEEP, ids=["k%d_n%d" % s for s in SWEEP])
def test_knn_stress_sweep_points_on_tensor_cores(k, N):
    """All 12 points of BASELINE configs[4] (k in {9,16,32} x N in {256..2048}, C = 64, dilation 2, and dilation 3
    where 3k <= 64): the tcgen05 engine takes every one of them (knn_big.cu) and the ordered lists equal the oracle's
    off documented ties."""
    ops = _ops()
    B = 3 if N <= 512 else 2
    for d in (2, 3):
        if k * d > 64:
            continue
        assert ops.knn_engine(B, N, 64, k, d) == "tcgen05"
        x = synth.synth_normal((B, 64, N, 1), 500 + N + k + d)
        _knn_check(x, k, d, tol=4e-6)


def test_knn_big_duplicates_zeros_and_other_widths():
    """knn_big.cu edge cases: exact duplicates and all-zero nodes (mass exact ties: the candidate lists grow to the
    whole graph), post-ReLU features, C in {32, 96, 128, 256}, k*d = 64 = the list limit, un-normalised input."""
    ops = _ops()
    x = torch.relu(synth.synth_normal((2, 64, 512, 1), 17))
    x[:, :, 5] = x[:, :, 9]
    x[:, :, 300] = x[:, :, 9]
    x[:, :, 17:40] = 0.0
    assert ops.knn_engine(2, 512, 64, 9, 2) == "tcgen05"
    _knn_check(x, 9, 2)
    _knn_check(torch.zeros((1, 64, 256, 1)), 16, 2)                       # every distance ties
    for C, N, k, d in ((96, 256, 17, 1), (32, 512, 8, 4), (128, 256, 32, 2), (256, 512, 5, 1), (64, 768, 6, 3)):
        assert ops.knn_engine(2, N, C, k, d) == "tcgen05"
        _knn_check(synth.synth_normal((2, C, N, 1), 600 + C + N), k, d)
    xn = torch.nn.functional.normalize(synth.synth_normal((2, 64, 512, 1), 18), dim=1)
    _knn_check(xn, 20, 1, normalize=False)
    # shapes neither tcgen05 kernel takes stay on the exact SIMT engine
    assert ops.knn_engine(1, 200, 64, 9, 2) == "simt" and ops.knn_engine(1, 512, 64, 40, 2) == "simt"
    assert ops.knn_engine(1, 512, 16, 9, 2) == "simt"
    _knn_check(synth.synth_normal((1, 16, 512, 1), 19), 9, 2)


@pytest.mark.parametrize("engine", ["simt", "3xtf32"])
@pytest.mark.parametrize("B,C,N,k,d", [
    (6, 64, 256, 3, 1), (6, 128, 128, 3, 1), (6, 256, 64, 3, 1), (6, 512, 32, 3, 1),   # size-'t' stages
    (5, 64, 256, 5, 1), (3, 64, 256, 8, 2), (3, 64, 128, 4, 3), (131, 64, 16, 4, 1), (7, 32, 64, 9, 1),
])
def test_knn_engines(B, C, N, k, d, engine):
    """Both engines against the oracle at the same documented-tie window (4e-6: adjacent reference distances within
    a few fp32 ulp at distance ~1)."""
    x = torch.relu(synth.synth_normal((B, C, N, 1), 300 + N + k)) + 0.05 * synth.synth_normal((B, C, N, 1), 301)
    _knn_check(x, k, d, tol=4e-6, engine=engine)


def test_knn_with_row_sumsq_from_gemm_epilogue():
    """The fc1 GEMM epilogue accumulates sum_c y^2 per node; the kNN then skips its own norm pass."""
    ops = _ops()
    from neuralsampleid_b200 import _prep
    B, N, C = 5, 256, 64
    a = synth.synth_normal((B * N, C), 1).to(DEV)
    w = (synth.synth_normal((C, C), 2) / 8.0).to(DEV)
    lin = _prep.make_linear(w, None, None)
    rs = torch.zeros(B * N, device=DEV)
    y = ops.linear(a, lin, row_sumsq=rs)
    assert torch.allclose(rs.cpu(), (y.cpu().double() ** 2).sum(1).float(), rtol=1e-5)
    got = ops.knn(y, B, N, 3, 1, row_sumsq=rs).cpu().long()
    x = y.cpu().view(B, N, C).transpose(1, 2).unsqueeze(-1).contiguous()
    edge, dist = O.dilated_knn_graph(x, 3, 1)
    tie = O.knn_tie_rows(dist, 3, 4e-6)
    assert not ((got != edge[0]).any(-1) & ~tie).any()
    rs2 = torch.zeros(B * N, device=DEV)                       # exact engine accumulates the same sums
    ops.linear(a, lin, row_sumsq=rs2, engine=1)
    assert torch.allclose(rs2, rs, rtol=1e-4)


def test_knn_post_relu_features_and_duplicates():
    # post-ReLU style features with exact duplicates and all-zero nodes: the reference has exact
    # ties here; every differing row must be a documented tie and indices must stay valid
    x = torch.relu(synth.synth_normal((2, 64, 128, 1), 7))
    x[:, :, 5] = x[:, :, 9]
    x[:, :, 17] = 0.0
    x[:, :, 18] = 0.0
    _knn_check(x, 5, 1)


def test_knn_unnormalised_entry_point():
    x = torch.nn.functional.normalize(synth.synth_normal((2, 32, 64, 1), 8), dim=1)
    _knn_check(x, 6, 1, normalize=False)


def test_knn_golden_dygraph_inputs(golden_dir):
    for fname, k, d in (("dygraph_k9_d2.npz", 9, 2), ("dygraph_k4_d3_n96.npz", 4, 3)):
        g = np.load(os.path.join(golden_dir, fname))
        x = torch.from_numpy(g["x"])
        ops = _ops()
        B, C, N = x.shape[:3]
        got = ops.knn(_nodes(x).to(DEV), B, N, k, d).cpu().long()
        _, dist = O.dilated_knn_graph(x, k, d)
        tie = O.knn_tie_rows(dist, k * d, 4e-6)
        diff = (got != torch.from_numpy(g["idx"].astype(np.int64))).any(-1)
        assert not (diff & ~tie).any()


def test_knn_rejects_bad_arguments():
    ops = _ops()
    from neuralsampleid_b200._lib import GrafpError
    x = torch.zeros((2 * 16, 8), device=DEV)
    with pytest.raises(GrafpError):
        ops.knn(x, 2, 16, 9, 2)            # k*d > N
    with pytest.raises(GrafpError):
        ops.knn(torch.zeros((16, 6), device=DEV), 1, 16, 3, 1)   # C % 4
    assert ops.knn(torch.zeros((0, 8), device=DEV), 0, 16, 3, 1).shape == (0, 16, 3)   # empty batch


# ------------------------------------------------------------------------------------------
# gather + max-relative (bit-exact)
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("B,C,N,k", [(5, 64, 256, 3), (300, 64, 256, 3), (4, 128, 128, 5), (3, 256, 64, 9),
                                     (2, 512, 32, 3), (2, 64, 2048, 16), (3, 32, 40, 4), (1, 8, 16, 16)])
def test_mr_aggregate_bit_exact(B, C, N, k):
    ops = _ops()
    x = synth.synth_normal((B, C, N, 1), 11 + B)
    rng = np.random.Generator(np.random.PCG64(5))
    idx = torch.from_numpy(rng.integers(0, N, size=(B, N, k), dtype=np.int64))
    idx[:, :, 0] = torch.arange(N)                      # self at rank 0, as the kNN produces
    center = torch.arange(N).view(1, N, 1).expand(B, N, k)
    want = O.max_relative(x, torch.stack((idx, center)))           # (B, C, N, 1)
    m, arg = ops.mr_aggregate(_nodes(x).to(DEV), idx.int().to(DEV), B, N, want_arg=True)
    assert torch.equal(m.cpu(), _nodes(want))
    # arg-max ranks reproduce m exactly
    xn = _nodes(x).view(B, N, C)
    a = arg.cpu().view(B, N, C).long()
    assert int(a.max()) < k
    nb = torch.gather(idx.unsqueeze(-1).expand(B, N, k, C), 2, a.unsqueeze(2)).squeeze(2)   # (B,N,C)
    vals = xn[torch.arange(B).view(B, 1, 1), nb, torch.arange(C).view(1, 1, C)] - xn
    assert torch.equal(vals.reshape(B * N, C), m.cpu())


@pytest.mark.parametrize("B,C,N,k", [(5, 64, 256, 3), (3, 128, 64, 9), (2, 32, 40, 4), (2, 64, 2048, 5)])
def test_nbr_reduce_modes(B, C, N, k):
    """grafp_nbr_reduce_fwd: plain max, (1+eps) x + sum, and max_k act(scale (x_j - x_i) + shift), staged and
    direct forms, contiguous and strided outputs."""
    ops = _ops()
    x = synth.synth_normal((B, C, N, 1), 61)
    rng = np.random.Generator(np.random.PCG64(7))
    idx = torch.from_numpy(rng.integers(0, N, size=(B, N, k), dtype=np.int64))
    xj = O.gather_nodes(x, idx)                                        # (B, C, N, k)
    nodes = _nodes(x).to(DEV)
    i32 = idx.int().to(DEV)
    got = ops.nbr_reduce(nodes, i32, B, N, ops.NBR_MAX).cpu()
    assert torch.equal(got, _nodes(xj.max(-1, keepdim=True)[0]))
    eps = torch.tensor([0.25], device=DEV)
    got = ops.nbr_reduce(nodes, i32, B, N, ops.NBR_SUM_SELF, eps=eps).cpu()
    want = _nodes(1.25 * x + xj.sum(-1, keepdim=True))
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-5)
    scale = synth.synth_uniform((C,), 62, -1.2, 1.2)
    shift = synth.synth_uniform((C,), 63, -0.5, 0.5)
    for act, fn in (("relu", torch.relu), ("gelu", torch.nn.functional.gelu), (None, lambda t: t)):
        wide = torch.full((B * N, 2 * C), -7.0, device=DEV)
        ops.nbr_reduce(nodes, i32, B, N, ops.NBR_EDGE_MAX, scale.to(DEV), shift.to(DEV), act, 0.0, out=wide[:, C:])
        e = fn((xj - x) * scale.view(1, C, 1, 1) + shift.view(1, C, 1, 1)).max(-1, keepdim=True)[0]
        assert torch.allclose(wide[:, C:].cpu(), _nodes(e), rtol=1e-5, atol=1e-5)
        assert bool((wide[:, :C] == -7.0).all())                      # the other column half is untouched


def test_index_select_matches_oracle():
    ops = _ops()
    x = synth.synth_normal((3, 16, 40, 1), 12)
    rng = np.random.Generator(np.random.PCG64(6))
    idx = torch.from_numpy(rng.integers(0, 40, size=(3, 40, 5), dtype=np.int64))
    got = ops.index_select(_nodes(x).to(DEV), idx.int().to(DEV), 3, 40)
    assert torch.equal(got.cpu(), O.gather_nodes(x, idx))


def test_mr_aggregate_bwd_matches_autograd():
    ops = _ops()
    B, C, N, k = 3, 32, 48, 4
    x = synth.synth_normal((B, C, N, 1), 13).requires_grad_(True)
    rng = np.random.Generator(np.random.PCG64(7))
    idx = torch.from_numpy(rng.integers(0, N, size=(B, N, k), dtype=np.int64))
    center = torch.arange(N).view(1, N, 1).expand(B, N, k)
    m = O.max_relative(x, torch.stack((idx, center)))
    gm = synth.synth_normal(tuple(m.shape), 14)
    m.backward(gm)
    xm, arg = ops.mr_aggregate(_nodes(x.detach()).to(DEV), idx.int().to(DEV), B, N, want_arg=True)
    dx = torch.zeros((B * N, C), device=DEV)
    ops.mr_aggregate_bwd(_nodes(gm).to(DEV), idx.int().to(DEV), arg, B, N, dx)
    assert torch.allclose(dx.cpu(), _nodes(x.grad), rtol=1e-5, atol=1e-6)


# ------------------------------------------------------------------------------------------
# GEMM family (every engine) against fp64-accumulated torch
# ------------------------------------------------------------------------------------------
def _gemm_ref(a1, w, scale, shift, act, slope, residual, a2, groups, tap3_nodes):
    a1, w = a1.double(), w.double()
    if tap3_nodes:
        cin = a1.shape[1]
        xb = a1.view(-1, 2 * tap3_nodes, cin)
        pad = torch.cat([torch.zeros_like(xb[:, :1]), xb], dim=1)            # node -1 = 0
        A = torch.cat([pad[:, 0:-1:2], pad[:, 1::2], pad[:, 2::2]], dim=2).reshape(-1, 3 * cin)
        y = A @ w.T
    else:
        n = w.shape[0] // groups
        outs = []
        for g in range(groups):
            k1 = a1.shape[1] // groups
            A = a1[:, g * k1:(g + 1) * k1]
            if a2 is not None:
                k2 = a2.shape[1] // groups
                A = torch.cat([A, a2.double()[:, g * k2:(g + 1) * k2]], dim=1)
            outs.append(A @ w[g * n:(g + 1) * n].T)
        y = torch.cat(outs, dim=1)
    if scale is not None:
        y = y * scale.double()
    if shift is not None:
        y = y + shift.double()
    if act == "relu":
        y = torch.relu(y)
    elif act == "leakyrelu":
        y = torch.nn.functional.leaky_relu(y, slope)
    elif act == "gelu":
        y = torch.nn.functional.gelu(y)
    elif act == "elu":
        y = torch.nn.functional.elu(y)
    if residual is not None:
        y = y + residual.double()
    return y


GEMM_CASES = [
    # M, k1, k2, n, groups, act, residual, tap3
    (1000, 8, 0, 64, 1, "leakyrelu", False, 0),        # stem
    (777, 64, 0, 64, 1, None, False, 0),               # Grapher.fc1
    (1024, 16, 16, 32, 4, "relu", False, 0),           # MRConv stage 1 (grouped, dual source)
    (512, 32, 32, 64, 4, "relu", False, 0),            # MRConv stage 2
    (300, 128, 0, 64, 1, None, True, 0),               # Grapher.fc2 + residual
    (640, 64, 0, 256, 1, "relu", False, 0),            # FFN.fc1
    (640, 256, 0, 64, 1, None, True, 0),               # FFN.fc2 + shortcut
    (256, 512, 0, 2048, 1, "gelu", False, 0),          # stage-4 FFN.fc1
    (130, 2048, 0, 512, 1, None, True, 0),             # stage-4 FFN.fc2, ragged M
    (3 * 64, 3 * 64, 0, 128, 1, None, False, 64),      # Downsample 64 -> 128, 128 -> 64 nodes
    (5 * 16, 3 * 256, 0, 512, 1, None, False, 16),     # Downsample 256 -> 512
    (33, 512, 0, 1024, 1, None, False, 0),             # proj
    (33, 1024, 0, 4096, 1, "elu", False, 0),           # projector.0
    (33, 4096, 0, 128, 1, None, False, 0),             # projector.2
    (50, 12, 0, 20, 1, "relu", True, 0),               # odd small shape (SIMT only)
]


@pytest.mark.parametrize("engine", ["simt", "3xtf32", "tf32", "bf16x3", "bf16", "f16x3", "auto"])
@pytest.mark.parametrize("case", GEMM_CASES, ids=lambda c: "m%d_k%d+%d_n%d_g%d_%s%s%s" % (
    c[0], c[1], c[2], c[3], c[4], c[5], "_res" if c[6] else "", "_tap3" if c[7] else ""))
def test_gemm_engines(case, engine):
    ops = _ops()
    from neuralsampleid_b200 import _lib, _prep
    M, k1, k2, n, groups, act, use_res, tap3 = case
    if tap3:
        cin = k1 // 3
        a1 = synth.synth_normal((2 * M, cin), 20)
    else:
        a1 = synth.synth_normal((M, groups * k1), 20)
    a2 = synth.synth_normal((M, groups * k2), 21) if k2 else None
    w = synth.synth_normal((groups * n, k1 + k2), 22) / float(np.sqrt(k1 + k2))
    scale = synth.synth_uniform((groups * n,), 23, 0.5, 1.5)
    shift = synth.synth_uniform((groups * n,), 24, -0.5, 0.5)
    res = synth.synth_normal((M, groups * n), 25) if use_res else None
    want = _gemm_ref(a1, w, scale, shift, act, 0.2, res, a2, groups, tap3)
    lin = _prep.make_linear(w.to(DEV), scale.to(DEV), shift.to(DEV), groups, dual=k2 > 0)
    eng = _lib.ENGINES[engine]
    # tcgen05 takes k % 32 == 0, n % 32 == 0; the Downsample form when Cin % 32 == 0 and the output
    # rows per graph tile 128 evenly
    tc_ok = lin.w_split is not None and n % 32 == 0 and \
        (not tap3 or ((k1 // 3) % 32 == 0 and (128 % tap3 == 0 or tap3 % 128 == 0)))
    if engine in ("3xtf32", "tf32", "bf16x3", "bf16", "f16x3") and not tc_ok:
        with pytest.raises(_lib.GrafpError):
            ops.linear(a1.to(DEV), lin, act, 0.2, res.to(DEV) if use_res else None,
                       a2.to(DEV) if k2 else None, tap3, eng)
        return
    got = ops.linear(a1.to(DEV), lin, act, 0.2, res.to(DEV) if use_res else None,
                     a2.to(DEV) if k2 else None, tap3, eng).cpu().double()
    assert got.shape == want.shape
    err = float((got - want).abs().max() / want.abs().max())
    # fp32 FFMA engine: 1e-5 of the output scale.  3xTF32: the tensor core's fp32 accumulator
    # truncates on every MMA, a bias that grows with the number of k-steps (measured 1.3e-5 at
    # k=2048, 3.0e-5 at k=4096) -> 5e-5.  Single-pass TF32: 2e-3.
    # bf16x3 (the default tensor-core engine): <= 3 * 2^-18 per product + the same accumulator bias -> 5e-5.
    # Plain bf16 operands: 1e-2.
    # f16x3 (the default tensor-core engine since ABI 4): ~3 * 2^-24 per product, so only the accumulator bias
    # is left: 4e-6 + 1.2e-8 per k element (measured: see profiles/r2*_parity.json).
    tol = {"tf32": 2e-3, "bf16": 1e-2, "simt": 1e-5}.get(engine, 5e-5)
    if engine in ("f16x3", "auto") and tc_ok:
        tol = 4e-6 + 1.2e-8 * (k1 + k2)
    assert err < tol, "engine %s rel err %.3g" % (engine, err)


@pytest.mark.parametrize("engine", ["bf16x3", "bf16", "f16x3", "auto"])
@pytest.mark.parametrize("M,k,hid,n,groups", [(640, 64, 256, 64, 1), (300, 256, 1024, 256, 1), (130, 512, 2048, 512, 1),
                                              (1000, 64, 128, 64, 4), (128 * 5 + 7, 128, 512, 128, 1)])
def test_gemm_split_bf16_activations_bit_exact(M, k, hid, n, groups, engine):
    """The split-bf16 activation format (include/grafp.h, ABI 2): a producer writes bf16 [hi ; lo] planes,
    the consumer reads them as its MMA operands.  Both must be BIT-identical to the fp32 route (the planes
    are exactly what the consumer's in-kernel conversion derives from the fp32 tensor)."""
    ops = _ops()
    from neuralsampleid_b200 import _lib, _prep
    eng = _lib.ENGINES[engine]
    a = synth.synth_normal((M, groups * k), 30).to(DEV)
    w1 = (synth.synth_normal((groups * hid, k), 31) / float(np.sqrt(k))).to(DEV)
    w2 = (synth.synth_normal((n, groups * hid), 32) / float(np.sqrt(groups * hid))).to(DEV)
    sc1 = synth.synth_uniform((groups * hid,), 33, 0.5, 1.5).to(DEV)
    sh1 = synth.synth_uniform((groups * hid,), 34, -0.5, 0.5).to(DEV)
    res = synth.synth_normal((M, n), 35).to(DEV)
    l1 = _prep.make_linear(w1, sc1, sh1, groups)
    l2 = _prep.make_linear(w2, None, None, 1)
    assert ops.split_ok(l1, groups * k) and ops.split_ok(l2, groups * hid)
    h32 = ops.linear(a, l1, "relu", 0.0, engine=eng)
    hs = ops.linear(a, l1, "relu", 0.0, engine=eng, out_split=True)
    planes = 1 if engine == "bf16" else 2                 # the 1-pass engine carries the hi plane only
    dt = torch.float16 if engine in ("f16x3", "auto") else torch.bfloat16    # fp16 planes on the f16x3 engine
    assert isinstance(hs, ops.SplitAct) and hs.t.shape == (planes, M, groups * hid) and hs.t.dtype == dt
    hi = h32.to(dt)
    lo = (h32 - hi.float()).to(dt)
    assert torch.equal(hs.t[0], hi) and (planes == 1 or torch.equal(hs.t[1], lo))
    y32 = ops.linear(h32, l2, None, 0.0, res, engine=eng)
    ys = ops.linear(hs, l2, None, 0.0, res, engine=eng)
    assert torch.equal(y32, ys)
    # dual output: one epilogue writes the fp32 tensor and its split copy (residual stream that is also the next
    # GEMM's operand); both bit-identical to the single-output launches
    yb32, ybs = ops.linear(hs, l2, None, 0.0, res, engine=eng, out_split="both")
    assert torch.equal(yb32, y32)
    yh = y32.to(dt)
    assert torch.equal(ybs.t[0], yh) and (planes == 1 or torch.equal(ybs.t[1], (y32 - yh.float()).to(dt)))


def test_gemm_pair_mma_bit_exact(tmp_path):
    """CTA-pair MMA (tcgen05.mma.cta_group::2, gemm_tc.cu kPair): the pre-split f16x3 GEMMs on 256-wide tiles must give the
    same bits whether a pair of CTAs computes a 256-row tile with one MMA or each CTA its own 128 rows.  The kernel choice
    is read from the environment once per process, so the same chain (single, split and dual output, ragged M, k from 256
    to 1024, odd and even numbers of row tiles) runs in two processes: pairs forced on every eligible shape, and pairs off."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = []
    for name, env in (("pair", {"GRAFP_TC_CLUSTER_MIN_WAVES": "0", "GRAFP_TC_PAIR_MIN_K": "32", "GRAFP_TC_PAIR": "1"}),
                      ("solo", {"GRAFP_TC_PAIR": "0", "GRAFP_TC_CLUSTER": "1"})):
        f = str(tmp_path / (name + ".pt"))
        e = dict(os.environ); e.update(env)
        r = subprocess.run([sys.executable, os.path.join(root, "tests", "pair_mma_case.py"), f], env=e, cwd=root,
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
        assert r.returncode == 0, r.stdout[-2000:]
        outs.append(torch.load(f))
    assert outs[0].pop("pair_launches") >= 12 and outs[1].pop("pair_launches") == 0      # the pair kernel really ran / did not
    assert outs[0].keys() == outs[1].keys() and len(outs[0]) == 16
    for k in outs[0]:
        assert torch.equal(outs[0][k], outs[1][k]), k


@pytest.mark.parametrize("B,N,C,k,groups", [(5, 256, 64, 3, 1), (9, 128, 128, 3, 4), (7, 64, 256, 5, 4), (11, 32, 512, 3, 4),
                                            (3, 64, 256, 9, 4)])
@pytest.mark.parametrize("out_split", [False, True])
def test_gemm_fused_max_relative_bit_exact(B, N, C, k, groups, out_split):
    """Fused MRConv2d (grafp_gemm_args.a2_gather_idx): the max-relative aggregate is computed by the GEMM's
    transform warps instead of being read -- BIT-identical to mr_aggregate + dual-source GEMM (ragged last
    tile included: B*N is not a multiple of 128 for the small-N cases)."""
    ops = _ops()
    from neuralsampleid_b200 import _prep
    x = synth.synth_normal((B * N, C), 70).to(DEV)
    rng = np.random.Generator(np.random.PCG64(9))
    idx = torch.from_numpy(rng.integers(0, N, size=(B, N, k), dtype=np.int64)).int().to(DEV)
    idx[:, :, 0] = torch.arange(N, device=DEV)
    n_out = 2 * C
    w = (synth.synth_normal((n_out, 2 * C // groups), 71) / float(np.sqrt(2 * C // groups))).to(DEV)
    scale = synth.synth_uniform((n_out,), 72, 0.5, 1.5).to(DEV)
    shift = synth.synth_uniform((n_out,), 73, -0.5, 0.5).to(DEV)
    lin = _prep.make_linear(w, scale, shift, groups, dual=True)
    os.environ["GRAFP_FUSED_MR"] = "1"
    ops._engine_override = "bf16x3"             # the fused gather exists on the bf16 engines only
    try:
        assert ops.fused_mr_ok(lin, C)
        m = ops.mr_aggregate(x, idx, B, N)
        want = ops.linear(x, lin, "relu", 0.0, a2=m, out_split=out_split)
        got = ops.linear(x, lin, "relu", 0.0, out_split=out_split, a2_gather=(idx, N))
    finally:
        del os.environ["GRAFP_FUSED_MR"]
        ops._engine_override = None
    if out_split:
        assert torch.equal(got.t, want.t)
    else:
        assert torch.equal(got, want)


@pytest.mark.parametrize("act", ["relu", "gelu"])
@pytest.mark.parametrize("M,C", [(128 * 5 + 7, 64), (4096, 64), (300, 128), (128 * 37, 128), (20000, 128)])
def test_ffn_fused_bit_exact(M, C, act):
    """grafp_ffn_fused_fwd (fc1 -> activation -> fc2 -> + shortcut in one kernel, hidden tile on chip) is BIT-identical
    to the two f16x3 GEMM launches it replaces (same operand values, same accumulation order), ragged last tile and
    several tiles per CTA included; and within the engine's tolerance of an fp64 restatement."""
    ops = _ops()
    from neuralsampleid_b200 import _prep
    hid = 4 * C
    x = synth.synth_normal((M, C), 40).to(DEV)
    w1 = (synth.synth_normal((hid, C), 41) / float(np.sqrt(C))).to(DEV)
    w2 = (synth.synth_normal((C, hid), 42) / float(np.sqrt(hid))).to(DEV)
    sc1, sh1 = synth.synth_uniform((hid,), 43, 0.5, 1.5).to(DEV), synth.synth_uniform((hid,), 44, -0.5, 0.5).to(DEV)
    sc2, sh2 = synth.synth_uniform((C,), 45, 0.5, 1.5).to(DEV), synth.synth_uniform((C,), 46, -0.5, 0.5).to(DEV)
    l1 = _prep.make_linear(w1, sc1, sh1, 1)
    l2 = _prep.make_linear(w2, sc2, sh2, 1)
    assert ops.ffn_fused_ok(l1, l2, x)
    h = ops.linear(x, l1, act, 0.0, out_split=True)
    want = ops.linear(h, l2, None, 0.0, x)
    got = ops.ffn_fused(x, l1, l2, act, 0.0)
    assert torch.equal(got, want), float((got - want).abs().max())
    assert torch.equal(ops.ffn_fused(x, l1, l2, act, 0.0), got)                    # deterministic
    ref = _gemm_ref(x.cpu(), w1.cpu(), sc1.cpu(), sh1.cpu(), act, 0.0, None, None, 1, 0)
    ref = _gemm_ref(ref.float(), w2.cpu(), sc2.cpu(), sh2.cpu(), None, 0.0, x.cpu(), None, 1, 0)
    assert float((got.cpu().double() - ref).abs().max() / ref.abs().max()) < 2e-5


@pytest.mark.parametrize("act", ["relu", "gelu"])
@pytest.mark.parametrize("M,C", [(128 * 5 + 7, 64), (4096, 64), (300, 128), (128 * 37, 128), (20000, 128), (128 * 600, 64)])
def test_mrconv_fc2_fused_bit_exact(M, C, act):
    """grafp_mrconv_fc2_fused_fwd (MRConv2d's groups = 4 conv over [x, m] -> activation -> fc2 -> + shortcut in one
    kernel, the 2C-wide MRConv output on chip) is BIT-identical to the two f16x3 GEMM launches it replaces (grouped
    dual-source GEMM at C = 128, block-diagonal densified GEMM at C = 64), ragged last tile and several tiles per
    CTA included; and within the engine's tolerance of an fp64 restatement of the grouped conv."""
    ops = _ops()
    from neuralsampleid_b200 import _prep
    x = synth.synth_normal((M, C), 50).to(DEV)
    m = synth.synth_normal((M, C), 51).abs().to(DEV)
    res = synth.synth_normal((M, C), 52).to(DEV)
    kg = C // 2                                      # columns per group: C/4 of x then C/4 of m (de-interleaved)
    w1 = (synth.synth_normal((2 * C, kg), 53) / float(np.sqrt(kg))).to(DEV)
    w2 = (synth.synth_normal((C, 2 * C), 54) / float(np.sqrt(2 * C))).to(DEV)
    sc1, sh1 = synth.synth_uniform((2 * C,), 55, 0.5, 1.5).to(DEV), synth.synth_uniform((2 * C,), 56, -0.5, 0.5).to(DEV)
    sc2, sh2 = synth.synth_uniform((C,), 57, 0.5, 1.5).to(DEV), synth.synth_uniform((C,), 58, -0.5, 0.5).to(DEV)
    l1 = _prep.make_linear(w1, sc1, sh1, 4, dual=True)
    l2 = _prep.make_linear(w2, sc2, sh2, 1)
    assert ops.mrconv_fc2_fused_ok(l1, l2, x, res)
    h = ops.linear(x, l1, act, 0.0, a2=m, out_split=True)
    want = ops.linear(h, l2, None, 0.0, res)
    got = ops.mrconv_fc2_fused(x, m, l1, act, 0.0, l2, res)
    assert torch.equal(got, want), float((got - want).abs().max())
    assert torch.equal(ops.mrconv_fc2_fused(x, m, l1, act, 0.0, l2, res), got)     # deterministic
    # fp64: group g reads x[:, g C/4 : (g+1) C/4] and m[:, the same columns]
    xd, md, q = x.cpu().double(), m.cpu().double(), C // 4
    hid = torch.cat([torch.cat([xd[:, g * q:(g + 1) * q], md[:, g * q:(g + 1) * q]], 1)
                     @ w1.cpu().double()[g * (C // 2):(g + 1) * (C // 2)].T for g in range(4)], 1)
    hid = hid * sc1.cpu().double() + sh1.cpu().double()
    hid = torch.relu(hid) if act == "relu" else torch.nn.functional.gelu(hid)
    ref = (hid @ w2.cpu().double().T) * sc2.cpu().double() + sh2.cpu().double() + res.cpu().double()
    assert float((got.cpu().double() - ref).abs().max() / ref.abs().max()) < 2e-5


def test_gemm_split_bf16_needs_bf16_engine():
    ops = _ops()
    from neuralsampleid_b200 import _lib, _prep
    a = synth.synth_normal((256, 64), 30).to(DEV)
    w = (synth.synth_normal((64, 64), 31) / 8.0).to(DEV)
    lin = _prep.make_linear(w, None, None, 1)
    for name in ("simt", "3xtf32", "tf32"):
        with pytest.raises(_lib.GrafpError):
            ops.linear(a, lin, engine=_lib.ENGINES[name], out_split=True)
    hs = ops.linear(a, lin, out_split=True)
    with pytest.raises(_lib.GrafpError):
        ops.linear(hs, lin, engine=_lib.ENGINES["simt"])


@pytest.mark.parametrize("B,cin,N,cout", [(5, 8, 256, 64), (3, 4, 96, 32), (2, 16, 128, 128), (300, 8, 256, 64)])
def test_stem_kernel_matches_gemm_route(B, cin, N, cout):
    """grafp_stem_fwd (stem fused with the NCHW -> node-major layout change) against the generic route
    (transpose + fp32 SIMT GEMM) and the fp64 reference; both input layouts."""
    ops = _ops()
    from neuralsampleid_b200 import _lib, _prep
    x = synth.synth_uniform((B, cin, N), 40)
    w = synth.synth_normal((cout, cin), 41) / float(np.sqrt(cin))
    scale = synth.synth_uniform((cout,), 42, 0.5, 1.5)
    shift = synth.synth_uniform((cout,), 43, -0.5, 0.5)
    lin = _prep.make_linear(w.to(DEV), scale.to(DEV), shift.to(DEV), 1)
    assert ops.stem_supported(cin, cout, N)
    nodes = ops.nchw_to_nodes(x.to(DEV))
    via_gemm = ops.linear(nodes, lin, "leakyrelu", 0.2, engine=_lib.ENGINE_SIMT)
    got_nchw = ops.stem(x.to(DEV), lin, "leakyrelu", 0.2)
    got_nodes = ops.stem(nodes, lin, "leakyrelu", 0.2, B, N)
    assert torch.equal(got_nchw, got_nodes)
    want = _gemm_ref(nodes.cpu(), w, scale, shift, "leakyrelu", 0.2, None, None, 1, 0)
    for got in (got_nchw, via_gemm):
        err = float((got.cpu().double() - want).abs().max() / want.abs().max())
        assert err < 1e-5, err
    assert not ops.stem_supported(12, 64, 256) and not ops.stem_supported(8, 20, 256)
    with pytest.raises(_lib.GrafpError):
        ops.stem(torch.zeros((2, 12, 64), device=DEV), _prep.make_linear(torch.zeros((64, 12), device=DEV), None, None, 1))


def test_gemm_empty_and_errors():
    ops = _ops()
    from neuralsampleid_b200._lib import GrafpError
    w = torch.zeros((16, 8), device=DEV)
    assert ops.gemm(torch.zeros((0, 8), device=DEV), w).shape == (0, 16)
    with pytest.raises(GrafpError):
        ops.gemm(torch.zeros((4, 6), device=DEV), torch.zeros((16, 6), device=DEV))      # k % 4
    with pytest.raises(NotImplementedError):
        ops.gemm(torch.zeros((4, 8), device=DEV), w, act="prelu")
    with pytest.raises(GrafpError):
        ops.gemm(torch.zeros((4, 8)), w)                                                  # CPU tensor


# ------------------------------------------------------------------------------------------
# small stages
# ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("P,Nq,Nk,H,Dh", [(3, 32, 32, 4, 128), (2, 20, 20, 2, 64), (5, 7, 13, 1, 32), (1, 64, 64, 4, 128)])
def test_mha_pool_matches_torch(P, Nq, Nk, H, Dh):
    """grafp_mha_pool_fwd: per-head softmax attention + mean over the queries, k / v as column halves of one
    fused projection output (strided views), against fp64 torch."""
    ops = _ops()
    E = H * Dh
    q = synth.synth_normal((P * Nq, E), 80)
    kv = synth.synth_normal((P * Nk, 2 * E), 81)
    got = ops.mha_pool(q.to(DEV), kv.to(DEV)[:, :E], kv.to(DEV)[:, E:], P, Nq, Nk, H).cpu().double()
    qq = q.double().view(P, Nq, H, Dh).transpose(1, 2)
    kk = kv[:, :E].double().reshape(P, Nk, H, Dh).transpose(1, 2)
    vv = kv[:, E:].double().reshape(P, Nk, H, Dh).transpose(1, 2)
    att = torch.softmax(qq @ kk.transpose(-1, -2) / float(Dh) ** 0.5, dim=-1)
    want = (att @ vv).mean(dim=2).reshape(P, E)                       # (P, H, Dh) -> (P, E)
    assert got.shape == want.shape and float((got - want).abs().max()) < 1e-5
    # positional add fused with the layout change
    x = synth.synth_normal((P, 24, Nq), 82)
    pos = synth.synth_normal((Nq, 24), 83)
    nodes = ops.nchw_to_nodes_add(x.to(DEV), pos.to(DEV)).cpu()
    assert torch.equal(nodes, (x.transpose(1, 2) + pos).reshape(P * Nq, 24))
    assert torch.equal(ops.nchw_to_nodes_add(x.to(DEV), None).cpu(), x.transpose(1, 2).reshape(P * Nq, 24))


def test_gemm_sigmoid_epilogue_simt_only():
    ops = _ops()
    from neuralsampleid_b200 import _lib, _prep
    a = synth.synth_normal((70, 128), 84)
    w = synth.synth_normal((1, 128), 85) / 11.0
    b = synth.synth_uniform((1,), 86, -0.5, 0.5)
    lin = _prep.make_linear(w.to(DEV), None, b.to(DEV))
    got = ops.linear(a.to(DEV), lin, "sigmoid").cpu()
    want = torch.sigmoid(a.double() @ w.double().T + b.double()).float()
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)
    w64 = (synth.synth_normal((64, 128), 87) / 11.0).to(DEV)
    with pytest.raises(_lib.GrafpError):
        ops.linear(a.to(DEV), _prep.make_linear(w64, None, None), "sigmoid", engine=_lib.ENGINES["bf16x3"])


def test_node_mean_and_l2_normalize():
    ops = _ops()
    x = synth.synth_normal((6 * 32, 512), 30)
    got = ops.node_mean(x.to(DEV), 6, 32).cpu()
    assert torch.allclose(got, x.view(6, 32, 512).mean(1), rtol=1e-6, atol=1e-6)
    z = synth.synth_normal((9, 128), 31)
    z[3] = 0.0
    got = ops.l2_normalize_rows(z.to(DEV), 1e-10).cpu()
    assert torch.allclose(got, torch.nn.functional.normalize(z, p=2, eps=1e-10), rtol=1e-6, atol=1e-7)


def test_peak_extractor_matches_oracle():
    ops = _ops()
    cfg = dict(n_filters=8, patch_bins=4, patch_frames=8)
    spec = [("peak_extractor.convs.0.weight", (8, 3, 4, 8), "w"), ("peak_extractor.convs.0.bias", (8,), "b")]
    sd = synth.synth_state(spec, 40)
    s = synth.synth_normal((5, 64, 128), 41)
    want = O.peak_extractor(sd, s)                                   # (B, 8, 256)
    got = ops.peak_extract(s.to(DEV), sd["peak_extractor.convs.0.weight"].to(DEV),
                           sd["peak_extractor.convs.0.bias"].to(DEV)).cpu()
    want_nodes = want.transpose(1, 2).reshape(5 * 256, 8)
    assert torch.allclose(got, want_nodes, rtol=1e-5, atol=2e-6)


@pytest.mark.parametrize("B,F,shape,patch", [(5, 8, (64, 128), (4, 8)), (700, 8, (64, 128), (4, 8)),
                                             (9, 16, (32, 64), (8, 4)), (3, 4, (64, 128), (2, 16))])
def test_peak_extractor_node_kernel_is_bit_identical(monkeypatch, B, F, shape, patch):
    """The one-thread-per-node peak extractor (persistent CTAs, ramp partial sums tabulated once, spectrogram normalised
    once per element) keeps the accumulation order of the simple one-CTA-per-segment kernel: bit-identical, more
    segments than CTAs, other patch shapes, NaN and constant segments included."""
    ops = _ops()
    s = synth.synth_normal((B,) + shape, 45).to(DEV)
    s[B // 2, 3, 5] = float("nan")                       # torch.min / max propagate NaN: the whole segment is NaN
    s[B - 1] = 0.25                                      # max == min: 0 / 0
    w = synth.synth_normal((F, 3) + patch, 46).to(DEV)
    b = synth.synth_normal((F,), 47).to(DEV)
    got = ops.peak_extract(s, w, b)
    monkeypatch.setenv("GRAFP_PEAK_SIMPLE", "1")
    want = ops.peak_extract(s, w, b)
    assert got.shape == want.shape
    assert torch.equal(torch.isnan(got), torch.isnan(want))
    assert torch.equal(torch.nan_to_num(got, nan=7.0), torch.nan_to_num(want, nan=7.0))
    assert bool(torch.isnan(got).any()) and bool((~torch.isnan(got)).any())


# ------------------------------------------------------------------------------------------
# NT-Xent
# ------------------------------------------------------------------------------------------
def test_ntxent_matches_golden(golden_dir):
    from neuralsampleid_b200.simclr.ntxent import ntxent_loss
    g = np.load(os.path.join(golden_dir, "ntxent_b16.npz"))
    zi = torch.from_numpy(g["z_i"]).to(DEV).requires_grad_(True)
    zj = torch.from_numpy(g["z_j"]).to(DEV).requires_grad_(True)
    loss = ntxent_loss(zi, zj, {"tau": float(g["tau"])})
    loss.backward()
    np.testing.assert_allclose(loss.item(), g["loss"], rtol=1e-5)
    np.testing.assert_allclose(zi.grad.cpu().numpy(), g["g_i"], rtol=1e-3, atol=1e-6)
    np.testing.assert_allclose(zj.grad.cpu().numpy(), g["g_j"], rtol=1e-3, atol=1e-6)


@pytest.mark.parametrize("B", [1, 7, 256, 1000])
def test_ntxent_matches_oracle_sizes(B):
    from neuralsampleid_b200.simclr.ntxent import ntxent_loss
    z_i = torch.nn.functional.normalize(synth.synth_normal((B, 128), 50), dim=1)
    z_j = torch.nn.functional.normalize(z_i + 0.3 * synth.synth_normal((B, 128), 51), dim=1)
    a = z_i.clone().requires_grad_(True)
    b = z_j.clone().requires_grad_(True)
    want = O.ntxent(a, b, 0.05)
    want.backward()
    zi = z_i.to(DEV).requires_grad_(True)
    zj = z_j.to(DEV).requires_grad_(True)
    loss = ntxent_loss(zi, zj, {"tau": 0.05})
    (2.0 * loss).backward()
    np.testing.assert_allclose(loss.item(), want.item(), rtol=2e-5)
    scale = float(a.grad.abs().max())
    assert float((zi.grad.cpu() - 2 * a.grad).abs().max()) < 2e-4 * 2 * scale + 1e-7
    assert float((zj.grad.cpu() - 2 * b.grad).abs().max()) < 2e-4 * 2 * scale + 1e-7


def test_ntxent_sharded_rows_equal_global():
    """The data-parallel form: two ranks' row ranges sum to the global loss and the local slices of
    the gradient concatenate to the global gradient (no collective needed to check the kernel)."""
    ops = _ops()
    B = 24
    z = torch.nn.functional.normalize(synth.synth_normal((2 * B, 128), 52), dim=1).to(DEV)
    loss, lse = ops.ntxent_fwd(z, 0.05)
    l0, lse0 = ops.ntxent_fwd(z, 0.05, 0, 20)
    l1, lse1 = ops.ntxent_fwd(z, 0.05, 20, 2 * B - 20)
    assert torch.allclose(l0 + l1, loss, rtol=1e-6)
    assert torch.equal(torch.cat([lse0, lse1]), lse)
    one = torch.ones(1, device=DEV)
    dz = ops.ntxent_bwd(z, lse, 0.05, one)
    d0 = ops.ntxent_bwd(z, lse, 0.05, one, 0, 20)
    d1 = ops.ntxent_bwd(z, lse, 0.05, one, 20, 2 * B - 20)
    assert torch.equal(torch.cat([d0, d1]), dz)


# ------------------------------------------------------------------------------------------
# exact fingerprint search
# ------------------------------------------------------------------------------------------
def _search_check(index, db, q, k, tol=2e-4):
    from oracle.flat_l2 import flat_l2_search
    D, I = index.search(q, k)
    Dw, Iw = flat_l2_search(db, q, k)
    assert D.shape == (q.shape[0], k) and I.shape == (q.shape[0], k) and I.dtype == np.int64
    fin = np.isfinite(Dw)
    assert np.array_equal(np.isfinite(D), fin) and bool((I[~fin] == -1).all())
    assert np.allclose(D[fin], Dw[fin], rtol=0, atol=tol)
    Df = np.where(fin, D, np.finfo(np.float32).max)
    assert bool((np.diff(Df, axis=1) >= 0).all())                                  # ascending, padding last
    # ids agree except where the oracle's own distances are within the engine's error of each other
    bad = (I != Iw) & fin
    for r, c in zip(*np.nonzero(bad)):
        alt = np.abs(Dw[r] - Dw[r, c]) < 2 * tol
        assert I[r, c] in Iw[r][alt] or abs(float(D[r, c]) - float(Dw[r, min(c + 1, k - 1)])) < 2 * tol, (r, c)
    return D, I


def test_flat_l2_index_matches_exact_search():
    from neuralsampleid_b200.db import FlatL2Index
    rng = np.random.Generator(np.random.PCG64(11))
    n, d, k = 150000, 128, 20                                  # three database chunks, the last one ragged
    db = rng.standard_normal((n, d)).astype(np.float32)
    db /= np.linalg.norm(db, axis=1, keepdims=True)
    index = FlatL2Index(d, DEV)
    index.add(db[:70000])
    index.add(db[70000:])
    assert index.ntotal == n
    qi = rng.integers(0, n, size=300)
    q = db[qi] + 0.05 * rng.standard_normal((300, d)).astype(np.float32)            # distorted copies of items
    q = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    D, I = _search_check(index, db, q, k)
    assert float((I[:, 0] == qi).mean()) > 0.99                                      # the source item ranks first
    _search_check(index, db, q[:3], 5)                                              # few queries: 16 column splits
    _search_check(index, db, (0.3 * rng.standard_normal((40, d))).astype(np.float32), 32, tol=1e-3)   # un-normalised, k = 32


def test_flat_l2_index_matches_the_independent_search_fixture(golden_dir):
    """The GPU index over a database in the reference's on-disk layout against the committed fixture of an independent
    exact float64 search (tests/golden/make_search_golden.py; FAISS was never run: it is absent from the image and
    from /root/reference): identical ids -- the duplicate pair resolves to the lower id -- and distances to 1e-5."""
    from neuralsampleid_b200.db import FlatL2Index, load_fingerprints
    g = np.load(os.path.join(golden_dir, "search_expected.npz"))
    emb, _ = load_fingerprints(os.path.join(golden_dir, "search_db"), "ref_db")
    index = FlatL2Index(128, DEV)
    index.add(np.asarray(emb))
    k = int(g["k"])
    D, I = index.search(g["q"], k)
    Dw, Iw = g["D"], g["I"]
    assert np.allclose(D, Dw, rtol=0, atol=1e-5)
    bad = I != Iw
    for r, c in zip(*np.nonzero(bad)):                     # only where the fixture's own distances are within 4e-6
        assert abs(Dw[r, c] - Dw[r, min(c + 1, k - 1)]) < 4e-6 or abs(Dw[r, c] - Dw[r, max(c - 1, 0)]) < 4e-6, (r, c)
    assert I[0, 0] == 33 and I[0, 1] == 700                # the exact duplicates, lower id first
    assert float(bad.mean()) < 0.01


def test_flat_l2_index_small_and_duplicates():
    from neuralsampleid_b200.db import FlatL2Index
    rng = np.random.Generator(np.random.PCG64(12))
    db = rng.standard_normal((7, 128)).astype(np.float32)
    db /= np.linalg.norm(db, axis=1, keepdims=True)
    index = FlatL2Index(128, DEV)
    index.add(db)
    D, I = _search_check(index, db, db[:4], 10)                 # k > ntotal: -1 / inf padding
    assert bool((I[:, 7:] == -1).all()) and bool(np.isinf(D[:, 7:]).all())
    assert bool((I[:, 0] == np.arange(4)).all()) and float(np.abs(D[:, 0]).max()) < 1e-4
    dup = np.repeat(db[:1], 50, axis=0)                          # exact duplicates: ties -> lower index first
    index2 = FlatL2Index(128, DEV)
    index2.add(np.concatenate([dup, db], axis=0))
    D2, I2 = index2.search(db[:1], 8)
    assert sorted(I2[0].tolist()) == list(range(8)) or bool((I2[0] < 51).all())
    assert float(np.abs(D2).max()) < 1e-4
    assert FlatL2Index(128, DEV).search(db[:2], 3)[1].tolist() == [[-1] * 3] * 2   # empty index


def test_song_level_ranking_matches_oracle():
    """eval.py:300-336: segment-level search -> candidate sequences -> per-file score histogram."""
    from neuralsampleid_b200.db import FlatL2Index, song_level_ranking
    from oracle import flat_l2
    rng = np.random.Generator(np.random.PCG64(21))
    n_files, per, d = 40, 60, 128
    db = rng.standard_normal((n_files * per, d)).astype(np.float32)
    db /= np.linalg.norm(db, axis=1, keepdims=True)
    file_of = np.repeat(np.arange(n_files), per)
    index = FlatL2Index(d, DEV)
    index.add(db)
    for start, sl in ((7 * per + 11, 9), (n_files * per - 5, 5), (3, 19)):          # incl. a sequence running off the end
        q = db[start:start + sl] + 0.08 * rng.standard_normal((min(sl, db.shape[0] - start), d)).astype(np.float32)
        q = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
        files, scores = song_level_ranking(index, q, 20, file_of, query_file=-1, first_valid=per)   # file 0 = dummy DB
        wf, ws = flat_l2.song_level_ranking(db, q, 20, file_of, -1, per)
        # the winning file and its score (a k_probe-boundary near-tie between the engines can move a weak candidate in
        # or out, never the true match)
        assert files[0] == wf[0] and (start < per or files[0] == file_of[start])
        assert abs(float(scores[0]) - float(ws[0])) < 1e-3
        # the scoring kernel itself, on the oracle's own candidate list: deterministic, tight
        _, Iw = flat_l2.flat_l2_search(db, q, 20)
        cand = Iw[np.where(Iw >= 0)].flatten()
        got = index.sequence_scores(q, cand).cpu().numpy()
        want = np.array([np.mean(np.sum(q[:db[c:c + q.shape[0]].shape[0]].astype(np.float64) *
                                        db[c:c + q.shape[0]].astype(np.float64), axis=1)) for c in cand])
        assert np.allclose(got, want, rtol=0, atol=2e-6)
    sc = index.sequence_scores(db[100:104], np.array([100, -1, db.shape[0] - 2, db.shape[0] + 5]))
    assert abs(float(sc[0]) - 1.0) < 1e-5 and float(sc[1]) == 0.0 and float(sc[3]) == 0.0
    want2 = float(np.mean(np.sum(db[100:102].astype(np.float64) * db[-2:].astype(np.float64), axis=1)))
    assert abs(float(sc[2]) - want2) < 1e-5
