"""Contrastive train step (train.py:48-83 semantics) against the CPU oracle's autograd and the golden
vectors minted from the reference: train-mode BatchNorm forward, hand-written backward, NT-Xent,
clip_grad_norm_(1.0) + Adam.

Tolerances: loss 1e-3 relative (north_star).  Gradients are compared (a) per layer against torch
autograd on identical inputs (1e-4 of the gradient scale) and (b) end to end by direction (cosine
> 0.999) and norm (2e-2): train-mode near-tie neighbour flips legitimately perturb a few rows."""
import os

import numpy as np
import pytest
import torch

from oracle import grafp_oracle as O
from oracle import synth

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
CFG = dict(n_mels=64, n_frames=128, patch_bins=4, patch_frames=8, n_filters=8, tau=0.05,
           d=128, h=1024, u=32, dim=2048, arch="grafp", bsz_train=256, lr=8.0e-5)


def _model(k):
    from neuralsampleid_b200.encoder.graph_encoder import GraphEncoder
    from neuralsampleid_b200.simclr.simclr import SimCLR
    sd = synth.synth_state(synth.simclr_state_spec(CFG, "t"), 1236)
    model = SimCLR(CFG, encoder=GraphEncoder(cfg=CFG, in_channels=CFG["n_filters"], k=k))
    model.load_state_dict(sd)
    return model.to(DEV), sd


def _inputs(B):
    s_i = synth.synth_normal((B, 64, 128), 21)
    s_j = s_i + 0.1 * synth.synth_normal((B, 64, 128), 22)
    return s_i, s_j


@pytest.mark.parametrize("act,use_bn,use_res,groups,dual", [
    ("relu", True, False, 1, False), (None, True, True, 1, False), ("relu", True, False, 4, True),
    ("elu", False, False, 1, False), ("leakyrelu", True, False, 1, False), ("gelu", True, True, 1, False)])
def test_layer_forward_backward_matches_autograd(act, use_bn, use_res, groups, dual):
    from neuralsampleid_b200 import autograd as A
    M, k, n = 600, 64, 96
    kk = k * (2 if dual else 1)
    a1 = synth.synth_normal((M, groups * k), 1)
    a2 = synth.synth_normal((M, groups * k), 2) if dual else None
    w = (synth.synth_normal((groups * n, kk), 3) / np.sqrt(kk)).requires_grad_(True)
    bias = synth.synth_uniform((groups * n,), 4, -0.1, 0.1).requires_grad_(True)
    res = synth.synth_normal((M, groups * n), 5) if use_res else None
    gout = synth.synth_normal((M, groups * n), 6)
    bn = torch.nn.BatchNorm2d(groups * n)
    with torch.no_grad():
        bn.weight.copy_(synth.synth_uniform((groups * n,), 7, 0.6, 1.2))
        bn.bias.copy_(synth.synth_uniform((groups * n,), 8, -0.2, 0.2))
    # torch reference
    x1 = a1.clone().requires_grad_(True)
    x2 = a2.clone().requires_grad_(True) if dual else None
    outs = []
    for g in range(groups):
        A_g = x1[:, g * k:(g + 1) * k]
        if dual:
            A_g = torch.cat([A_g, x2[:, g * k:(g + 1) * k]], dim=1)
        outs.append(A_g @ w[g * n:(g + 1) * n].T)
    y = torch.cat(outs, dim=1) + bias
    bn_ref = torch.nn.BatchNorm2d(groups * n)
    bn_ref.load_state_dict(bn.state_dict())
    if use_bn:
        y = bn_ref.train()(y.t().reshape(1, groups * n, M, 1)).reshape(groups * n, M).t()
    y = {"relu": torch.relu, "elu": torch.nn.functional.elu, "gelu": torch.nn.functional.gelu,
         "leakyrelu": lambda t: torch.nn.functional.leaky_relu(t, 0.2), None: lambda t: t}[act](y)
    if use_res:
        y = y + res
    y.backward(gout)
    # kernels
    wp = torch.nn.Parameter(w.detach().to(DEV))
    bp = torch.nn.Parameter(bias.detach().to(DEV))
    bn = bn.to(DEV).train()
    tape = []
    out = A.layer_fwd(tape, a1.to(DEV), wp, wp.detach(), "dense", bp, bn if use_bn else None, act, 0.2,
                      res.to(DEV) if use_res else None, a2.to(DEV) if dual else None, groups)
    assert torch.allclose(out.cpu(), y.detach(), rtol=1e-4, atol=1e-4)
    if use_bn:
        assert torch.allclose(bn.running_mean.cpu(), bn_ref.running_mean, rtol=1e-4, atol=1e-5)
        assert torch.allclose(bn.running_var.cpu(), bn_ref.running_var, rtol=1e-4, atol=1e-5)
    grads = {}
    da1, da2 = A.layer_bwd(tape[0], gout.to(DEV), grads)

    def close(a, b, name):
        scale = float(b.abs().max()) + 1e-12
        assert float((a.cpu() - b).abs().max()) < 2e-4 * scale + 1e-6, name
    close(da1, x1.grad, "da1")
    if dual:
        close(da2, x2.grad, "da2")
    close(grads[wp], w.grad, "dw")
    if use_bn:
        close(grads[bn.weight], bn_ref.weight.grad, "dgamma")
        close(grads[bn.bias], bn_ref.bias.grad, "dbeta")
        assert float(grads[bp].abs().max()) == 0.0
    else:
        close(grads[bp], bias.grad, "dbias")


def _oracle_train(sd, s_i, s_j, k, names):
    params = {n: (t.clone().requires_grad_(True) if n in names else t.clone()) for n, t in sd.items()}
    stats = {}
    h_i, h_j, z_i, z_j = O.simclr_forward(params, s_i, s_j, k=k, training=True, stats=stats)
    loss = O.ntxent(z_i, z_j, CFG["tau"])
    loss.backward()
    return loss.detach(), z_i.detach(), z_j.detach(), {n: params[n].grad for n in names}, stats


def test_train_step_autograd_path_matches_oracle_and_golden(golden_dir):
    from neuralsampleid_b200.simclr.ntxent import ntxent_loss
    g = np.load(os.path.join(golden_dir, "simclr_train_b8.npz"))
    names = [str(n) for n in g["grad_names"]]
    model, sd = _model(5)
    model.train()
    s_i, s_j = _inputs(8)
    loss_o, zi_o, zj_o, grads_o, stats_o = _oracle_train(sd, s_i, s_j, 5, names)
    h_i, h_j, z_i, z_j = model(s_i.to(DEV), s_j.to(DEV))
    loss = ntxent_loss(z_i, z_j, CFG)
    loss.backward()
    # loss / embeddings: vs the oracle on this machine and vs the reference's golden values
    assert abs(loss.item() - loss_o.item()) < 1e-3 * abs(loss_o.item())
    assert abs(loss.item() - float(g["loss"])) < 1e-3 * abs(float(g["loss"]))
    rel = (z_i.detach().cpu() - zi_o).norm(dim=1) / zi_o.norm(dim=1)
    assert float(rel.median()) < 1e-3 and float(rel.max()) < 5e-2, rel
    # gradients: direction and size of the full gradient, and per-parameter norms
    named = dict(model.named_parameters())
    got = torch.cat([named[n].grad.detach().cpu().reshape(-1) for n in names]).double()
    want = torch.cat([grads_o[n].reshape(-1) for n in names]).double()
    cos = float((got @ want) / (got.norm() * want.norm()))
    assert cos > 0.999, cos
    assert abs(float(got.norm()) - float(want.norm())) < 2e-2 * float(want.norm())
    assert abs(float(got.norm()) - float(g["grad_total"])) < 2e-2 * float(g["grad_total"])
    norms = np.array([float(named[n].grad.double().norm()) for n in names])
    big = g["grad_norms"] > 1e-3 * g["grad_norms"].max()
    np.testing.assert_allclose(norms[big], g["grad_norms"][big], rtol=5e-2)
    # BatchNorm running statistics were updated like the reference's (two views -> two updates)
    buf = dict(model.named_buffers())
    for n in ("encoder.stem.1.running_mean", "encoder.backbone.0.0.fc1.1.running_var",
              "encoder.backbone.14.1.fc2.1.running_mean"):
        assert torch.allclose(buf[n].cpu(), stats_o[n], rtol=2e-3, atol=2e-4), n
    assert int(buf["encoder.stem.1.num_batches_tracked"]) == 2


def test_fused_train_step_matches_reference_post_step(golden_dir):
    from neuralsampleid_b200.train import FusedClipAdam, train_step
    g = np.load(os.path.join(golden_dir, "simclr_train_b8.npz"))
    model, sd = _model(5)
    model.train()
    opt = FusedClipAdam(model.parameters(), lr=CFG["lr"], max_norm=1.0)
    s_i, s_j = _inputs(8)
    loss = train_step(model, s_i.to(DEV), s_j.to(DEV), CFG, opt)
    assert abs(loss.item() - float(g["loss"])) < 1e-3 * abs(float(g["loss"]))
    assert abs(opt.grad_norm() - float(g["grad_total"])) < 2e-2 * float(g["grad_total"])
    named = dict(model.named_parameters())
    got = np.concatenate([named[n].detach().cpu().reshape(-1)[:: max(1, named[n].numel() // 16)][:16].numpy()
                          for n in ("encoder.stem.0.weight", "encoder.backbone.0.0.fc1.0.weight",
                                    "encoder.backbone.7.1.fc2.0.weight", "encoder.proj.weight",
                                    "projector.2.weight", "peak_extractor.convs.0.weight")])
    before = np.concatenate([sd[n].reshape(-1)[:: max(1, sd[n].numel() // 16)][:16].numpy()
                             for n in ("encoder.stem.0.weight", "encoder.backbone.0.0.fc1.0.weight",
                                       "encoder.backbone.7.1.fc2.0.weight", "encoder.proj.weight",
                                       "projector.2.weight", "peak_extractor.convs.0.weight")])
    # first Adam step moves every weight by ~lr * sign(grad): compare the applied update
    upd, upd_ref = got - before, g["post_step_sample"] - before
    assert np.abs(upd).max() <= 1.01 * CFG["lr"] and np.abs(upd).max() > 0.5 * CFG["lr"]
    agree = np.mean(np.abs(upd - upd_ref) < 0.1 * CFG["lr"])
    assert agree > 0.9, agree
    # a second step runs and changes the loss
    loss2 = train_step(model, s_i.to(DEV), s_j.to(DEV), CFG, opt)
    assert torch.isfinite(loss2) and opt.step_count == 2
    # eval forward sees the updated weights (prepared-weight caches are invalidated)
    model.eval()
    with torch.no_grad():
        out = model(s_i.to(DEV), s_j.to(DEV))
    assert bool(torch.isfinite(out[2]).all())


def test_clip_adam_kernel_matches_oracle():
    from neuralsampleid_b200 import ops
    n = 10007
    p = synth.synth_normal((n,), 1)
    gr = synth.synth_normal((n,), 2) * 3.0
    p_ref, m_ref, v_ref = p.clone(), torch.zeros(n), torch.zeros(n)
    pd, gd = p.to(DEV), gr.to(DEV)
    md, vd = torch.zeros(n, device=DEV), torch.zeros(n, device=DEV)
    for step in (1, 2, 3):
        g2 = gr.clone()
        O.clip_grad_norm_([g2], 1.0)
        O.adam_step(p_ref, g2, m_ref, v_ref, step, 8e-5)
        sq = torch.zeros(1, device=DEV, dtype=torch.float64)
        ops.sq_norm(gd, sq)
        ops.adam_clip_step(pd, gd, md, vd, 8e-5, 0.9, 0.999, 1e-8, step, 1.0, sq)
        assert abs(float(sq.sqrt()) - float(gr.double().norm())) < 1e-6 * float(gr.norm())
    assert torch.allclose(pd.cpu(), p_ref, rtol=1e-6, atol=1e-7)
    assert torch.allclose(md.cpu(), m_ref, rtol=1e-5, atol=1e-9)
