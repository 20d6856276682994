"""Diagnostic: teacher-forced encoder train fwd/bwd vs the oracle for each GEMM engine."""
import sys, os, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import grafp_oracle as O, synth
from neuralsampleid_b200 import autograd as A, ops
from neuralsampleid_b200.encoder.graph_encoder import GraphEncoder
CFG = dict(n_mels=64, n_frames=128, patch_bins=4, patch_frames=8, n_filters=8)
DEV = "cuda:0"
k, B = 5, 6
sd = synth.synth_state(synth.encoder_state_spec("t", 8, 1024, 256), 1234)
x = synth.synth_uniform((B, 8, 256), 77)
names = [n for n, t in sd.items() if t.dtype == torch.float32 and "running" not in n and "relative_pos" not in n]
params = {n: (t.clone().requires_grad_(True) if n in names else t.clone()) for n, t in sd.items()}
taps = []
emb_o = O.encoder_forward(params, x, k=k, training=True, stats={}, taps=taps)
G = synth.synth_normal(tuple(emb_o.shape), 78)
(emb_o * G).sum().backward()
forced = [t["idx"].int().to(DEV) for t in taps if t["kind"] == "block"]
for eng in ("simt", "3xtf32", "bf16x3"):
    ops._engine_override = eng
    enc = GraphEncoder(cfg=CFG, in_channels=8, k=k)
    enc.load_state_dict(sd)
    enc = enc.to(DEV).train()
    emb, _, tape = A.encoder_train_fwd(enc, ops.nchw_to_nodes(x.to(DEV)), B, 256, forced)
    rel = float(((emb.cpu() - emb_o.detach()).norm(dim=1) / emb_o.detach().norm(dim=1)).max())
    grads = {}
    A.encoder_train_bwd(tape, G.to(DEV), grads)
    named = dict(enc.named_parameters())
    ga, wa, worst = [], [], (1.0, "", 1.0)
    for n in names:
        w = params[n].grad
        if w is None or float(w.norm()) / w.numel() ** 0.5 < 1e-6:
            continue
        g = grads[named[n]].cpu().double().reshape(-1); w = w.double().reshape(-1)
        if float(g.norm()) == 0: continue
        cos = float(g @ w / (g.norm() * w.norm()))
        if cos < worst[0]: worst = (cos, n, float(g.norm() / w.norm()))
        ga.append(g); wa.append(w)
    ga, wa = torch.cat(ga), torch.cat(wa)
    print("%-7s emb rel %.2e  global cos %.6f  norm ratio %.5f  worst param %s" %
          (eng, rel, float(ga @ wa / (ga.norm() * wa.norm())), float(ga.norm() / wa.norm()), worst))
ops._engine_override = None
