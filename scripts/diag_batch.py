"""Diagnostic: where does enc(x)[sub] first differ from enc(x[sub]) (and from a second enc(x))?"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import synth  # noqa: E402
from neuralsampleid_b200.encoder.graph_encoder import GraphEncoder  # noqa: E402

DEV = "cuda:0"
cfg = dict(n_mels=64, n_frames=128, patch_bins=4, patch_frames=8, n_filters=8)
sd = synth.synth_state(synth.encoder_state_spec("t", 8, 1024, 256), 1234)
enc = GraphEncoder(cfg=cfg, in_channels=8, k=3)
enc.load_state_dict(sd)
enc = enc.to(DEV).eval()
B = int(os.environ.get("DIAG_B", "4096"))
x = synth.synth_uniform((B, 8, 256), 90).to(DEV)
sub = torch.arange(0, B, 512, device=DEV)


def rows_of(t, nb, sel):
    n = t.shape[0] // nb
    return t.reshape(nb, n, *t.shape[1:])[sel]


def compare(tag, ta, tb, nb_a, sel_a):
    for i, (a, b) in enumerate(zip(ta, tb)):
        for key in ("in", "fc1", "idx", "grapher", "out"):
            va = a[key]
            vb = b[key]
            if key == "idx":
                va = va.reshape(nb_a, -1, va.shape[-1])[sel_a] if sel_a is not None else va
                vb = vb.reshape(va.shape)
            else:
                va = rows_of(va, nb_a, sel_a) if sel_a is not None else va
                vb = vb.reshape(va.shape)
            if not torch.equal(va, vb):
                diff = (va != vb)
                seg = diff.reshape(diff.shape[0], -1).any(1).nonzero().flatten().tolist() if sel_a is not None else None
                if key == "idx":
                    nbad = int(diff.any(-1).sum())
                    print("%s: block %d %s differs: %d rows, segments %s" % (tag, i, key, nbad, seg))
                else:
                    d = (va - vb).abs()
                    print("%s: block %d %s differs: %d elems, max abs %.3e (scale %.3e), segments %s"
                          % (tag, i, key, int(diff.sum()), float(d.max()), float(va.abs().max()), seg))
                    rows = diff.reshape(-1, diff.shape[-1]).any(1).nonzero().flatten()
                    print("   first differing flat rows:", rows[:16].tolist(), "of", diff.reshape(-1, diff.shape[-1]).shape[0])
                return False
    print("%s: all taps equal" % tag)
    return True


with torch.no_grad():
    t1, t2, t3 = [], [], []
    e1 = enc(x, taps=t1)
    e2 = enc(x, taps=t2)
    e3 = enc(x[sub], taps=t3)
print("determinism:", torch.equal(e1, e2))
compare("run1 vs run2", t1, t2, B, None)
print("batch independence:", torch.equal(e1[sub], e3))
compare("big[sub] vs sub", t1, t3, B, sub)
