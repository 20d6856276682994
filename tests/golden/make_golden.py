#!/usr/bin/env python
"""Golden-vector generator: runs the REFERENCE itself (imported read-only from
/root/reference with the stub recipe in refimport.py) on portable seeded weights / inputs
(oracle/synth.py) and writes small .npz fixtures next to this script.

Run here (the container that has /root/reference):   python tests/golden/make_golden.py
The fixtures are committed; the GPU box never needs the reference tree.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

from oracle import synth  # noqa: E402
from refimport import import_reference  # noqa: E402

WEIGHT_SEED = 1234


def capture_encoder(ref, k, B, x_seed, keep_graphs=1):
    enc = ref.GraphEncoder(cfg=ref.cfg, in_channels=ref.cfg["n_filters"], k=k).eval()
    sd = synth.synth_state(synth.encoder_state_spec("t", ref.cfg["n_filters"], 1024, 256), WEIGHT_SEED)
    assert list(sd.keys()) == list(enc.state_dict().keys()), "state_spec order drifted from the reference"
    for n, t in enc.state_dict().items():
        assert tuple(t.shape) == tuple(sd[n].shape), (n, t.shape, sd[n].shape)
    relpos = {n: t.clone() for n, t in enc.state_dict().items() if n.endswith("relative_pos")}
    enc.load_state_dict(sd)
    x = synth.synth_uniform((B, ref.cfg["n_filters"], 256), x_seed)
    out = {"x": x.numpy(), "weights_sha256": np.array(synth.state_sha256(sd))}

    taps = {}

    def hook_knn(i):
        def f(mod, inp, res):
            taps["knn_in_%d" % i] = inp[0].detach().clone()          # (B,C,N,1) un-normalised
            taps["idx_%d" % i] = res[0].detach().clone()             # (B,N,k)
        return f

    def hook_out(i):
        def f(mod, inp, res):
            taps["out_%d" % i] = res.detach().clone()
        return f

    bi = 0
    for i, m in enumerate(enc.backbone):
        m.register_forward_hook(hook_out(i))
        if hasattr(m, "conv"):
            continue
        m[0].graph_conv.dilated_knn_graph.register_forward_hook(hook_knn(i))
        bi += 1
    with torch.no_grad():
        emb = enc(x)
    out["emb"] = emb.numpy()
    for name, t in taps.items():
        if name.startswith("idx_"):
            out[name] = t.numpy().astype(np.int16)
        elif name.startswith("knn_in_"):
            out[name] = t[:keep_graphs, :, :, 0].numpy()             # (g,C,N)
        else:
            flat = t.reshape(t.shape[0], -1)
            out[name + "_sample"] = flat[:, ::61].numpy()
            out[name + "_absmean"] = flat.abs().mean(dim=1).numpy()
    return out, relpos


def capture_dygraph(ref, k, d, N, B, seed):
    m = ref.DyGraphConv2d(64, 128, k, d, "mr", "relu", "batch", True).eval()
    spec = [("gconv.nn.0.weight", (128, 32, 1, 1), "w"), ("gconv.nn.0.bias", (128,), "b")] + \
        synth._bn("gconv.nn.1", 128)
    sd = synth.synth_state(spec, WEIGHT_SEED + 1)
    m.load_state_dict(sd)
    x = synth.synth_normal((B, 64, N, 1), seed)
    with torch.no_grad():
        edge = m.dilated_knn_graph(x)
        y = m(x)
    return {"x": x.numpy(), "idx": edge[0].numpy().astype(np.int16), "y": y.numpy()}


def capture_graphconv(ref, conv, act, k, d, cin, cout, N, B, seed):
    """The reference's other GraphConv2d variants (edge / sage / gin) through DyGraphConv2d."""
    m = ref.DyGraphConv2d(cin, cout, k, d, conv, act, "batch", True).eval()
    sd = synth.synth_state(synth.graphconv_state_spec(conv, cin, cout), WEIGHT_SEED + 3)
    assert sorted(sd.keys()) == sorted(m.state_dict().keys()), (conv, sorted(m.state_dict().keys()))
    m.load_state_dict(sd)
    x = synth.synth_normal((B, cin, N, 1), seed)
    with torch.no_grad():
        edge = m.dilated_knn_graph(x)
        y = m(x)
    return {"x": x.numpy(), "idx": edge[0].numpy().astype(np.int16), "y": y.numpy()}


def capture_reranker(B, N, seed):
    """CrossAttentionClassifier (downstream.py:30-79) in eval mode on (B, 512, N) node matrices."""
    from refimport import import_reference_reranker
    m = import_reference_reranker()(512, 4, 128, 100, True).eval()
    sd = synth.synth_state(synth.reranker_state_spec(512, 128, 100), WEIGHT_SEED + 4)
    assert sorted(sd.keys()) == sorted(m.state_dict().keys()), sorted(m.state_dict().keys())
    m.load_state_dict(sd)
    x_i, x_j = synth.reranker_inputs(B, N, seed)                           # half matching, half unrelated pairs
    with torch.no_grad():
        y = m(x_i, x_j)
    return {"y": y.numpy(), "B": np.int64(B), "N": np.int64(N), "seed": np.int64(seed)}


def capture_logmel(n_samples, seed):
    """The reference's front end exactly as modules/transformations.py:27-34 builds it (torchaudio transforms with the
    grafp.yaml parameters) and its eval-branch segmentation (:96-104)."""
    from torchaudio.transforms import AmplitudeToDB, MelSpectrogram
    logmelspec = torch.nn.Sequential(MelSpectrogram(sample_rate=16000, win_length=1024, hop_length=512, n_fft=1024,
                                                    n_mels=64), AmplitudeToDB())
    wave = synth.synth_wave(n_samples, seed)
    with torch.no_grad():
        X = logmelspec(wave)                                            # (n_mels, T)
        seg = X.transpose(1, 0).unfold(0, size=128, step=int(128 * (1 - 0.875)))
    return {"db": X.numpy(), "n_segments": np.int64(seg.shape[0]), "seg_last": seg[-1].numpy(),
            "n_samples": np.int64(n_samples), "seed": np.int64(seed)}


def capture_ntxent(ref, B, seed):
    z_i = torch.nn.functional.normalize(synth.synth_normal((B, 128), seed), dim=1)
    z_j = torch.nn.functional.normalize(z_i + 0.3 * synth.synth_normal((B, 128), seed + 1), dim=1)
    zi = z_i.clone().requires_grad_(True)
    zj = z_j.clone().requires_grad_(True)
    loss = ref.ntxent_loss(zi, zj, {"tau": 0.05})
    loss.backward()
    return {"z_i": z_i.numpy(), "z_j": z_j.numpy(), "loss": loss.detach().numpy(),
            "g_i": zi.grad.numpy(), "g_j": zj.grad.numpy(), "tau": np.float32(0.05)}


def capture_simclr(ref, k, B, training):
    cfg = dict(ref.cfg)
    enc = ref.GraphEncoder(cfg=cfg, in_channels=cfg["n_filters"], k=k)
    model = ref.SimCLR(cfg, encoder=enc)
    sd = synth.synth_state(synth.simclr_state_spec(cfg, "t"), WEIGHT_SEED + 2)
    assert list(sd.keys()) == list(model.state_dict().keys()), "simclr spec order drifted"
    model.load_state_dict(sd)
    model.train(training)
    s_i = synth.synth_normal((B, cfg["n_mels"], cfg["n_frames"]), 21)
    s_j = s_i + 0.1 * synth.synth_normal((B, cfg["n_mels"], cfg["n_frames"]), 22)
    out = {"weights_sha256": np.array(synth.state_sha256(sd))}
    if not training:
        # per-block neighbour lists of both views (the encoder runs twice: view i, then view j)
        calls = []
        for m in enc.backbone:
            if not hasattr(m, "conv"):
                m[0].graph_conv.dilated_knn_graph.register_forward_hook(
                    lambda mod, inp, res: calls.append(res[0].detach().clone()))
        with torch.no_grad():
            h_i, h_j, z_i, z_j = model(s_i, s_j)
        nb = len(calls) // 2
        for b in range(nb):
            out["idx_i_%d" % b] = calls[b].numpy().astype(np.int16)
            out["idx_j_%d" % b] = calls[nb + b].numpy().astype(np.int16)
        out.update(h_i=h_i.numpy(), h_j=h_j.numpy(), z_i=z_i.numpy(), z_j=z_j.numpy())
        return out
    h_i, h_j, z_i, z_j = model(s_i, s_j)
    loss = ref.ntxent_loss(z_i, z_j, cfg)
    loss.backward()
    names, norms = [], []
    for n, p in model.named_parameters():
        if p.grad is not None:
            names.append(n)
            norms.append(float(p.grad.double().norm()))
    total = float(np.sqrt(sum(v * v for v in norms)))
    out.update(z_i=z_i.detach().numpy(), z_j=z_j.detach().numpy(), h_i=h_i.detach().numpy(),
               loss=loss.detach().numpy(), grad_names=np.array(names), grad_norms=np.array(norms),
               grad_total=np.float64(total))
    new = model.state_dict()
    rm = [float(new[n].double().sum()) for n in new if n.endswith("running_mean")]
    rv = [float(new[n].double().sum()) for n in new if n.endswith("running_var")]
    out.update(running_mean_sums=np.array(rm), running_var_sums=np.array(rv))
    # one clip + Adam step as train.py:70-75 does it
    opt = torch.optim.Adam(model.parameters(), lr=cfg["lr"])
    torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm=1.0)
    opt.step()
    new = model.state_dict()
    out["post_step_sample"] = np.concatenate(
        [new[n].reshape(-1)[:: max(1, new[n].numel() // 16)][:16].numpy()
         for n in ("encoder.stem.0.weight", "encoder.backbone.0.0.fc1.0.weight",
                   "encoder.backbone.7.1.fc2.0.weight", "encoder.proj.weight",
                   "projector.2.weight", "peak_extractor.convs.0.weight")])
    return out


def main():
    torch.set_num_threads(1)          # single-thread reductions: reproducible fixtures
    ref = import_reference()
    if "--only-simclr-eval" in sys.argv:      # re-mint one fixture without touching the others
        np.savez_compressed(os.path.join(HERE, "simclr_eval_b4.npz"), **capture_simclr(ref, 3, 4, False))
        return
    g, relpos = capture_encoder(ref, k=3, B=4, x_seed=11)
    np.savez_compressed(os.path.join(HERE, "encoder_t_k3.npz"), **g)
    np.savez_compressed(os.path.join(HERE, "relative_pos_checksums.npz"),
                        **{n: np.array([float(t.double().sum()), float(t.double().abs().sum()),
                                        float(t[0, 0, -1]), float(t[0, -1, 0])]) for n, t in relpos.items()})
    g5, _ = capture_encoder(ref, k=5, B=2, x_seed=12, keep_graphs=0)
    g5 = {n: v for n, v in g5.items() if n in ("x", "emb", "weights_sha256") or n.startswith("idx_")}
    np.savez_compressed(os.path.join(HERE, "encoder_t_k5.npz"), **g5)
    np.savez_compressed(os.path.join(HERE, "dygraph_k9_d2.npz"), **capture_dygraph(ref, 9, 2, 256, 2, 31))
    np.savez_compressed(os.path.join(HERE, "dygraph_k4_d3_n96.npz"), **capture_dygraph(ref, 4, 3, 96, 3, 32))
    for conv, act in (("edge", "gelu"), ("edge", "relu"), ("sage", "relu"), ("gin", "leakyrelu")):
        np.savez_compressed(os.path.join(HERE, "graphconv_%s_%s.npz" % (conv, act)),
                            **capture_graphconv(ref, conv, act, 4, 2, 64, 128, 64, 3, 51))
    np.savez_compressed(os.path.join(HERE, "reranker_b16_n32.npz"), **capture_reranker(16, 32, 61))
    np.savez_compressed(os.path.join(HERE, "logmel_5s.npz"), **capture_logmel(81920, 71))
    np.savez_compressed(os.path.join(HERE, "ntxent_b16.npz"), **capture_ntxent(ref, 16, 41))
    np.savez_compressed(os.path.join(HERE, "simclr_eval_b4.npz"), **capture_simclr(ref, 3, 4, False))
    np.savez_compressed(os.path.join(HERE, "simclr_train_b8.npz"), **capture_simclr(ref, 5, 8, True))
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".npz"):
            print("%-32s %8.1f KB" % (f, os.path.getsize(os.path.join(HERE, f)) / 1024))


if __name__ == "__main__":
    main()
