"""Parity diagnostic (GPU): how often, and at which reference gap, do the kNN neighbour lists of the sm_100a path
differ from the oracle's -- per block, per GEMM engine, teacher-forced and free-running -- and how large is the error
of the selected distances.  Prints one JSON line per (engine, mode).  Usage:
    python scripts/diag_parity.py [--segments 512] [--engines auto,bf16x3,3xtf32,simt] [--k 3]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import grafp_oracle as O  # noqa: E402
from oracle import synth  # noqa: E402

BINS = [1e-6, 4e-6, 1e-5, 1e-4]
CFG = dict(n_mels=64, n_frames=128, patch_bins=4, patch_frames=8, n_filters=8)


def min_gap(dist, kk):
    kk1 = min(kk + 1, dist.shape[-1])
    vals = -torch.topk(-dist, k=kk1).values
    return (vals[..., 1:] - vals[..., :-1]).min(dim=-1).values          # (B, N)


def bin_counts(gaps):
    """rows per reference-gap bin: <=1e-6, (1e-6,4e-6], (4e-6,1e-5], (1e-5,1e-4], >1e-4"""
    edges = [-1.0] + BINS + [float("inf")]
    return [int(((gaps > lo) & (gaps <= hi)).sum()) for lo, hi in zip(edges[:-1], edges[1:])]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--segments", type=int, default=512)
    ap.add_argument("--engines", default="auto,bf16x3,3xtf32,simt")
    ap.add_argument("--k", type=int, default=3)
    args = ap.parse_args()
    from neuralsampleid_b200 import ops, _lib
    from neuralsampleid_b200.encoder.graph_encoder import GraphEncoder
    dev = "cuda:0"
    k, B = args.k, args.segments
    sd = synth.synth_state(synth.encoder_state_spec("t", 8, 1024, 256), 1234)
    enc = GraphEncoder(cfg=CFG, in_channels=8, k=k)
    enc.load_state_dict(sd)
    enc = enc.to(dev).eval()
    x = synth.synth_uniform((B, 8, 256), 4242)
    taps = []
    with torch.no_grad():
        want = O.encoder_forward(sd, x, k=k, taps=taps)
    blocks = [t for t in taps if t["kind"] == "block"]
    gaps = [min_gap(o["dist"], k) for o in blocks]
    ref_sorted = [(-torch.topk(-o["dist"], k=k).values) for o in blocks]
    forced = [t["idx"].int().to(dev) for t in blocks]
    rel = lambda a, b: ((a.double() - b.double()).norm(dim=1) / b.double().norm(dim=1))

    # (0) the kNN kernels alone on the oracle's own kNN input
    for name, eng in (("tc", _lib.ENGINE_AUTO), ("simt", _lib.ENGINE_SIMT)):
        per = []
        for i, o in enumerate(blocks):
            Bc, C, N = o["knn_in"].shape[:3]
            nodes = o["knn_in"].reshape(Bc, C, N).transpose(1, 2).reshape(Bc * N, C).contiguous().to(dev)
            idx, dist = ops.knn(nodes, Bc, N, k, 1, return_dist=True, engine=eng)
            diff = (idx.cpu().long() != o["idx"]).any(-1)
            same = ~diff
            derr = (dist.cpu() - ref_sorted[i])[same].abs()
            per.append({"block": i, "N": N, "C": C, "rows": diff.numel(), "flips": int(diff.sum()),
                        "flips_by_ref_gap": bin_counts(gaps[i][diff]), "rows_by_ref_gap": bin_counts(gaps[i]),
                        "dist_err_max": float(derr.max()), "dist_err_p999": float(derr.flatten().kthvalue(
                            max(1, int(0.999 * derr.numel()))).values)})
        print(json.dumps({"mode": "knn_kernel_on_oracle_input", "knn_engine": name, "k": k, "segments": B,
                          "gap_bins": BINS, "blocks": per}), flush=True)

    for ename in args.engines.split(","):
        ops._engine_override = None if ename == "auto" else ename
        try:
            with torch.no_grad():
                tf, tr_ = [], []
                emb_f = enc(x.to(dev), forced_idx=forced, taps=tf)
                emb = enc(x.to(dev), taps=tr_)
                per = []
                for i, (t, o) in enumerate(zip(tf, blocks)):
                    N = o["idx"].shape[1]
                    idx, dist = ops.knn(t["fc1"], B, N, k, 1, return_dist=True)
                    diff = (idx.cpu().long() != o["idx"]).any(-1)
                    derr = (dist.cpu() - ref_sorted[i])[~diff].abs()
                    ki = o["knn_in"].reshape(B, -1, N).transpose(1, 2).reshape(B * N, -1)
                    ferr = (t["fc1"].cpu() - ki).abs().max() / ki.abs().max()
                    per.append({"block": i, "flips": int(diff.sum()), "rows": diff.numel(),
                                "flips_by_ref_gap": bin_counts(gaps[i][diff]),
                                "dist_err_max": float(derr.max()), "knn_input_err_rel_max": float(ferr)})
        finally:
            ops._engine_override = None
        out = {"mode": "encoder", "gemm_engine": ename, "k": k, "segments": B, "gap_bins": BINS,
               "teacher_forced_emb_rel_max": float(rel(emb_f.cpu(), want).max()),
               "teacher_forced_blocks": per}
        for tol in (1e-6, 4e-6, 1e-5):
            tr = O.CascadeTracker(B)
            first = []
            for i, (t, o) in enumerate(zip(tr_, blocks)):
                before = int(tr.alive.sum())
                tr.update(i, t["idx"].cpu(), o["idx"], o["dist"], k, tol)
                first.append(before - int(tr.alive.sum()))
            r = rel(emb.cpu(), want)
            out["free_running_tol_%g" % tol] = {
                "off_tie_rows": tr.bad, "tie_flip_rows": tr.tie_flips, "segments_alive": int(tr.alive.sum()),
                "segments_lost_per_block": first,
                "alive_emb_rel_max": float(r[tr.alive].max()) if tr.alive.any() else None}
        out["free_running_emb_rel_median_all"] = float(rel(emb.cpu(), want).median())
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
