#!/usr/bin/env python
"""Secondary measurements for BASELINE.json configs 3-5 (bench.py is the judged contract, configs[1]).

  python bench_extra.py train  [--pairs 32] [--steps 10]     config 3: contrastive train step (N ranks via torchrun)
  python bench_extra.py db     [--segments 1000000]          config 4: spectrogram segments -> 128-d fingerprints
  python bench_extra.py sweep                                 config 5: dynamic-graph stress sweep (kNN + aggregate)
  python bench_extra.py chunks                                generate.py call shape: chunks of 128 segments
  python bench_extra.py search [--segments 1000000]           SURVEY 8f rank 4: exact L2 search of the fingerprint DB

Each mode prints one JSON line (rank 0).  Synthetic data, random-init weights.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

CFG = dict(n_mels=64, n_frames=128, patch_bins=4, patch_frames=8, n_filters=8, tau=0.05,
           d=128, h=1024, u=32, dim=2048, arch="grafp", bsz_train=256, lr=8.0e-5)


def _setup():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    return world, rank, dev


def _model(dev, k):
    from neuralsampleid_b200.encoder.graph_encoder import GraphEncoder
    from neuralsampleid_b200.simclr.simclr import SimCLR
    torch.manual_seed(0)
    model = SimCLR(CFG, encoder=GraphEncoder(cfg=CFG, in_channels=CFG["n_filters"], k=k))
    return model.to(dev)


def _finish(world):
    """End of a mode: flush and leave WITHOUT tearing NCCL down.  destroy_process_group() after a CUDA graph that
    captured NCCL collectives (the graphed train step) hung all 8 ranks until the job's timeout."""
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
    sys.stdout.flush()
    sys.stderr.flush()
    if world > 1:
        os._exit(0)


def _timed(fn, steps, world, dev):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    return ms


def mode_train(args):
    from neuralsampleid_b200.train import FusedClipAdam, GraphedTrainStep, train_step
    world, rank, dev = _setup()
    model = _model(dev, 5).train()                     # train.py default --k 5
    opt = FusedClipAdam(model.parameters(), lr=CFG["lr"], max_norm=1.0)
    g = torch.Generator().manual_seed(2 + rank)
    x_i = torch.randn((args.pairs, 64, 128), generator=g).to(dev)
    x_j = (x_i.cpu() + 0.1 * torch.randn((args.pairs, 64, 128), generator=g)).to(dev)
    losses = []

    def step():
        losses.append(train_step(model, x_i, x_j, CFG, opt).clone())
    for _ in range(3):
        step()
    ms_eager = _timed(step, args.steps, world, dev)
    g = GraphedTrainStep(model, CFG, opt, args.pairs)

    def gstep():
        losses.append(g(x_i, x_j).clone())
    for _ in range(2):
        gstep()
    ms = _timed(gstep, args.steps, world, dev)
    if rank == 0:
        print(json.dumps({"mode": "train", "metric": "contrastive train step encoder-segments/s",
                          "value": 2 * args.pairs * world * args.steps / (ms * 1e-3), "unit": "segments/s",
                          "n_gpus": world, "ms_per_step": ms / args.steps, "ms_per_step_eager": ms_eager / args.steps,
                          "launch": "cuda graph replay of the whole step", "pairs_per_gpu": args.pairs,
                          "global_pairs": args.pairs * world, "loss_first": float(losses[0]),
                          "loss_last": float(losses[-1]),
                          "config": "SimCLR(GraphEncoder t, k=5) fwd+bwd, NT-Xent (global negatives via NCCL "
                                    "all-gather), summed grad all-reduce, clip 1.0, Adam 8e-5"}))
    _finish(world)


def mode_db(args):
    from neuralsampleid_b200.graphed import GraphedSimCLR
    from neuralsampleid_b200.parallel import shard_range
    world, rank, dev = _setup()
    model = _model(dev, 3).eval()
    B = args.batch
    lo, hi = shard_range(args.segments, rank, world)
    n_batches = (hi - lo + B - 1) // B
    g = GraphedSimCLR(model, B)
    spec = torch.randn((B, 64, 128), generator=torch.Generator().manual_seed(rank)).pin_memory()
    out = torch.empty((n_batches * B, 128), dtype=torch.float32).pin_memory()     # this rank's slice of the DB
    for _ in range(2):
        g(spec.to(dev))
    it = [0]

    def step():
        i = it[0]
        g.input.copy_(spec, non_blocking=True)
        g.replay()
        out[i * B:(i + 1) * B].copy_(g.z, non_blocking=True)
        it[0] = i + 1
    t0 = time.perf_counter()
    ms = _timed(step, n_batches, world, dev)
    wall = time.perf_counter() - t0
    if rank == 0:
        print(json.dumps({"mode": "db", "metric": "reference-database fingerprinting segments/s (spectrogram -> 128-d)",
                          "value": args.segments / (ms * 1e-3), "unit": "segments/s", "n_gpus": world,
                          "segments": args.segments, "batch": B, "seconds": ms * 1e-3, "wall_seconds": wall,
                          "h2d_bytes_per_segment": 64 * 128 * 4, "d2h_bytes_per_segment": 128 * 4,
                          "config": "SimCLR eval (peak extractor + GraphEncoder t k=3 + projector + L2 norm), CUDA graph "
                                    "replay per batch, pinned host in/out, contiguous segment ranges per rank"}))
    _finish(world)


def mode_sweep(args):
    """BASELINE configs[4] / SURVEY 8(d) config 5: k in {9,16,32} x N in {256,512,1024,2048}, C = 64, dilation 2,
    B * N = 2^20 nodes.  Per point: kNN ms (+ its ratio to the 3-pass TF32 tensor ceiling of 2 N^2 C flops per graph),
    aggregate ms + GB/s on algorithmic bytes, which kNN kernel family ran, and index parity of the first graphs against
    the CPU oracle (rows that differ off documented ties: must be 0)."""
    from neuralsampleid_b200 import ops
    from oracle import grafp_oracle as O
    world, rank, dev = _setup()
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tf32_peak = peaks.get("bf16_tflops_sustained", 1400.0) / 2.0 * 1e12          # dense tf32 = half the bf16 rate
    hbm = peaks.get("hbm_gbs", 6650.0)
    rows = []
    nodes = 1 << args.log2_nodes
    for N in (256, 512, 1024, 2048):
        for k in (9, 16, 32):
            for d in (2,):
                B = nodes // N
                x = torch.randn((B * N, 64), device=dev, generator=torch.Generator(device=dev).manual_seed(4))
                idx = ops.knn(x, B, N, k, d)
                ops.mr_aggregate(x, idx, B, N)
                torch.cuda.synchronize()
                reps = 3
                ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                ev[0].record()
                for _ in range(reps):
                    idx = ops.knn(x, B, N, k, d)
                ev[1].record()
                for _ in range(reps):
                    m = ops.mr_aggregate(x, idx, B, N)
                ev[2].record()
                torch.cuda.synchronize()
                t_knn, t_agg = ev[0].elapsed_time(ev[1]) / reps, ev[1].elapsed_time(ev[2]) / reps
                agg_bytes = B * (2 * N * 64 * 4 + 4 * N * k)
                ceil_ms = 3.0 * 2.0 * N * N * 64 * B / tf32_peak * 1e3
                # index parity of the first graphs against the oracle
                gchk = 2
                xs = x[:gchk * N].cpu().view(gchk, N, 64).transpose(1, 2).unsqueeze(-1).contiguous()
                edge, dist = O.dilated_knn_graph(xs, k, d)
                tie = O.knn_tie_rows(dist, k * d, 4e-6)
                diff = (idx[:gchk].cpu().long() != edge[0]).any(-1)
                rows.append({"N": N, "k": k, "d": d, "graphs": B, "knn_ms": round(t_knn, 3), "agg_ms": round(t_agg, 3),
                             "knn_x_of_3xtf32_ceiling": round(t_knn / ceil_ms, 2), "knn_ceiling_ms": round(ceil_ms, 3),
                             "agg_GBps": round(agg_bytes / (t_agg * 1e-3) / 1e9, 1),
                             "agg_frac_of_hbm": round(agg_bytes / (t_agg * 1e-3) / 1e9 / hbm, 3),
                             "knn_engine": ops.knn_engine(B, N, 64, k, d),
                             "idx_parity": {"rows_checked": int(diff.numel()), "rows_differing": int(diff.sum()),
                                            "off_tie_rows": int((diff & ~tie).sum())}})
                del x, idx, m
    if rank == 0:
        print(json.dumps({"mode": "sweep", "metric": "dynamic-graph stress sweep (C=64, dilation 2)", "nodes": nodes,
                          "rows": rows}))


def mode_chunks(args):
    from neuralsampleid_b200.graphed import GraphedSimCLR
    world, rank, dev = _setup()
    model = _model(dev, 3).eval()
    x = torch.randn((128, 64, 128), device=dev)
    with torch.no_grad():
        for _ in range(3):
            model(x, x)
    ms_eager = _timed(lambda: model._one_view(x), 20, 1, dev) / 20
    g = GraphedSimCLR(model, 128)
    g(x)
    ms_graph = _timed(g.replay, 50, 1, dev) / 50
    if rank == 0:
        print(json.dumps({"mode": "chunks", "metric": "generate.py call shape: 128-segment chunk, one view",
                          "eager_ms": ms_eager, "graph_ms": ms_graph, "eager_seg_s": 128 / (ms_eager * 1e-3),
                          "graph_seg_s": 128 / (ms_graph * 1e-3)}))


def mode_search(args):
    """exact squared-L2 search of the fingerprint database (eval.py's index type 'l2')."""
    from neuralsampleid_b200.db import FlatL2Index
    world, rank, dev = _setup()
    n, nq, k, d = args.segments, args.queries, 20, 128
    g = torch.Generator(device=dev).manual_seed(5)
    db = torch.nn.functional.normalize(torch.randn((n, d), device=dev, generator=g), dim=1)
    q = torch.nn.functional.normalize(db[torch.randint(0, n, (nq,), device=dev, generator=g)] +
                                      0.05 * torch.randn((nq, d), device=dev, generator=g), dim=1)
    index = FlatL2Index(d, dev)
    index.add(db)
    index.search(q, k)                                     # warm-up
    steps = 5
    ms = _timed(lambda: index.search(q, k), steps, world, dev) / steps
    D, I = index.search(q, k)
    t0 = time.perf_counter()
    sample = 64
    import numpy as np                                     # CPU baseline: plain numpy exact search (float64)
    dbn, qn = db.cpu().numpy().astype(np.float64), q[:sample].cpu().numpy().astype(np.float64)
    dist = (qn * qn).sum(1)[:, None] - 2.0 * (qn @ dbn.T) + (dbn * dbn).sum(1)[None, :]
    Iw = np.argsort(dist, axis=1, kind="stable")[:, :k]
    cpu_s = time.perf_counter() - t0
    agree = float((I[:sample].cpu().numpy() == Iw).mean())
    if rank == 0:
        print(json.dumps({"mode": "search", "metric": "exact L2 fingerprint search queries/s", "value": nq / (ms * 1e-3),
                          "unit": "queries/s", "n_gpus": world, "database": n, "queries": nq, "k": k, "ms_per_search": ms,
                          "pair_distances_per_s": nq * n / (ms * 1e-3),
                          "cpu_baseline": {"value": sample / cpu_s, "unit": "queries/s", "kind": "port",
                                           "sample": "%d queries, numpy float64 exact search" % sample},
                          "id_agreement_with_exact": agree,
                          "config": "FlatL2Index: bf16x3 tcgen05 GEMM per 65536-row chunk + threshold top-k scan + merge"}))
    _finish(world)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["train", "db", "sweep", "chunks", "search"])
    ap.add_argument("--pairs", type=int, default=32)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--segments", type=int, default=1000000)
    ap.add_argument("--batch", type=int, default=4096)
    ap.add_argument("--queries", type=int, default=2048)
    ap.add_argument("--log2-nodes", type=int, default=20, help="sweep: B * N = 2^this nodes per point")
    args = ap.parse_args()
    with torch.no_grad():
        {"train": mode_train, "db": mode_db, "sweep": mode_sweep, "chunks": mode_chunks,
         "search": mode_search}[args.mode](args)


if __name__ == "__main__":
    main()
