"""Times the kNN + aggregate kernels at the four encoder stages (B = 4096), CUDA events.
usage: python scripts/knn_bench.py"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neuralsampleid_b200 import ops

DEV = "cuda:0"
B = int(os.environ.get("KB", "4096"))
torch.manual_seed(0)
for N, C in ((256, 64), (128, 128), (64, 256), (32, 512)):
    x = torch.randn(B * N, C, device=DEV)
    rs = (x * x).sum(1).contiguous()
    for k in (3,):
        for _ in range(3):
            idx = ops.knn(x, B, N, k, 1, row_sumsq=rs)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
        ev[0].record()
        for i in range(10):
            idx = ops.knn(x, B, N, k, 1, row_sumsq=rs)
            ev[i + 1].record()
        torch.cuda.synchronize()
        ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(10))
        by = B * N * C * 4 + B * N * k * 4
        for _ in range(3):
            m = ops.mr_aggregate(x, idx, B, N)
        ev2 = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
        ev2[0].record()
        for i in range(10):
            m = ops.mr_aggregate(x, idx, B, N)
            ev2[i + 1].record()
        torch.cuda.synchronize()
        ta = sorted(ev2[i].elapsed_time(ev2[i + 1]) for i in range(10))
        bya = 2 * B * N * C * 4 + B * N * k * 4
        print("N=%4d C=%4d k=%d  knn med %.1f us (%.2f TB/s)   aggregate med %.1f us (%.2f TB/s)" % (
            N, C, k, 1e3 * ts[5], by / (ts[5] * 1e-3) / 1e12, 1e3 * ta[5], bya / (ta[5] * 1e-3) / 1e12), flush=True)
    del x, rs, idx, m
