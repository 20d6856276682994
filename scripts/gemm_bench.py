"""Times individual GEMM shapes of the encoder through ops.linear (CUDA events, L2 flushed by size).
usage: python scripts/gemm_bench.py [tag]   (engine / cluster mode via GRAFP_* env)"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neuralsampleid_b200 import ops, _prep

DEV = "cuda:0"
SHAPES = [  # M, K, N, a_split, out_split, residual
    (262144, 1024, 256, True, False, True),
    (262144, 256, 1024, False, True, False),
    (131072, 2048, 512, True, False, True),
    (131072, 512, 2048, False, True, False),
    (262144, 512, 256, True, False, True),
    (262144, 256, 256, False, False, False),
    (1048576, 256, 64, True, False, True),
    (1048576, 64, 256, False, True, False),
    (524288, 512, 128, True, False, True),
    # narrow fp32-A shapes of stages 1-2 (memory-bound)
    (1048576, 64, 64, False, False, False),
    (524288, 128, 128, False, False, False),
    (524288, 128, 512, False, True, False),
    (1048576, 128, 64, True, False, True),
    # what the fp32-A shapes would cost with a pre-split A (dual-output producers)
    (262144, 256, 1024, True, True, False),
    (131072, 512, 2048, True, True, False),
    (1048576, 64, 256, True, True, False),
    (262144, 256, 256, True, False, False),
]
torch.manual_seed(0)
for M, K, N, a_split, out_split, res in SHAPES:
    w = (torch.randn(N, K, device=DEV) / K ** 0.5)
    lin = _prep.make_linear(w, torch.ones(N, device=DEV), torch.zeros(N, device=DEV), 1)
    a = torch.randn(M, K, device=DEV)
    if a_split:
        hi = a.bfloat16(); lo = (a - hi.float()).bfloat16()
        a_in = ops.SplitAct(torch.stack([hi, lo]).contiguous())
        del hi, lo, a
    else:
        a_in = a
    r = torch.randn(M, N, device=DEV) if res else None
    for _ in range(3):
        y = ops.linear(a_in, lin, "relu" if out_split else None, 0.0, r, out_split=out_split)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(11)]
    ev[0].record()
    for i in range(10):
        y = ops.linear(a_in, lin, "relu" if out_split else None, 0.0, r, out_split=out_split)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(10))
    fl = 2.0 * M * K * N
    print("m=%8d k=%5d n=%5d %s%s%s  med %.1f us  min %.1f us  %.0f TF/s useful" % (
        M, K, N, "As " if a_split else "A32", " Ys" if out_split else " Y32", " res" if res else "    ",
        1e3 * ts[5], 1e3 * ts[0], fl / (ts[5] * 1e-3) / 1e12), flush=True)
    del a_in, y, r, lin, w
