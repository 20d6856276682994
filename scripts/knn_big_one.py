"""One knn_big configuration (for ncu): python scripts/knn_big_one.py N k d [log2_nodes] [reps]"""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neuralsampleid_b200 import ops
N, k, d = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
lg = int(sys.argv[4]) if len(sys.argv) > 4 else 20
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
B = (1 << lg) // N
x = torch.randn((B * N, 64), device="cuda:0", generator=torch.Generator(device="cuda:0").manual_seed(4))
ev = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
ev[0].record()
for i in range(reps):
    idx = ops.knn(x, B, N, k, d)
    ev[i + 1].record()
torch.cuda.synchronize()
print(N, k, d, ops.knn_engine(B, N, 64, k, d), [round(ev[i].elapsed_time(ev[i + 1]), 3) for i in range(reps)])
