"""One eager contrastive train step inside a profiler range (for `ncu --profile-from-start off`)."""
import os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neuralsampleid_b200.train import FusedClipAdam, train_step
import bench_extra as BE
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = "cuda:0"
model = BE._model(dev, 5).train()
opt = FusedClipAdam(model.parameters(), lr=BE.CFG["lr"], max_norm=1.0)
g = torch.Generator().manual_seed(2)
x_i = torch.randn((pairs, 64, 128), generator=g).to(dev)
x_j = (x_i.cpu() + 0.1 * torch.randn((pairs, 64, 128), generator=g)).to(dev)
with torch.no_grad():
    for _ in range(2):
        train_step(model, x_i, x_j, BE.CFG, opt)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    train_step(model, x_i, x_j, BE.CFG, opt)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
