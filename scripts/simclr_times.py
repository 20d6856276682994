"""Per-entry-point CUDA-event times of one eager SimCLR eval forward (spectrogram segments -> fingerprints) at B segments."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from neuralsampleid_b200 import _lib
from neuralsampleid_b200.encoder.graph_encoder import GraphEncoder
from neuralsampleid_b200.simclr.simclr import SimCLR
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
cfg = dict(bench.CFG, tau=0.05, d=128, h=1024, u=32, dim=2048, arch="grafp", bsz_train=256, lr=8.0e-5)
torch.manual_seed(0)
model = SimCLR(cfg, encoder=GraphEncoder(cfg=cfg, in_channels=cfg["n_filters"], k=3)).to("cuda:0").eval()
x = torch.randn((B, 64, 128), device="cuda:0")
with torch.no_grad():
    for _ in range(2):
        model._one_view(x)
    timer = bench.KernelTimer()
    _lib.set_profiler(timer)
    model._one_view(x)
    _lib.set_profiler(None)
rows = sorted(timer.summary().items(), key=lambda kv: -kv[1]["ms"])
tot = sum(v["ms"] for _, v in rows)
print("total %.3f ms over %d entries" % (tot, len(rows)))
ALL = len(sys.argv) > 2
for k, v in rows:
    if ALL or not k.startswith(("gemm m=26", "gemm m=13", "knn", "aggregate", "ffn_fused", "mrconv", "gemm m=52", "gemm m=10")):
        print("%-50s %.3f ms (%d calls)" % (k, v["ms"], v["calls"]))
