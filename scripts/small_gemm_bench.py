"""Back-to-back launch time of the train step's small GEMM / wgrad shapes (warm caches, CUDA events)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neuralsampleid_b200 import ops, _prep
dev = "cuda:0"
torch.manual_seed(0)
def t(fn, n=50):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return 1e3 * e0.elapsed_time(e1) / n
for (M, K, N) in ((8192, 64, 64), (8192, 64, 256), (8192, 256, 64), (4096, 128, 512), (2048, 256, 1024), (2048, 1024, 256), (1024, 512, 2048), (1024, 2048, 512), (32, 1024, 4096)):
    a = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev) / K ** 0.5
    lin = _prep.make_linear(w, None, None)
    dy = torch.randn(M, N, device=dev)
    us_f = t(lambda: ops.linear(a, lin))
    us_s = t(lambda: ops.linear(a, lin, engine=1))
    us_w = t(lambda: ops.gemm_wgrad(dy, a, None, N))
    us_ws = t(lambda: ops.gemm_wgrad(dy, a, None, N, engine=1))
    us_cs = t(lambda: ops.col_stats(dy))
    print("M=%5d K=%4d N=%4d  fwd tc %6.1f us  simt %6.1f us | wgrad tc %6.1f us  simt %6.1f us | col_stats %5.1f us" % (M, K, N, us_f, us_s, us_w, us_ws, us_cs), flush=True)
