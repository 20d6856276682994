"""Debug: per-role cycle accounting of the fused MRConv -> fc2 kernel (build with GRAFP_NVCC_EXTRA=-DFF_TRACE):
python scripts/mr_trace.py C M"""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neuralsampleid_b200 import ops, _prep
C = int(sys.argv[1]); M = int(sys.argv[2]); FFN = len(sys.argv) > 3 and sys.argv[3] == "ffn"
dev = "cuda:0"
torch.manual_seed(0)
x, m, res = torch.randn(M, C, device=dev), torch.randn(M, C, device=dev).abs(), torch.randn(M, C, device=dev)
l1 = _prep.make_linear(torch.randn(2 * C, C // 2, device=dev) / (C // 2) ** 0.5, torch.ones(2 * C, device=dev),
                       torch.zeros(2 * C, device=dev), 4, dual=True)
l2 = _prep.make_linear(torch.randn(C, 2 * C, device=dev) / (2 * C) ** 0.5, torch.ones(C, device=dev), torch.zeros(C, device=dev))
if FFN:
    l1 = _prep.make_linear(torch.randn(4 * C, C, device=dev) / C ** 0.5, torch.ones(4 * C, device=dev), torch.zeros(4 * C, device=dev))
    l2 = _prep.make_linear(torch.randn(C, 4 * C, device=dev) / (4 * C) ** 0.5, torch.ones(C, device=dev), torch.zeros(C, device=dev))
    run = lambda: ops.ffn_fused(x, l1, l2, "relu")
else:
    run = lambda: ops.mrconv_fc2_fused(x, m, l1, "relu", 0.0, l2, res)
for _ in range(20): y = run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): y = run()
e1.record(); torch.cuda.synchronize()
us = 20 * e0.elapsed_time(e1)
print("%s C %d M %d: %.1f us, %.2f TB/s on its algorithmic bytes" % ("ffn" if FFN else "mrconv_fc2", C, M, us, (8.0 if FFN else 16.0) * M * C / us / 1e6))
lib = ctypes.CDLL(os.path.join(ROOT, "neuralsampleid_b200", "libgrafp_sm100a.so"))
if hasattr(lib, "grafp_debug_ffn_trace"):
    buf = (ctypes.c_ulonglong * 3072)()
    assert lib.grafp_debug_ffn_trace(buf) == 0
    b = [int(buf[i]) for i in range(24)]
    t = max(b[6], 1)
    print("tiles %d, MMA warp cycles/tile %.0f; waits: w_full %.0f h_full %.0f acc1_empty %.0f a_full %.0f acc2_empty %.0f; issue+other %.0f"
          % (t, b[0] / t, b[1] / t, b[2] / t, b[3] / t, b[4] / t, b[5] / t, (b[0] - sum(b[1:6])) / t))
    print("TMA A: %.0f/tile, waiting a_empty %.0f" % (b[8] / t, b[9] / t))
    print("transform: %.0f/tile, waiting data %.0f" % (b[10] / t, b[11] / t))
    print("epilogue 1: %.0f/tile, waiting acc1_full %.0f, h_empty %.0f" % (b[12] / t, b[13] / t, b[14] / t))
    print("   tmem ld %.0f, act %.0f, pack+sts %.0f, fence+arrive %.0f" % (b[17] / t, b[18] / t, b[19] / t, b[20] / t))
    print("epilogue 2: %.0f/tile, waiting acc2_full %.0f" % (b[15] / t, b[16] / t))
    print("   tmem ld %.0f, scale+sts %.0f, lds+ldg wait+stg %.0f" % (b[21] / t, b[22] / t, b[23] / t))
