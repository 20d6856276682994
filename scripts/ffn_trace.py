"""Debug: where the fused FFN kernel's MMA-issuing thread spends its cycles (build with GRAFP_NVCC_EXTRA=-DFF_TRACE)."""
import ctypes, os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neuralsampleid_b200 import ops, _prep
C = int(sys.argv[1]); M = int(sys.argv[2])
dev = "cuda:0"
torch.manual_seed(0)
x = torch.randn(M, C, device=dev)
l1 = _prep.make_linear(torch.randn(4 * C, C, device=dev) / C ** 0.5, torch.ones(4 * C, device=dev), torch.zeros(4 * C, device=dev))
l2 = _prep.make_linear(torch.randn(C, 4 * C, device=dev) / (4 * C) ** 0.5, torch.ones(C, device=dev), torch.zeros(C, device=dev))
for _ in range(2): y = ops.ffn_fused(x, l1, l2, "relu")
torch.cuda.synchronize()
lib = ctypes.CDLL(os.path.join(ROOT, "neuralsampleid_b200", "libgrafp_sm100a.so"))
buf = (ctypes.c_ulonglong * 3072)()
assert lib.grafp_debug_ffn_trace(buf) == 0
tot, w, h, a, xw, a2, tiles = [int(buf[i]) for i in range(7)]
print("C %d: tiles %d, cycles/tile %.0f; waits per tile: w_full %.0f  h_full %.0f  acc1_empty %.0f  xop_full %.0f  acc2_empty %.0f; issue+other %.0f"
      % (C, tiles, tot / tiles, w / tiles, h / tiles, a / tiles, xw / tiles, a2 / tiles, (tot - w - h - a - xw - a2) / tiles))
