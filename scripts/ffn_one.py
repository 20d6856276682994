import os, sys, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neuralsampleid_b200 import ops, _prep
C = int(sys.argv[1]); M = int(sys.argv[2])
dev = "cuda:0"
torch.manual_seed(0)
x = torch.randn(M, C, device=dev)
l1 = _prep.make_linear(torch.randn(4 * C, C, device=dev) / C ** 0.5, torch.ones(4 * C, device=dev), torch.zeros(4 * C, device=dev))
l2 = _prep.make_linear(torch.randn(C, 4 * C, device=dev) / (4 * C) ** 0.5, torch.ones(C, device=dev), torch.zeros(C, device=dev))
for _ in range(3): y = ops.ffn_fused(x, l1, l2, "relu")
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): y = ops.ffn_fused(x, l1, l2, "relu")
e1.record(); torch.cuda.synchronize()
print("C", C, "M", M, "fused us", 200 * e0.elapsed_time(e1))
