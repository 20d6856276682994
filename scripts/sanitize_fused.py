"""Small invocations of the round-2 kernels for compute-sanitizer (memcheck): fused FFN, fused MRConv -> fc2, peak
extractor node kernel.  usage: compute-sanitizer --tool memcheck python scripts/sanitize_fused.py"""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neuralsampleid_b200 import ops, _prep
dev = "cuda:0"
torch.manual_seed(0)
for C, M in ((64, 128 * 3 + 5), (128, 128 * 2 + 77)):
    x, m, res = torch.randn(M, C, device=dev), torch.randn(M, C, device=dev).abs(), torch.randn(M, C, device=dev)
    l1 = _prep.make_linear(torch.randn(2 * C, C // 2, device=dev), torch.ones(2 * C, device=dev), torch.zeros(2 * C, device=dev), 4, dual=True)
    l2 = _prep.make_linear(torch.randn(C, 2 * C, device=dev), torch.ones(C, device=dev), torch.zeros(C, device=dev))
    y = ops.mrconv_fc2_fused(x, m, l1, "relu", 0.0, l2, res)
    f1 = _prep.make_linear(torch.randn(4 * C, C, device=dev), torch.ones(4 * C, device=dev), torch.zeros(4 * C, device=dev))
    f2 = _prep.make_linear(torch.randn(C, 4 * C, device=dev), torch.ones(C, device=dev), torch.zeros(C, device=dev))
    z = ops.ffn_fused(x, f1, f2, "gelu")
    torch.cuda.synchronize()
    print("C", C, "ok", float(y.abs().mean()), float(z.abs().mean()))
s = torch.randn(7, 64, 128, device=dev)
o = ops.peak_extract(s, torch.randn(8, 3, 4, 8, device=dev), torch.randn(8, device=dev))
torch.cuda.synchronize()
print("peak ok", tuple(o.shape))
