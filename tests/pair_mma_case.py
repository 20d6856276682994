"""One pre-split f16x3 GEMM chain (fc1 -> split hidden -> fc2 + shortcut, single / split / dual output, ragged M) written to a
file: run by tests/test_gpu_kernels.py::test_gemm_pair_mma_bit_exact in two processes, with and without the CTA-pair kernel
(the kernel selection reads its environment once per process)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neuralsampleid_b200 import ops, _prep
import numpy as np
dev = "cuda:0"
gen = torch.Generator().manual_seed(130)


def normal(shape):
    return torch.randn(shape, generator=gen)


def uniform(shape, lo, hi):
    return torch.rand(shape, generator=gen) * (hi - lo) + lo


out = {}
for (M, k, hid, n) in ((128 * 9 + 37, 256, 1024, 256), (128 * 4, 512, 512, 512), (128 * 7 + 1, 64, 256, 256), (128 * 6 + 5, 128, 512, 256)):
    a = normal((M, k)).to(dev)
    l1 = _prep.make_linear((normal((hid, k)) / float(np.sqrt(k))).to(dev),
                           uniform((hid,), 0.5, 1.5).to(dev), uniform((hid,), -0.5, 0.5).to(dev))
    l2 = _prep.make_linear((normal((n, hid)) / float(np.sqrt(hid))).to(dev), None, None)
    res = normal((M, n)).to(dev)
    hs = ops.linear(a, l1, "relu", 0.0, out_split=True)
    y = ops.linear(hs, l2, None, 0.0, res)
    y2, ys = ops.linear(hs, l2, None, 0.0, res, out_split="both")
    h2 = ops.linear(hs, _prep.make_linear((normal((hid, hid)) / float(np.sqrt(hid))).to(dev), None, None), "gelu", 0.0,
                    out_split=True)
    tag = "%d_%d_%d_%d" % (M, k, hid, n)
    out[tag + "_y"] = y.cpu(); out[tag + "_y2"] = y2.cpu(); out[tag + "_ys"] = ys.t.cpu(); out[tag + "_h2"] = h2.t.cpu()
torch.cuda.synchronize()
import ctypes
lib = ctypes.CDLL(os.path.join(ROOT, "neuralsampleid_b200", "libgrafp_sm100a.so"))
lib.grafp_debug_pair_launches.restype = ctypes.c_longlong
out["pair_launches"] = int(lib.grafp_debug_pair_launches())
torch.save(out, sys.argv[1])
