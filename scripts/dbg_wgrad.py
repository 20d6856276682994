import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neuralsampleid_b200 import ops, _lib
torch.manual_seed(0)
M, n, k = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
dy = torch.randn(M, n, device="cuda"); a = torch.randn(M, k, device="cuda")
want = (dy.double().T @ a.double())
got = ops.gemm_wgrad(dy, a, None, n, 1, 0, engine=_lib.ENGINES["3xtf32"]).double()
print("want absmax", float(want.abs().max()), "got absmax", float(got.abs().max()), "nonzero frac", float((got != 0).float().mean()))
print("err", float((got - want).abs().max()))
# structure probes: dy = one-hot rows/cols
for name, (dyp, ap) in {
    "dy=e(m0,n3), a=e(m0,k5)": (torch.zeros(M, n, device="cuda").index_put_((torch.tensor([0]), torch.tensor([3])), torch.tensor(1.0)),
                                 torch.zeros(M, k, device="cuda").index_put_((torch.tensor([0]), torch.tensor([5])), torch.tensor(1.0))),
    "dy=e(m9,n40), a=e(m9,k33)": (torch.zeros(M, n, device="cuda").index_put_((torch.tensor([9]), torch.tensor([40])), torch.tensor(1.0)),
                                   torch.zeros(M, k, device="cuda").index_put_((torch.tensor([9]), torch.tensor([33])), torch.tensor(1.0))),
}.items():
    g = ops.gemm_wgrad(dyp, ap, None, n, 1, 0, engine=_lib.ENGINES["3xtf32"])
    nz = torch.nonzero(g)
    print(name, "->", nz[:8].tolist(), [float(g[i, j]) for i, j in nz[:8].tolist()])
