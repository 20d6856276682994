import torch
x = torch.empty(1 << 30, dtype=torch.float32, device="cuda")   # 4 GB
y = torch.empty_like(x)
def t(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e-3
tw = t(lambda: x.zero_()); print("pure write  %.0f GB/s" % (x.numel() * 4 / tw / 1e9))
tr = t(lambda: x.sum()); print("pure read   %.0f GB/s" % (x.numel() * 4 / tr / 1e9))
tc = t(lambda: y.copy_(x)); print("copy (r+w)  %.0f GB/s" % (2 * x.numel() * 4 / tc / 1e9))
row = torch.randn(1 << 18, device="cuda")                      # 1 MB, L2-resident source
xv = x.view(-1, 1 << 18)
tw2 = t(lambda: xv.copy_(row.expand_as(xv))); print("write of incompressible data (source in L2)  %.0f GB/s" % (x.numel() * 4 / tw2 / 1e9))
tf = t(lambda: x.fill_(1.5)); print("fill(const) %.0f GB/s" % (x.numel() * 4 / tf / 1e9))
