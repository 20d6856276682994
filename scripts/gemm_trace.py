"""Debug: cycle accounting of gemm_tc's MMA warp (block 0) on the step's big GEMM shapes.
Build with GRAFP_NVCC_EXTRA=-DTC_TRACE first."""
import ctypes, os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neuralsampleid_b200 import ops, _prep
dev = "cuda:0"
torch.manual_seed(0)
lib = ctypes.CDLL(os.path.join(ROOT, "neuralsampleid_b200", "libgrafp_sm100a.so"))
buf = (ctypes.c_ulonglong * 24)()

def lin(n, k):
    return _prep.make_linear(torch.randn(n, k, device=dev) / k ** 0.5, torch.ones(n, device=dev), torch.zeros(n, device=dev))

def run(tag, fn):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    assert lib.grafp_debug_tc_trace(buf) == 0
    tot, ce, cw, ca, tiles, nkb, S, RAW = [int(b) for b in buf][:8]
    et = [int(b) for b in buf][8:19]
    kb = max(1, tiles * nkb)
    print("%-34s %7.1f us | block 0: %3d tiles x %2d k-blocks, W stages %d, A ring %d | cycles per k-block %5.0f: wait tmem_empty %4.0f, W %4.0f, A %4.0f, issue+rest %4.0f"
          % (tag, 1e3 * e0.elapsed_time(e1), tiles, nkb, S, RAW, tot / kb, ce / kb, cw / kb, ca / kb, (tot - ce - cw - ca) / kb), flush=True)
    ch = max(1, et[10])
    names = ["tile-start barrier", "wait tmem_full", "tmem ld", "wait store read", "barrier A", "pack + st.shared", "fence", "barrier B", "TMA store issue"]
    per_tile = ", ".join("%s %.0f" % (n, et[1 + i] / max(1, tiles)) for i, n in enumerate(names[:2]))
    per_chunk = ", ".join("%s %.0f" % (n, et[1 + i] / ch) for i, n in enumerate(names) if i >= 2)
    acc = sum(et[1:10])
    print("      epilogue warp 2: %.0f cycles per tile (%d chunks of mine per tile); per tile: %s; per chunk: %s, rest (residual loads, apply) %.0f"
          % (et[0] / max(1, tiles), ch // max(1, tiles), per_tile, per_chunk, (et[0] - acc) / ch), flush=True)

for (M, C) in ((262144, 256), (131072, 512)):
    x = torch.randn(M, C, device=dev)
    ident = lin(C, C)
    xs = ops.linear(x, ident, out_split=True)            # a SplitAct, like the Grapher fc2 dual output
    fc1, fc2 = lin(4 * C, C), lin(C, 4 * C)
    h = ops.linear(xs, fc1, "gelu", 0.0, out_split=True)
    run("fc1 split-in gelu split-out C=%d" % C, lambda: ops.linear(xs, fc1, "gelu", 0.0, out_split=True))
    run("fc1 split-in NO act split-out C=%d" % C, lambda: ops.linear(xs, fc1, out_split=True))
    run("fc1 split-in gelu fp32-out C=%d" % C, lambda: ops.linear(xs, fc1, "gelu", 0.0))
    run("fc1 split-in NO act fp32-out C=%d" % C, lambda: ops.linear(xs, fc1))
    run("fc1 split-in relu split-out C=%d" % C, lambda: ops.linear(xs, fc1, "relu", 0.0, out_split=True))
    run("fc1 fp32-in gelu split-out C=%d" % C, lambda: ops.linear(x, fc1, "gelu", 0.0, out_split=True))
    run("fc2 split-in + residual C=%d" % C, lambda: ops.linear(h, fc2, residual=x))
    run("fc2 split-in no residual C=%d" % C, lambda: ops.linear(h, fc2))
    g2 = lin(C, 2 * C)
    m = ops.linear(x, lin(2 * C, C), out_split=True)
    run("grapher fc2 split-in + res C=%d" % C, lambda: ops.linear(m, g2, residual=x))
