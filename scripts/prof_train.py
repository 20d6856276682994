"""Per-entry-point device time of one contrastive train step (eager, CUDA events around every C-ABI call) next to the
CUDA-graph replay time of the whole step.  usage: python scripts/prof_train.py [pairs]"""
import json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from neuralsampleid_b200 import _lib
from neuralsampleid_b200.train import FusedClipAdam, GraphedTrainStep, train_step
sys.path.insert(0, ROOT)
import bench_extra as BE

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = "cuda:0"
model = BE._model(dev, 5).train()
opt = FusedClipAdam(model.parameters(), lr=BE.CFG["lr"], max_norm=1.0)
g = torch.Generator().manual_seed(2)
x_i = torch.randn((pairs, 64, 128), generator=g).to(dev)
x_j = (x_i.cpu() + 0.1 * torch.randn((pairs, 64, 128), generator=g)).to(dev)
with torch.no_grad():
    for _ in range(3):
        train_step(model, x_i, x_j, BE.CFG, opt)
    rec = []

    def timer(name, fn, args):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        rec.append((name, e0, e1))
        return rc
    torch.cuda.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    _lib.set_profiler(timer)
    t0.record()
    train_step(model, x_i, x_j, BE.CFG, opt)
    t1.record()
    _lib.set_profiler(None)
    torch.cuda.synchronize()
    agg = {}
    for name, e0, e1 in rec:
        a = agg.setdefault(name, [0.0, 0])
        a[0] += e0.elapsed_time(e1); a[1] += 1
    gs = GraphedTrainStep(model, BE.CFG, opt, pairs)
    for _ in range(2):
        gs(x_i, x_j)
    torch.cuda.synchronize()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(10):
        gs(x_i, x_j)
    g1.record()
    torch.cuda.synchronize()
print(json.dumps({"pairs": pairs, "eager_instrumented_ms": t0.elapsed_time(t1), "graph_ms": g0.elapsed_time(g1) / 10,
                  "calls": sum(a[1] for a in agg.values()), "sum_call_ms": sum(a[0] for a in agg.values()),
                  "by_entry": {k: {"ms": round(v[0], 3), "calls": v[1], "us_per_call": round(1e3 * v[0] / v[1], 1)}
                               for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])}}))
