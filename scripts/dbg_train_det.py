import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import synth
from neuralsampleid_b200 import ops
from neuralsampleid_b200.autograd import view_bwd, view_fwd
from neuralsampleid_b200.encoder.graph_encoder import GraphEncoder
from neuralsampleid_b200.simclr.simclr import SimCLR
from neuralsampleid_b200.train import FusedClipAdam, train_step
CFG = dict(n_mels=64, n_frames=128, patch_bins=4, patch_frames=8, n_filters=8, tau=0.05, d=128, h=1024, u=32, dim=2048, arch="grafp", bsz_train=256, lr=8.0e-5)
dev = "cuda:0"
def build():
    sd = synth.synth_state(synth.simclr_state_spec(CFG, "t"), 1236)
    m = SimCLR(CFG, encoder=GraphEncoder(cfg=CFG, in_channels=8, k=5)); m.load_state_dict(sd); return m.to(dev).train()
s_i = synth.synth_normal((8, 64, 128), 21).to(dev); s_j = (s_i.cpu() + 0.1 * synth.synth_normal((8, 64, 128), 22)).to(dev)
def run_sink():
    m = build(); o = FusedClipAdam(m.parameters(), lr=CFG["lr"]); train_step(m, s_i, s_j, CFG, o); return o.flat_g.clone(), m
def run_dict():
    m = build(); o = FusedClipAdam(m.parameters(), lr=CFG["lr"])
    with torch.no_grad():
        _, z_i, c_i = view_fwd(m, s_i); _, z_j, c_j = view_fwd(m, s_j)
        z = torch.stack((z_i, z_j), dim=1).reshape(16, -1).contiguous()
        l, lse = ops.ntxent_fwd(z, CFG["tau"]); dz = ops.ntxent_bwd(z, lse, CFG["tau"], torch.ones(1, device=dev)).view(-1, 2, z.shape[1])
        g = {}; view_bwd(m, c_i, None, dz[:, 0].contiguous(), g); view_bwd(m, c_j, None, dz[:, 1].contiguous(), g)
        o.zero_grad(); o.accumulate(g)
    return o.flat_g.clone(), m
a1, m1 = run_sink(); a2, _ = run_sink(); b1, m2 = run_dict(); b2, _ = run_dict()
rel = lambda x, y: float((x.double() - y.double()).norm() / y.double().norm())
print("sink vs sink %.2e  dict vs dict %.2e  sink vs dict %.2e" % (rel(a1, a2), rel(b1, b2), rel(a1, b1)))
# per-parameter breakdown of sink vs dict
off = 0
worst = []
for (n, p) in m1.named_parameters():
    if not p.requires_grad: continue
    k = p.numel(); x, y = a1[off:off + k].double(), b1[off:off + k].double(); off += k
    worst.append((float((x - y).norm() / (y.norm() + 1e-30)), n, float(y.norm())))
tot = float((a1.double() - a2.double()).norm())
off = 0; contrib = []
for (n, p) in m1.named_parameters():
    if not p.requires_grad: continue
    k = p.numel(); x, y = a1[off:off + k].double(), a2[off:off + k].double(); off += k
    contrib.append((float((x - y).norm()) / tot, n, float(y.norm()), float((x - y).norm() / (y.norm() + 1e-30))))
contrib.sort(reverse=True)
print("sink run 1 vs run 2: share of the total difference, name, |g|, relative diff")
for w in contrib[:14]: print("%.3f %-50s |g|=%.3e rel %.2e" % w)
