"""Run under torchrun (>= 2 ranks, one GPU each): the data-parallel contrastive train step
(NCCL all-gather of z for global negatives + summed gradient all-reduce) must equal a sequential
single-GPU emulation of the same sharded step (DataParallel semantics: per-shard BatchNorm statistics,
one global NT-Xent, gradients summed).  Prints one line: MULTI_GPU_CHECK PASS|FAIL ..."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import synth                                             # portable weights / inputs only
from neuralsampleid_b200 import ops
from neuralsampleid_b200.autograd import view_bwd, view_fwd
from neuralsampleid_b200.encoder.graph_encoder import GraphEncoder
from neuralsampleid_b200.parallel import shard_range
from neuralsampleid_b200.simclr.simclr import SimCLR
from neuralsampleid_b200.train import FusedClipAdam, train_step

CFG = dict(n_mels=64, n_frames=128, patch_bins=4, patch_frames=8, n_filters=8, tau=0.05,
           d=128, h=1024, u=32, dim=2048, arch="grafp", bsz_train=256, lr=8.0e-5)


def build(dev):
    sd = synth.synth_state(synth.simclr_state_spec(CFG, "t"), 1236)
    model = SimCLR(CFG, encoder=GraphEncoder(cfg=CFG, in_channels=8, k=5))
    model.load_state_dict(sd)
    return model.to(dev).train()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    Bg = 8 * world
    s_i = synth.synth_normal((Bg, 64, 128), 21)
    s_j = s_i + 0.1 * synth.synth_normal((Bg, 64, 128), 22)
    lo, hi = shard_range(Bg, rank, world)
    model = build(dev)
    opt = FusedClipAdam(model.parameters(), lr=CFG["lr"], max_norm=1.0)
    loss = train_step(model, s_i[lo:hi].to(dev), s_j[lo:hi].to(dev), CFG, opt)
    flat = opt.flat_p.clone()
    # every rank must hold identical parameters after the step
    ref = flat.clone()
    dist.broadcast(ref, 0)
    same = bool(torch.allclose(flat, ref, rtol=0, atol=0))
    ok, msg = same, "params identical across ranks: %s" % same
    if rank == 0:
        # sequential emulation on one GPU
        m2 = build(dev)
        o2 = FusedClipAdam(m2.parameters(), lr=CFG["lr"], max_norm=1.0)
        with torch.no_grad():
            ctxs, zs = [], []
            for r in range(world):
                a, b = shard_range(Bg, r, world)
                _, z_i, c_i = view_fwd(m2, s_i[a:b].to(dev))
                _, z_j, c_j = view_fwd(m2, s_j[a:b].to(dev))
                ctxs.append((c_i, c_j))
                zs.append(torch.stack((z_i, z_j), dim=1).reshape(2 * (b - a), -1))
            z_all = torch.cat(zs).contiguous()
            l2, lse = ops.ntxent_fwd(z_all, CFG["tau"])
            dz = ops.ntxent_bwd(z_all, lse, CFG["tau"], torch.ones(1, device=dev)).view(-1, 2, z_all.shape[1])
            grads = {}
            off = 0
            for r in range(world):
                a, b = shard_range(Bg, r, world)
                n = b - a
                view_bwd(m2, ctxs[r][0], None, dz[off:off + n, 0].contiguous(), grads)
                view_bwd(m2, ctxs[r][1], None, dz[off:off + n, 1].contiguous(), grads)
                off += n
            o2.zero_grad()
            o2.accumulate(grads)
            o2.step()
        dl = abs(loss.item() - l2.item())
        # The all-reduced flat gradient must equal the emulation's (atomics make the summation
        # order differ: relative L2 1e-4).  Post-Adam parameters are NOT compared element-wise: the
        # first Adam step moves every weight by lr * sign(grad), and parameters whose true gradient
        # is zero (biases in front of BatchNorm) carry sign noise.  BatchNorm running statistics
        # differ by design (each rank keeps its own, as under nn.DataParallel replicas).
        g1, g2 = opt.flat_g.double(), o2.flat_g.double()
        rel = float((g1 - g2).norm() / g2.norm())
        big = g2.abs() > 1e-3 * float(g2.abs().max())
        upd = float((flat - o2.flat_p)[big].abs().max())
        ok = ok and dl < 1e-4 * abs(l2.item()) and rel < 1e-4 and upd < 0.05 * CFG["lr"]
        msg += "; loss %.6f vs %.6f; grad rel diff %.2e; max update diff on non-noise grads %.2e (lr %.1e); grad norm %.4f vs %.4f" % (
            loss.item(), l2.item(), rel, upd, CFG["lr"], opt.grad_norm(), o2.grad_norm())
        print("MULTI_GPU_CHECK %s world=%d %s" % ("PASS" if ok else "FAIL", world, msg), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
