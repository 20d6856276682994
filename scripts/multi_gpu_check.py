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
    with torch.no_grad():                                # this rank's z rows under the pre-step weights (a fresh copy:
        m0 = build(dev)                                  # the train-mode forward also moves BatchNorm statistics)
        _, z_a, _ = view_fwd(m0, s_i[lo:hi].to(dev))
        _, z_b, _ = view_fwd(m0, s_j[lo:hi].to(dev))
        z_rank0 = torch.stack((z_a, z_b), dim=1).reshape(2 * (hi - lo), -1).clone()
        del m0
    z_gather = torch.empty((world * z_rank0.shape[0], z_rank0.shape[1]), device=dev)
    dist.all_gather_into_tensor(z_gather, z_rank0)       # (test bookkeeping) every rank's pre-step z rows
    loss = train_step(model, s_i[lo:hi].to(dev), s_j[lo:hi].to(dev), CFG, opt)
    # ---- teacher-forced N-rank step for the ORACLE comparison: train-mode near-tie neighbour flips cascade through
    # the per-shard BatchNorm statistics (8 of 32 z rows moved by > 1e-3 in a free-running 2-rank run), so the
    # oracle's own per-block graphs of this rank's shard are forced, as in tests/test_gpu_train.py
    from oracle import grafp_oracle as O
    sd = synth.synth_state(synth.simclr_state_spec(CFG, "t"), 1236)

    def oracle_shard(params, a, b):
        enc = {n[len("encoder."):]: t for n, t in params.items() if n.startswith("encoder.")}
        zs, forced = [], []
        for x in (s_i[a:b], s_j[a:b]):
            taps = []
            h = O.encoder_forward(enc, O.peak_extractor(params, x), k=5, training=True, stats={}, taps=taps)
            zs.append(O.projector(params, h))
            forced.append([t["idx"].int() for t in taps if t["kind"] == "block"])
        return zs, forced
    with torch.no_grad():
        _, forced = oracle_shard({n: t.clone() for n, t in sd.items()}, lo, hi)
    forced_dev = tuple([t.to(dev) for t in f] for f in forced)
    model_f = build(dev)
    opt_f = FusedClipAdam(model_f.parameters(), lr=CFG["lr"], max_norm=1.0)
    with torch.no_grad():
        m0 = build(dev)
        _, z_a, _ = view_fwd(m0, s_i[lo:hi].to(dev), forced_dev[0])
        _, z_b, _ = view_fwd(m0, s_j[lo:hi].to(dev), forced_dev[1])
        z_forced = torch.stack((z_a, z_b), dim=1).reshape(2 * (hi - lo), -1).clone()
        del m0
    zf_gather = torch.empty((world * z_forced.shape[0], z_forced.shape[1]), device=dev)
    dist.all_gather_into_tensor(zf_gather, z_forced)
    loss_f = train_step(model_f, s_i[lo:hi].to(dev), s_j[lo:hi].to(dev), CFG, opt_f, forced_idx=forced_dev)
    flat_f = opt_f.flat_p.clone()
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
        # The all-reduced flat gradient must equal the emulation's (relative L2 1e-4).  Post-Adam parameters are NOT
        # compared element-wise: the first Adam step moves every weight by lr * sign(grad), and parameters whose true
        # gradient is zero (biases in front of BatchNorm) carry sign noise.  BatchNorm running statistics differ by
        # design (each rank keeps its own, as under nn.DataParallel replicas).
        g1, g2 = opt.flat_g.double(), o2.flat_g.double()
        rel = float((g1 - g2).norm() / g2.norm())
        big = g2.abs() > 1e-3 * float(g2.abs().max())
        upd = float((flat - o2.flat_p)[big].abs().max())
        ok = ok and dl < 1e-4 * abs(l2.item()) and rel < 1e-4 and upd < 0.05 * CFG["lr"]
        msg += "; loss %.6f vs %.6f; grad rel diff %.2e; max update diff on non-noise grads %.2e (lr %.1e); grad norm %.4f vs %.4f" % (
            loss.item(), l2.item(), rel, upd, CFG["lr"], opt.grad_norm(), o2.grad_norm())
        # ---- the ORACLE's DataParallel emulation (reference train.py:48-83, 117-120): every
        # shard through the CPU restatement of the reference in train mode (per-replica BatchNorm statistics), z
        # concatenated, ONE global NT-Xent, torch autograd backward (gradients of the shared parameters sum over
        # the shards) -- compared with the teacher-forced N-rank NCCL step: loss, every rank's z rows, the global
        # gradient norm, the gradient direction, and the applied clip + Adam update.
        names = [n for n, p in m2.named_parameters() if p.requires_grad]
        params = {n: (t.clone().requires_grad_(True) if n in names else t.clone()) for n, t in sd.items()}
        zi, zj = [], []
        for r in range(world):
            a, b = shard_range(Bg, r, world)
            zs, _ = oracle_shard(params, a, b)
            zi.append(zs[0])
            zj.append(zs[1])
        loss_o = O.ntxent(torch.cat(zi), torch.cat(zj), CFG["tau"])
        loss_o.backward()
        g_o = torch.cat([params[n].grad.reshape(-1) for n in names]).double()
        g_n = opt_f.flat_g.detach().cpu().double()
        cos = float((g_o * g_n).sum() / (g_o.norm() * g_n.norm()))
        nrm = abs(float(g_n.norm()) - float(g_o.norm())) / float(g_o.norm())
        dlo = abs(loss_f.item() - loss_o.item()) / abs(loss_o.item())
        z_o = torch.stack((torch.cat(zi), torch.cat(zj)), dim=1).reshape(2 * Bg, -1).detach()
        zr = ((zf_gather.cpu() - z_o).norm(dim=1) / z_o.norm(dim=1))
        zfree = ((z_gather.cpu() - z_o).norm(dim=1) / z_o.norm(dim=1))
        # clip_grad_norm_(1.0) + Adam on the oracle's gradients: the applied update agrees where sign(grad) is defined
        glist = [params[n].grad.clone() for n in names]
        O.clip_grad_norm_(glist, 1.0)
        agree = count = 0
        flat_new = flat_f.detach().cpu()
        off = 0
        for n, g_ in zip(names, glist):
            p_ = sd[n].clone()
            O.adam_step(p_, g_, torch.zeros_like(p_), torch.zeros_like(p_), 1, CFG["lr"])
            k_ = p_.numel()
            upd_ref = (p_ - sd[n]).reshape(-1)
            upd_n = flat_new[off:off + k_] - sd[n].reshape(-1)
            bigm = g_.reshape(-1).abs() > 1e-6 * float(g_.abs().max() + 1e-30)
            agree += int(((upd_n - upd_ref).abs() < 0.05 * CFG["lr"])[bigm].sum())
            count += int(bigm.sum())
            off += k_
        ok_o = dlo < 1e-3 and cos > 0.999 and nrm < 2e-2 and float(zr.max()) < 1e-3 and agree > 0.99 * count
        ok = ok and ok_o
        msg += "; ORACLE DataParallel emulation (graphs forced): loss %.6f vs %.6f (rel %.2e), grad cosine %.6f, grad norm " \
               "rel %.2e, z rel max %.2e, update agreement %.4f; free-running z rows off by > 1e-3: %d of %d (train-mode " \
               "tie flips); gradient all-reduce %s" % (
                   loss_f.item(), loss_o.item(), dlo, cos, nrm, float(zr.max()), agree / max(1, count),
                   int((zfree > 1e-3).sum()), zfree.numel(),
                   "unbucketed" if os.environ.get("GRAFP_TRAIN_NO_OVERLAP") else "bucketed, overlapped with the backward")
        print("MULTI_GPU_CHECK %s world=%d %s" % ("PASS" if ok else "FAIL", world, msg), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
