"""B200 counterpart of the reference's simclr/ntxent.py (``ntxent_loss``, :5-30): one fused
similarity-GEMM + masked log-sum-exp kernel forward and one fused backward instead of a Python
loop over the 2B rows.  ``ntxent_loss_distributed`` is the data-parallel form: an NCCL all-gather
of the local embeddings provides the global negatives (what nn.DataParallel's gather to GPU 0
gives the reference, train.py:61-63,117-120), and the backward returns only the local slice of
the gradient of the global mean loss."""
from __future__ import annotations

import torch
import torch.distributed as dist

from .. import ops


class _NTXent(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z_i, z_j, tau):
        z = torch.stack((z_i, z_j), dim=1).reshape(2 * z_i.shape[0], z_i.shape[1]).contiguous()
        loss, lse = ops.ntxent_fwd(z, tau)
        ctx.save_for_backward(z, lse)
        ctx.tau = tau
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        z, lse = ctx.saved_tensors
        dz = ops.ntxent_bwd(z, lse, ctx.tau, g.contiguous())
        dz = dz.view(-1, 2, z.shape[1])
        return dz[:, 0].contiguous(), dz[:, 1].contiguous(), None


def ntxent_loss(z_i, z_j, cfg):
    """z_i, z_j: (B, D) L2-normalised embeddings of the two views -> scalar loss."""
    return _NTXent.apply(z_i.float(), z_j.float(), float(cfg["tau"]))


class _NTXentDist(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z_i, z_j, tau, group):
        world = dist.get_world_size(group)
        rank = dist.get_rank(group)
        z_loc = torch.stack((z_i, z_j), dim=1).reshape(2 * z_i.shape[0], z_i.shape[1]).contiguous()
        rows = z_loc.shape[0]
        z_all = torch.empty((world * rows, z_loc.shape[1]), device=z_loc.device, dtype=z_loc.dtype)
        dist.all_gather_into_tensor(z_all, z_loc, group=group)
        loss, lse = ops.ntxent_fwd(z_all, tau, rank * rows, rows)
        lse_all = torch.empty((world * rows,), device=z_loc.device, dtype=torch.float32)
        dist.all_gather_into_tensor(lse_all, lse, group=group)
        dist.all_reduce(loss, group=group)           # every rank reports the global mean loss
        ctx.save_for_backward(z_all, lse_all)
        ctx.tau, ctx.row0, ctx.rows = tau, rank * rows, rows
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        z_all, lse_all = ctx.saved_tensors
        # d(global loss)/d(local rows): already includes the terms where local rows act as
        # negatives of remote rows, so parameter gradients must be SUMMED across ranks.
        dz = ops.ntxent_bwd(z_all, lse_all, ctx.tau, g.contiguous(), ctx.row0, ctx.rows)
        dz = dz.view(-1, 2, z_all.shape[1])
        return dz[:, 0].contiguous(), dz[:, 1].contiguous(), None, None


def ntxent_loss_distributed(z_i, z_j, cfg, group=None):
    """Global-batch NT-Xent over all ranks of ``group`` (rank r holds pairs [r*B_loc, (r+1)*B_loc))."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return ntxent_loss(z_i, z_j, cfg)
    return _NTXentDist.apply(z_i.float(), z_j.float(), float(cfg["tau"]), group)
