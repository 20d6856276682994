"""The contrastive train step of train.py:48-83 as a fused driver: both views forward in train mode,
NT-Xent over the (global) batch, hand-written backward, gradient all-reduce, clip_grad_norm_(1.0)
and Adam -- every arithmetic step a libgrafp_sm100a kernel, no autograd engine in the loop.

Data parallelism follows nn.DataParallel's semantics (train.py:117-120) with one process per GPU:
the global batch is split on dim 0, BatchNorm statistics are per rank, the loss sees the whole
gathered batch (NCCL all-gather of z), and parameter gradients are SUMMED across ranks (each rank
back-propagates d(global loss)/d(its rows)).
"""
from __future__ import annotations

from typing import Dict, Iterable, Optional

import torch
import torch.distributed as dist

from . import ops
from .autograd import view_bwd, view_fwd


class FusedClipAdam:
    """torch.optim.Adam(lr) + clip_grad_norm_(max_norm) over ONE flat parameter buffer: two kernels
    per step (squared-norm reduction, fused clip + Adam update).  Parameters are re-pointed at views
    of the flat buffer, so the module / state_dict see the updates."""

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 8e-5, betas=(0.9, 0.999),
                 eps: float = 1e-8, max_norm: float = 1.0):
        self.params = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("no trainable parameters")
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat_p = torch.empty(n, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(n, device=dev, dtype=torch.float32)
        self.m = torch.zeros(n, device=dev, dtype=torch.float32)
        self.v = torch.zeros(n, device=dev, dtype=torch.float32)
        self.views: Dict[torch.nn.Parameter, torch.Tensor] = {}
        off = 0
        with torch.no_grad():
            for p in self.params:
                k = p.numel()
                self.flat_p[off:off + k].copy_(p.detach().reshape(-1))
                p.data = self.flat_p[off:off + k].view(p.shape)
                gview = self.flat_g[off:off + k].view(p.shape)
                p.grad = gview
                self.views[p] = gview
                off += k
        self.lr, self.betas, self.eps, self.max_norm = lr, betas, eps, max_norm
        self.step_count = 0
        self.sq = torch.zeros(1, device=dev, dtype=torch.float64)
        self.last_grad_norm: Optional[torch.Tensor] = None

    def zero_grad(self, set_to_none: bool = False) -> None:
        self.flat_g.zero_()

    def accumulate(self, grads: Dict[torch.nn.Parameter, torch.Tensor]) -> None:
        for p, g in grads.items():
            view = self.views.get(p)
            if view is not None:
                ops.add_inplace(view, g.reshape(view.shape).contiguous())

    def step(self) -> None:
        self.step_count += 1
        self.sq.zero_()
        ops.sq_norm(self.flat_g, self.sq)
        self.last_grad_norm = self.sq
        ops.adam_clip_step(self.flat_p, self.flat_g, self.m, self.v, self.lr, self.betas[0], self.betas[1],
                           self.eps, self.step_count, self.max_norm, self.sq)
        # bump the version counters so prepared (folded / split) weights are rebuilt
        for p in self.params:
            p.data = p.data

    def grad_norm(self) -> float:
        return float(self.sq.sqrt().item())


def train_step(model, x_i: torch.Tensor, x_j: torch.Tensor, cfg, optimizer: FusedClipAdam,
               group=None, skip_nan: bool = True, forced_idx=None):
    """One step of train.py::train on this rank's shard of the batch.  Returns the (global) loss
    tensor; a NaN loss skips the update like the reference (train.py:65-68)."""
    if not model.training:
        raise RuntimeError("train_step needs model.train()")
    tau = float(cfg["tau"])
    world = dist.get_world_size(group) if (dist.is_initialized()) else 1
    rank = dist.get_rank(group) if world > 1 else 0
    optimizer.zero_grad()
    with torch.no_grad():
        h_i, z_i, c_i = view_fwd(model, x_i, forced_idx[0] if forced_idx else None)
        h_j, z_j, c_j = view_fwd(model, x_j, forced_idx[1] if forced_idx else None)
        z_loc = torch.stack((z_i, z_j), dim=1).reshape(2 * z_i.shape[0], z_i.shape[1]).contiguous()
        rows = z_loc.shape[0]
        if world > 1:
            z_all = torch.empty((world * rows, z_loc.shape[1]), device=z_loc.device, dtype=torch.float32)
            dist.all_gather_into_tensor(z_all, z_loc, group=group)
        else:
            z_all = z_loc
        loss, lse = ops.ntxent_fwd(z_all, tau, rank * rows, rows)
        if world > 1:
            lse_all = torch.empty((world * rows,), device=z_loc.device, dtype=torch.float32)
            dist.all_gather_into_tensor(lse_all, lse, group=group)
            dist.all_reduce(loss, group=group)
        else:
            lse_all = lse
        if skip_nan and bool(torch.isnan(loss).item()):
            return loss.reshape(())
        one = torch.ones(1, device=z_loc.device, dtype=torch.float32)
        dz = ops.ntxent_bwd(z_all, lse_all, tau, one, rank * rows, rows).view(-1, 2, z_loc.shape[1])
        grads: Dict = {}
        view_bwd(model, c_i, None, dz[:, 0].contiguous(), grads)
        view_bwd(model, c_j, None, dz[:, 1].contiguous(), grads)
        optimizer.accumulate(grads)
        if world > 1:
            dist.all_reduce(optimizer.flat_g, op=dist.ReduceOp.SUM, group=group)
        optimizer.step()
    return loss.reshape(())
