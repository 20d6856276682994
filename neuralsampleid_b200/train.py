"""The contrastive train step of train.py:48-83 as a fused driver: both views forward in train mode,
NT-Xent over the (global) batch, hand-written backward, gradient all-reduce, clip_grad_norm_(1.0)
and Adam -- every arithmetic step a libgrafp_sm100a kernel, no autograd engine in the loop.
`GraphedTrainStep` captures the whole step into one CUDA graph.

Data parallelism follows nn.DataParallel's semantics (train.py:117-120) with one process per GPU:
the global batch is split on dim 0, BatchNorm statistics are per rank, the loss sees the whole
gathered batch (NCCL all-gather of z), and parameter gradients are SUMMED across ranks (each rank
back-propagates d(global loss)/d(its rows)).
"""
from __future__ import annotations

import os
from typing import Dict, Iterable, Optional

import torch
import torch.distributed as dist

from . import _prep, ops
from .autograd import GradSink, view_bwd_steps, view_fwd


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
        self._lr, self.betas, self.eps, self.max_norm = lr, betas, eps, max_norm
        # step counter and learning rate also live on the device so a CUDA-graph-captured step
        # (GraphedTrainStep) stays valid across replays and LR-schedule changes
        self.step_dev = torch.zeros(1, device=dev, dtype=torch.int32)
        self.lr_dev = torch.full((1,), lr, device=dev, dtype=torch.float32)
        self.sq = torch.zeros(1, device=dev, dtype=torch.float64)
        self.last_grad_norm: Optional[torch.Tensor] = None
        opt = self

        class _Group(dict):                  # a torch-style param group whose 'lr' writes through to the device scalar
            def __setitem__(self, key, value):
                super().__setitem__(key, value)
                if key == "lr":
                    opt._lr = float(value)
                    opt.lr_dev.fill_(float(value))
        self._groups = [_Group(lr=lr, initial_lr=lr, betas=betas, eps=eps, params=self.params)]

    def zero_grad(self, set_to_none: bool = False) -> None:
        self.flat_g.zero_()

    def accumulate(self, grads: Dict[torch.nn.Parameter, torch.Tensor]) -> None:
        for p, g in grads.items():
            view = self.views.get(p)
            if view is not None:
                ops.add_inplace(view, g.reshape(view.shape).contiguous())

    @property
    def lr(self) -> float:
        return self._lr

    @lr.setter
    def lr(self, value: float) -> None:          # e.g. CosineAnnealingLR per epoch (train.py:127)
        self._groups[0]["lr"] = float(value)

    @property
    def step_count(self) -> int:
        return int(self.step_dev.item())

    def step(self, loss_guard: Optional[torch.Tensor] = None) -> None:
        """clip_grad_norm_(max_norm) + Adam.  ``loss_guard``: a device scalar; NaN skips the update."""
        self.check_views()
        self.sq.zero_()
        ops.sq_norm(self.flat_g, self.sq)
        self.last_grad_norm = self.sq
        ops.adam_clip_step_dev(self.flat_p, self.flat_g, self.m, self.v, self.lr_dev, self.betas[0],
                               self.betas[1], self.eps, self.step_dev, self.max_norm, self.sq, loss_guard)
        _prep.bump_weights()        # parameters changed through raw pointers: rebuild prepared weights

    def grad_norm(self) -> float:
        return float(self.sq.sqrt().item())

    # ---- torch.optim.Adam-compatible checkpointing (reference checkpoint format: util.py:149-158 stores
    # optimizer.state_dict(); train.py resumes it) and the param_groups view LR schedulers drive -------------------
    @property
    def param_groups(self):
        return self._groups

    def state_dict(self) -> dict:
        """The layout of torch.optim.Adam.state_dict(): per-parameter 'step', 'exp_avg', 'exp_avg_sq'."""
        state, off = {}, 0
        step = float(self.step_count)
        for i, p in enumerate(self.params):
            k = p.numel()
            state[i] = {"step": torch.tensor(step), "exp_avg": self.m[off:off + k].view(p.shape).clone(),
                        "exp_avg_sq": self.v[off:off + k].view(p.shape).clone()}
            off += k
        group = {"lr": self._lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": 0, "amsgrad": False,
                 "maximize": False, "params": list(range(len(self.params)))}
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, sd: dict) -> None:
        off = 0
        steps = []
        with torch.no_grad():
            for i, p in enumerate(self.params):
                k = p.numel()
                st = sd["state"].get(i)
                if st is not None:
                    self.m[off:off + k].copy_(st["exp_avg"].reshape(-1))
                    self.v[off:off + k].copy_(st["exp_avg_sq"].reshape(-1))
                    steps.append(int(float(st["step"])))
                off += k
            if steps:
                self.step_dev.fill_(max(steps))
        g = sd["param_groups"][0]
        self.betas, self.eps = tuple(g["betas"]), g["eps"]
        self.lr = g["lr"]

    def check_views(self) -> None:
        """Raises if a parameter no longer aliases the flat buffer (model.to(...), p.data reassigned elsewhere)."""
        off = 0
        for p in self.params:
            if p.data_ptr() != self.flat_p.data_ptr() + 4 * off:
                raise RuntimeError("FusedClipAdam: a parameter was detached from the flat buffer (model.to() / p.data "
                                   "reassigned after the optimizer was built); rebuild the optimizer")
            off += p.numel()


class _BucketReducer:
    """Gradient all-reduce (SUM) overlapped with the backward.  The flat gradient buffer is cut at sub-module
    boundaries (projector, proj, every backbone entry, stem, peak extractor); a sub-module's range is ready once BOTH
    views' backward have passed it, ready neighbours are merged, and a merged range of at least ``min_bytes`` is
    all-reduced asynchronously (NCCL runs on its own stream, ordered after the kernels enqueued so far) while the
    backward of the shallower layers continues.  ``finish`` flushes the rest and joins."""

    def __init__(self, optimizer, group, passes: int = 2, min_bytes: int = 4 << 20):
        self.opt, self.group, self.passes, self.min_bytes = optimizer, group, passes, min_bytes
        self.offset = {}
        off = 0
        for p in optimizer.params:
            self.offset[p] = (off, off + p.numel())
            off += p.numel()
        self.seen = {}
        self.ready = []            # disjoint (lo, hi) element ranges, complete and not yet reduced
        self.works = []
        self.ranges = []
        self.launched = 0

    def _range(self, module):
        spans = [self.offset[p] for p in module.parameters() if p in self.offset]
        if not spans:
            return None
        return min(a for a, _ in spans), max(b for _, b in spans)

    def done(self, module) -> None:
        """One view's backward has completed ``module``'s parameter gradients."""
        self.seen[module] = self.seen.get(module, 0) + 1
        if self.seen[module] < self.passes:
            return
        r = self._range(module)
        if r is None:
            return
        self.ready.append(r)
        self.ready.sort()
        merged = [self.ready[0]]
        for a, b in self.ready[1:]:
            if a <= merged[-1][1]:
                merged[-1] = (merged[-1][0], max(b, merged[-1][1]))
            else:
                merged.append((a, b))
        self.ready = merged
        self._launch(False)

    def _launch(self, everything: bool) -> None:
        keep = []
        for a, b in self.ready:
            if everything or (b - a) * 4 >= self.min_bytes:
                self.works.append(dist.all_reduce(self.opt.flat_g[a:b], op=dist.ReduceOp.SUM, group=self.group,
                                                  async_op=True))
                self.ranges.append((a, b))
                self.launched += 1
            else:
                keep.append((a, b))
        self.ready = keep

    def finish(self) -> None:
        self._launch(True)
        self.opt.last_buckets = list(self.ranges)      # (bench bookkeeping: the element ranges that were reduced)
        for w in self.works:
            w.wait()               # the current stream waits for the collective
        self.works = []


def train_step(model, x_i: torch.Tensor, x_j: torch.Tensor, cfg, optimizer: FusedClipAdam,
               group=None, skip_nan: bool = True, forced_idx=None):
    """One step of train.py::train on this rank's shard of the batch.  Returns the (global) loss
    tensor; a NaN loss skips the update like the reference (train.py:65-68).

    Data parallel (one process per GPU): ONE small collective before the loss -- the all-gather of this rank's
    normalised z rows; every rank then evaluates the whole NT-Xent (log-sum-exp of all rows and the scalar loss are
    recomputed redundantly, 2 n^2 D flops, instead of a second all-gather and an all-reduce) -- and the summed
    gradient all-reduce, bucketed and overlapped with the backward (_BucketReducer)."""
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
        loss, lse_all = ops.ntxent_fwd(z_all, tau)             # all n rows: the global loss, on every rank
        one = torch.ones(1, device=z_loc.device, dtype=torch.float32)
        dz = ops.ntxent_bwd(z_all, lse_all, tau, one, rank * rows, rows).view(-1, 2, z_loc.shape[1])
        grads = GradSink(optimizer.views)
        reducer = _BucketReducer(optimizer, group) if world > 1 and not os.environ.get("GRAFP_TRAIN_NO_OVERLAP") else None
        # the two views' backward advance in lockstep, sub-module by sub-module, so that a layer's summed gradient is
        # final -- and its all-reduce can start -- as early as possible
        gi = view_bwd_steps(model, c_i, None, dz[:, 0].contiguous(), grads)
        gj = view_bwd_steps(model, c_j, None, dz[:, 1].contiguous(), grads)
        for mi, mj in zip(gi, gj):
            if reducer is not None:
                reducer.done(mi)
                reducer.done(mj)
        optimizer.accumulate(grads)                            # whatever did not go to the flat buffer directly
        if reducer is not None:
            reducer.finish()
        elif world > 1:
            dist.all_reduce(optimizer.flat_g, op=dist.ReduceOp.SUM, group=group)
        optimizer.step(loss if skip_nan else None)     # a NaN loss skips the update on the device
    return loss.reshape(())


class GraphedTrainStep:
    """One contrastive train step captured into a CUDA graph (every arithmetic node is a
    libgrafp_sm100a kernel; the NCCL all-gather / all-reduce are captured too).  At the reference's
    per-GPU batch (32 pairs) the eager step is bound by ~2000 host-side launches; replaying the
    graph leaves only the kernels.

        g = GraphedTrainStep(model, cfg, optimizer, pairs=32)
        loss = g(x_i, x_j)            # copies the inputs into the static buffers and replays
    """

    def __init__(self, model, cfg, optimizer: FusedClipAdam, pairs: int, group=None, warmup: int = 2):
        if not model.training:
            raise RuntimeError("GraphedTrainStep needs model.train()")
        dev = optimizer.flat_p.device
        self.model, self.optimizer = model, optimizer
        self.x_i = torch.zeros((pairs, cfg["n_mels"], cfg["n_frames"]), device=dev, dtype=torch.float32)
        self.x_j = torch.zeros_like(self.x_i)
        # warm-up steps run for real: snapshot and restore every piece of mutable state
        keep = [optimizer.flat_p.clone(), optimizer.m.clone(), optimizer.v.clone(), optimizer.step_dev.clone()]
        bufs = [(b, b.clone()) for b in model.buffers()]
        self.x_i.normal_(generator=torch.Generator(device=dev).manual_seed(1))
        self.x_j.copy_(self.x_i).add_(0.1)
        stream = torch.cuda.Stream(device=dev)
        stream.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(stream):
            for _ in range(warmup):
                train_step(model, self.x_i, self.x_j, cfg, optimizer, group)
        torch.cuda.current_stream(dev).wait_stream(stream)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = train_step(model, self.x_i, self.x_j, cfg, optimizer, group)
        with torch.no_grad():
            optimizer.flat_p.copy_(keep[0]); optimizer.m.copy_(keep[1]); optimizer.v.copy_(keep[2])
            optimizer.step_dev.copy_(keep[3])
            for b, saved in bufs:
                b.copy_(saved)
        _prep.bump_weights()

    def __call__(self, x_i: torch.Tensor, x_j: torch.Tensor) -> torch.Tensor:
        self.x_i.copy_(x_i, non_blocking=True)
        self.x_j.copy_(x_j, non_blocking=True)
        self.graph.replay()
        _prep.bump_weights()
        return self.loss
