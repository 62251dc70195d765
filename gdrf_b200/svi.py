"""Minimal stand-in for ``pyro.infer.SVI`` around the fused ELBO op, with the call shape the reference's
training loop uses (``train_script.py:365-371,467``): ``SVI(model, guide, optim, loss).step(xs=, ws=,
subsample=False)`` returns the loss as a Python float after one optimiser update.  Observation-sharded
data parallelism: every rank evaluates its shard, the small parameter gradients are all-reduced.
"""
from __future__ import annotations

from typing import Optional

import torch
import torch.distributed as dist


def allreduce_grads_(params, group=None) -> None:
    """One flat all-reduce (sum) of every parameter gradient -- NCCL over NVLink on the GPU box, gloo in the
    CPU tests."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    o = 0
    for g in grads:
        g.copy_(flat[o:o + g.numel()].view_as(g))
        o += g.numel()


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of n observations for `rank`."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class SVI:
    def __init__(self, model, guide=None, optim: Optional[torch.optim.Optimizer] = None, loss=None, group=None):
        self.module = getattr(model, "__self__", model)
        self.optim = optim
        self.group = group

    def loss_and_grads(self, xs, ws, eps=None, n_global=None, n_offset=0) -> torch.Tensor:
        rank = dist.get_rank(self.group) if dist.is_available() and dist.is_initialized() else 0
        elbo = self.module.elbo(xs, ws, eps=eps, n_global=n_global, n_offset=n_offset, include_prior=(rank == 0))
        loss = -elbo
        loss.backward()
        params = [p for p in self.module.parameters() if p.requires_grad]
        allreduce_grads_(params, self.group)
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(self.group) > 1:
            loss = loss.detach().clone()
            dist.all_reduce(loss, op=dist.ReduceOp.SUM, group=self.group)
        return loss.detach()

    def step(self, xs, ws, subsample=False, eps=None, n_global=None, n_offset=0) -> float:
        if self.optim is not None:
            self.optim.zero_grad(set_to_none=True)
        loss = self.loss_and_grads(xs, ws, eps=eps, n_global=n_global, n_offset=n_offset)
        if self.optim is not None:
            self.optim.step()
        return float(loss.item())
