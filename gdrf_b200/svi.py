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


def global_count(n_local: int, device, group=None) -> int:
    """The N of the reference's ``poutine.scale(1/N)`` (train_script.py:365) when every rank holds a shard: the sum of
    the shard sizes.  A caller that does not pass ``n_global`` gets this instead of silently normalising by its own
    shard (which would make the all-reduced loss and gradients world_size times too large)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return int(n_local)
    t = torch.tensor([int(n_local)], dtype=torch.int64, device=device if dist.get_backend(group) == "nccl" else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t.item())


def shard_bounds(n: int, rank: int, world: int):
    """Contiguous shard [lo, hi) of n observations for `rank`."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class ClippedAdam(torch.optim.Optimizer):
    """``pyro.optim.ClippedAdam`` (pyro-ppl 1.8.0 ``pyro/optim/clipped_adam.py``; the optimiser of
    ``scripts/mvco.py:135`` and OPTIMIZER_DICT["clippedadam"], ``train_script.py:75``) for the un-fused path: the
    gradient is clamped element-wise to ``[-clip_norm, clip_norm]``, weight decay is L2, the learning rate decays by
    ``lrd`` every step, ``denom = sqrt(v) + eps`` and ``step = lr sqrt(1 - beta2^t) / (1 - beta1^t)``."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, clip_norm=10.0, lrd=1.0):
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, clip_norm=clip_norm,
                                      lrd=lrd))

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        for group in self.param_groups:
            group["lr"] *= group["lrd"]
            beta1, beta2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                grad = p.grad.clamp_(-group["clip_norm"], group["clip_norm"])
                state = self.state[p]
                if len(state) == 0:
                    state["step"] = 0
                    state["exp_avg"] = torch.zeros_like(grad)
                    state["exp_avg_sq"] = torch.zeros_like(grad)
                state["step"] += 1
                if group["weight_decay"] != 0:
                    grad = grad.add(p, alpha=group["weight_decay"])
                state["exp_avg"].mul_(beta1).add_(grad, alpha=1 - beta1)
                state["exp_avg_sq"].mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
                denom = state["exp_avg_sq"].sqrt().add_(group["eps"])
                bc1 = 1 - beta1 ** state["step"]
                bc2 = 1 - beta2 ** state["step"]
                p.addcdiv_(state["exp_avg"], denom, value=-group["lr"] * (bc2 ** 0.5) / bc1)
        return loss


class SVI:
    def __init__(self, model, guide=None, optim: Optional[torch.optim.Optimizer] = None, loss=None, group=None):
        self.module = getattr(model, "__self__", model)
        self.optim = optim
        self.group = group

    def loss_and_grads(self, xs, ws, eps=None, n_global=None, n_offset=0) -> torch.Tensor:
        rank = dist.get_rank(self.group) if dist.is_available() and dist.is_initialized() else 0
        if n_global is None:
            n_global = global_count(xs.shape[0], xs.device, self.group)
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


class FusedSVI:
    """SVI step with everything after the data on the device and fused: the ELBO + gradient op on the flat
    constrained buffer, then ONE kernel for the chain rule through the PyroParam constraints and the Adam / AdamW
    update of the flat unconstrained buffer (the reference does this with a per-parameter
    Python loop, ``train_script.py:325-327``).  Parameters are taken from / written back to a
    :class:`gdrf_b200.models.SparseMultinomialGDRF`.
    """

    def __init__(self, module, lr: float = 1e-3, betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0,
                 group=None, clip_norm: float = 0.0, lrd: float = 1.0):
        """``clip_norm > 0`` selects pyro's ClippedAdam (with its ``lrd`` decay and L2 ``weight_decay``) instead of
        Adam / AdamW."""
        self.clip_norm, self.lrd = float(clip_norm), float(lrd)
        import ctypes
        from . import _lib
        self._ct, self._lib = ctypes, _lib
        self.m = module
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), betas, float(eps), float(weight_decay)
        self.group = group
        self.t = 0
        dev = module.device
        K, M, V, D = module.K, module.M, module.V, module.D
        self.ls_dim = int(module._kernel.lengthscale.numel())
        self.learn_z = 0 if module._fixed_inducing_points else 1
        with torch.no_grad():
            z_u = module._inducing_points_fixed if not self.learn_z else module._inducing_points_unconstrained
            parts = [module.u_scale_tril_unconstrained.reshape(-1), module.u_loc_unconstrained.reshape(-1),
                     module._word_topic_matrix_map_unconstrained.reshape(-1), z_u.reshape(-1),
                     module._kernel.variance_unconstrained.reshape(-1), module._kernel.lengthscale_unconstrained.reshape(-1),
                     module.noise_unconstrained.reshape(-1)]
            if module._kernel_kind == "rationalquadratic":
                parts.append(module._kernel.scale_mixture_unconstrained.reshape(-1))
            self.theta_u = torch.cat([p.detach().float() for p in parts]).contiguous().to(dev)
            # the module's parameters become views of the flat buffer: what FusedSVI trains is what the module
            # evaluates (perplexity, kernel_lengthscale, checkpoints -- train_script.py:468-500) with nothing to sync
            o = 0
            for p_, n_ in self._targets():
                if p_ is not None:
                    p_.data = self.theta_u[o:o + n_].view_as(p_)
                o += n_
        self.theta_c = torch.empty_like(self.theta_u)
        self.mom1 = torch.zeros_like(self.theta_u)
        self.mom2 = torch.zeros_like(self.theta_u)
        self.row_scratch = torch.empty(K, dtype=torch.float32, device=dev)
        self.shape = _lib.Shape(n_local=1, n_offset=0, n_eps=1, d=D, m=M, k=K, v=V,
                                kernel_id=_lib.KERNEL_IDS[module._kernel_kind], ls_dim=self.ls_dim, chunk_rows=0, flags=0)

    def _targets(self):
        m = self.m
        K, M, V, D = m.K, m.M, m.V, m.D
        t = [(m.u_scale_tril_unconstrained, K * M * M), (m.u_loc_unconstrained, K * M),
             (m._word_topic_matrix_map_unconstrained, K * V),
             (m._inducing_points_unconstrained if self.learn_z else None, M * D),
             (m._kernel.variance_unconstrained, 1), (m._kernel.lengthscale_unconstrained, self.ls_dim),
             (m.noise_unconstrained, 1)]
        if m._kernel_kind == "rationalquadratic":
            t.append((m._kernel.scale_mixture_unconstrained, 1))
        return t

    def _views(self):
        from .elbo import split_grad
        K, M, V, D = self.m.K, self.m.M, self.m.V, self.m.D
        return split_grad(self.theta_c, K, M, V, D, self.ls_dim)

    def constrain(self):
        st = torch.cuda.current_stream(self.theta_u.device).cuda_stream
        self._lib.check(self._lib.load().gdrf_constrain(self._ct.byref(self.shape), self.theta_u.data_ptr(),
                                                        self.theta_c.data_ptr(), self.learn_z, st))
        return self._views()

    def step(self, xs, ws, subsample=False, eps=None, n_global=None, n_offset=0) -> float:
        from .elbo import elbo_value_and_grads
        m = self.m
        if getattr(m, "_reference_double_scale", False) and not m._unit_world:
            raise NotImplementedError("reference_double_scale on a non-unit world runs through SparseMultinomialGDRF.elbo "
                                      "(gdrf_b200.svi.SVI / torch optimisers), not the fused step")
        c = self.constrain()
        x = m._scaled(xs)
        N = x.shape[0]
        n_global = global_count(N, m.device, self.group) if n_global is None else int(n_global)
        if eps is None:     # drawn for exactly these observations; num_particles draws each (scripts/mvco.py:136)
            P = int(getattr(m, "num_particles", 1))
            eps = torch.randn(*((P,) if P > 1 else ()), m.K, N, device=m.device, generator=m._eps_generator)
            n_offset = 0
        rank = dist.get_rank(self.group) if dist.is_available() and dist.is_initialized() else 0
        terms, g, _ = elbo_value_and_grads(x, ws.to(m.device), c["Z"], c["variance"], c["lengthscale"], c["u_loc"],
                                           c["u_scale_tril"], c["noise"], c["phi"], m._dirichlet_param, eps,
                                           kernel=m._kernel_kind, jitter=m._jitter, maxjitter=m._maxjitter,
                                           n_global=n_global, n_offset=n_offset, include_prior=(rank == 0),
                                           scale_mixture=c.get("scale_mixture"), all_reduce=True, group=self.group)
        from .elbo import flat_gradient
        flat = flat_gradient(g)[:self.theta_u.numel()]      # summed over the ranks, like terms
        self.t += 1
        st = torch.cuda.current_stream(self.theta_u.device).cuda_stream
        if self.clip_norm > 0.0:
            self.lr *= self.lrd       # ClippedAdam decays before the update
            self._lib.check(self._lib.load().gdrf_clipped_adam_step(
                self._ct.byref(self.shape), self.theta_u.data_ptr(), self.theta_c.data_ptr(), flat.data_ptr(),
                self.mom1.data_ptr(), self.mom2.data_ptr(), self.row_scratch.data_ptr(), self.lr, self.betas[0],
                self.betas[1], self.eps, self.weight_decay, self.clip_norm, self.t, -1.0 / n_global, self.learn_z, st))
        else:
            self._lib.check(self._lib.load().gdrf_adam_step(
                self._ct.byref(self.shape), self.theta_u.data_ptr(), self.theta_c.data_ptr(), flat.data_ptr(),
                self.mom1.data_ptr(), self.mom2.data_ptr(), self.row_scratch.data_ptr(), self.lr, self.betas[0],
                self.betas[1], self.eps, self.weight_decay, self.t, -1.0 / n_global, self.learn_z, st))
        elbo = terms[0] + terms[3] + terms[2] - terms[1]
        return float((-elbo / n_global).item())

    def write_back(self) -> None:
        """Kept for callers of the first version: the module's parameters are views of the flat buffer, so there is
        nothing to copy (a parameter that was re-assigned since is re-attached)."""
        o = 0
        with torch.no_grad():
            for p, n in self._targets():
                if p is not None and p.data_ptr() != self.theta_u[o:o + n].data_ptr():
                    p.data = self.theta_u[o:o + n].view_as(p)
                o += n
