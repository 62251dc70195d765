"""Trace_ELBO for guides whose sites are all reparameterised (Normal.rsample, Delta): the surrogate loss equals
-ELBO, ELBO = sum_model scale * log p(site) - sum_guide scale * log q(site).  SVI = loss_and_grads + optimiser."""
import torch

import pyro
from pyro import poutine


class Trace_ELBO:
    def __init__(self, num_particles=1, vectorize_particles=False, max_plate_nesting=None, **kwargs):
        self.num_particles = num_particles

    @staticmethod
    def _site_log_prob(node):
        return (node["fn"].log_prob(node["value"]) * node["scale"]).sum()

    def _one(self, model, guide, args, kwargs):
        guide_trace = poutine.trace(guide).get_trace(*args, **kwargs)
        model_trace = poutine.trace(poutine.replay(model, trace=guide_trace)).get_trace(*args, **kwargs)
        for name, node in guide_trace.nodes.items():
            if not getattr(node["fn"], "has_rsample", False):
                raise NotImplementedError(f"shim: guide site {name} is not reparameterised")
        elbo = sum(self._site_log_prob(n) for n in model_trace.nodes.values())
        elbo = elbo - sum(self._site_log_prob(n) for n in guide_trace.nodes.values())
        self.last_traces = (model_trace, guide_trace)
        return elbo

    def differentiable_loss(self, model, guide, *args, **kwargs):
        elbo = sum(self._one(model, guide, args, kwargs) for _ in range(self.num_particles)) / self.num_particles
        return -elbo

    def loss(self, model, guide, *args, **kwargs):
        with torch.no_grad():
            return float(self.differentiable_loss(model, guide, *args, **kwargs))

    def loss_and_grads(self, model, guide, *args, **kwargs):
        loss = self.differentiable_loss(model, guide, *args, **kwargs)
        loss.backward()
        return float(loss)


TraceGraph_ELBO = Trace_ELBO          # identical for fully reparameterised guides


class SVI:
    def __init__(self, model, guide, optim, loss, **kwargs):
        self.model, self.guide, self.optim, self.loss = model, guide, optim, loss

    def step(self, *args, **kwargs):
        loss = self.loss.loss_and_grads(self.model, self.guide, *args, **kwargs)
        params = [u for u, _ in pyro.get_param_store().values() if u.grad is not None]
        self.optim(params)
        for p in params:
            p.grad = None
        return loss
