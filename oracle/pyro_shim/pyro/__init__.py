"""Minimal stand-in for pyro-ppl, written for ONE purpose: to execute the reference's own, unmodified
``gdrf/models/*.py`` in a container where pyro-ppl (pinned 1.8.0 by the reference's poetry.lock) is not installed.

TEST INFRASTRUCTURE ONLY (lives under oracle/; used by oracle/make_ref_fixtures.py, never by the product).

It restates, from the published behaviour of pyro-ppl 1.8, only the primitives the reference's SVI path touches
(SURVEY.md section 8c): ``pyro.sample / plate / param``, the effect handlers ``trace / replay / scale``,
``Trace_ELBO`` for fully reparameterised guides, ``PyroModule / PyroParam`` constraint handling (through torch's
own ``transform_to`` registry), ``pyro.contrib.gp`` ``Parameterized``, the isotropic kernels and
``gp.util.conditional``, and thin wrappers over ``torch.distributions``.  It is NOT pyro: whatever the reference
computes through these primitives is pinned only as far as this restatement is faithful -- DESIGN.md section 2 says so.
"""
from collections import OrderedDict

import torch

_HANDLERS = []            # innermost last
_PARAM_STORE = OrderedDict()
EPS_LOG = []              # standard-normal draws consumed by Normal.rsample, in order (fixture generation reads it)


class _Handler:
    def __enter__(self):
        _HANDLERS.append(self)
        return self

    def __exit__(self, *exc):
        assert _HANDLERS.pop() is self
        return False

    def process(self, msg):      # innermost first, before the value is drawn
        pass

    def postprocess(self, msg):  # after the value is known
        pass


def sample(name, fn, obs=None, **kwargs):
    msg = {"type": "sample", "name": name, "fn": fn, "value": obs, "is_observed": obs is not None, "scale": 1.0}
    for h in reversed(_HANDLERS):
        h.process(msg)
    if msg["value"] is None:
        msg["value"] = fn.rsample() if getattr(fn, "has_rsample", False) else fn.sample()
    for h in reversed(_HANDLERS):
        h.postprocess(msg)
    return msg["value"]


def deterministic(name, value, event_dim=None):
    return value


class plate:
    """Without subsampling a plate only declares independence: scale 1, indices = arange(size)."""

    def __init__(self, name, size=None, subsample_size=None, subsample=None, dim=None, use_cuda=None, device=None):
        if subsample_size is not None or subsample is not None:
            raise NotImplementedError("shim: plates without subsampling only")
        self.name, self.size, self.device = name, size, device

    def __enter__(self):
        return torch.arange(self.size, device=self.device)

    def __exit__(self, *exc):
        return False


def param(name, *args, **kwargs):
    if name not in _PARAM_STORE:
        raise KeyError(name)
    unconstrained, constraint = _PARAM_STORE[name]
    return torch.distributions.transform_to(constraint)(unconstrained)


class _Unit:
    """pyro.distributions.Unit(log_factor): an empty-valued site whose log-density is the given factor."""
    has_rsample = True

    def __init__(self, log_factor):
        self.log_factor = log_factor

    def log_prob(self, value):
        return self.log_factor


def factor(name, log_factor):
    """pyro.factor: adds ``log_factor`` to the model's log-density (observed Unit site)."""
    return sample(name, _Unit(log_factor), obs=torch.empty(0))


def module(name, nn_module, update_module_params=False):
    """pyro.module: registers every parameter of a torch module with the param store (unconstrained = the parameter
    itself, constraint real), so that SVI's optimiser steps it."""
    for pname, p in nn_module.named_parameters():
        _PARAM_STORE.setdefault(f"{name}$$${pname}", (p, torch.distributions.constraints.real))
    return nn_module


def clear_param_store():
    _PARAM_STORE.clear()


def get_param_store():
    return _PARAM_STORE


from . import distributions, nn, poutine, infer, ops, optim  # noqa: E402,F401
from . import contrib  # noqa: E402,F401
