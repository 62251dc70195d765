"""PyroModule / PyroParam: a constrained parameter ``x`` is stored as ``x_unconstrained`` (an nn.Parameter holding
``transform_to(constraint).inv(init)``) and read back as ``transform_to(constraint)(x_unconstrained)`` on every
attribute access; reads also register the unconstrained leaf in the global param store under its full name."""
from collections import OrderedDict, namedtuple

import torch
from torch.distributions import constraints, transform_to

import pyro


class PyroParam(namedtuple("PyroParam", ("init_value", "constraint", "event_dim"))):
    def __new__(cls, init_value=None, constraint=constraints.real, event_dim=None):
        return super().__new__(cls, init_value, constraint, event_dim)


class PyroSample(namedtuple("PyroSample", ("prior",))):
    pass


def pyro_method(fn):
    return fn


class PyroModule(torch.nn.Module):
    def __init__(self, name=""):
        self._pyro_name = name
        self._pyro_params = OrderedDict()
        self._pyro_samples = OrderedDict()
        super().__init__()

    def _pyro_get_fullname(self, name):
        return f"{self._pyro_name}.{name}" if self._pyro_name else name

    def _pyro_set_supermodule(self, name):
        self._pyro_name = name
        for key, value in self._modules.items():
            if isinstance(value, PyroModule):
                value._pyro_set_supermodule(f"{name}.{key}" if name else key)

    def __setattr__(self, name, value):
        if isinstance(value, PyroModule):
            value._pyro_set_supermodule(self._pyro_get_fullname(name))
            return super().__setattr__(name, value)
        if isinstance(value, PyroParam):
            init, constraint, event_dim = value
            self._pyro_params[name] = (constraint, event_dim)
            with torch.no_grad():
                unconstrained = transform_to(constraint).inv(init.detach()).contiguous().clone()
            return super().__setattr__(name + "_unconstrained", torch.nn.Parameter(unconstrained))
        if isinstance(value, PyroSample):
            raise NotImplementedError("shim: PyroSample attributes are not on the accelerated path")
        params = self.__dict__.get("_pyro_params")
        if params is not None and name in params and isinstance(value, torch.Tensor):
            constraint, _ = params[name]
            with torch.no_grad():
                getattr(self, name + "_unconstrained").data = transform_to(constraint).inv(value.detach()).contiguous()
            return None
        return super().__setattr__(name, value)

    def __getattr__(self, name):
        params = self.__dict__.get("_pyro_params")
        if params is not None and name in params:
            constraint, _ = params[name]
            unconstrained = super().__getattr__(name + "_unconstrained")
            pyro._PARAM_STORE[self._pyro_get_fullname(name)] = (unconstrained, constraint)
            return transform_to(constraint)(unconstrained)
        return super().__getattr__(name)
