"""torch.distributions with pyro's to_event / __call__ / has_rsample surface, plus Delta."""
import torch
from torch.distributions import constraints, transforms  # noqa: F401
from torch.distributions import biject_to, transform_to  # noqa: F401

import pyro


class _Mixin:
    def __call__(self, sample_shape=torch.Size()):
        return self.rsample(sample_shape) if self.has_rsample else self.sample(sample_shape)

    def to_event(self, reinterpreted_batch_ndims=None):
        if reinterpreted_batch_ndims is None:
            reinterpreted_batch_ndims = len(self.batch_shape)
        if reinterpreted_batch_ndims == 0:
            return self
        return Independent(self, reinterpreted_batch_ndims)


class Independent(torch.distributions.Independent, _Mixin):
    pass


class Normal(torch.distributions.Normal, _Mixin):
    def rsample(self, sample_shape=torch.Size()):
        shape = self._extended_shape(sample_shape)
        eps = torch.randn(shape, dtype=self.loc.dtype, device=self.loc.device)
        pyro.EPS_LOG.append(eps)
        return self.loc + eps * self.scale


class Dirichlet(torch.distributions.Dirichlet, _Mixin):
    pass


class Categorical(torch.distributions.Categorical, _Mixin):
    pass


class Multinomial(torch.distributions.Multinomial, _Mixin):
    pass


class Delta(torch.distributions.Distribution, _Mixin):
    """Point mass at ``v``: log_prob(x) = log(x == v) summed over the event dims + log_density."""
    has_rsample = True
    arg_constraints = {}

    def __init__(self, v, log_density=0.0, event_dim=0, validate_args=None):
        self.v, self.log_density = v, log_density
        batch_dim = v.dim() - event_dim
        super().__init__(v.shape[:batch_dim], v.shape[batch_dim:], validate_args=False)

    def rsample(self, sample_shape=torch.Size()):
        return self.v.expand(torch.Size(sample_shape) + self.v.shape)

    sample = rsample

    def log_prob(self, x):
        lp = (x == self.v).to(x.dtype).log()
        n = len(self.event_shape)
        if n:
            lp = lp.reshape(lp.shape[:lp.dim() - n] + (-1,)).sum(-1)
        return lp + self.log_density


class util:
    @staticmethod
    def eye_like(value, m, n=None):
        n = m if n is None else n
        eye = torch.zeros(m, n, dtype=value.dtype, device=value.device)
        eye.view(-1)[: min(m, n) * n: n + 1] = 1
        return eye
