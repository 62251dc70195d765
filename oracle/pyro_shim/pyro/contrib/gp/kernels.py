"""Isotropic stationary kernels with pyro.contrib.gp's arithmetic: squared distance through the
|x|^2 - 2 x.z + |z|^2 expansion clamped at 0, r = sqrt(r2 + 1e-12), positive-constrained variance / lengthscale."""
import numbers

import torch
from torch.distributions import constraints

from pyro.nn import PyroParam
from .parameterized import Parameterized


def _torch_sqrt(x, eps=1e-12):
    return (x + eps).sqrt()


class Kernel(Parameterized):
    def __init__(self, input_dim, active_dims=None):
        super().__init__()
        if active_dims is None:
            active_dims = list(range(input_dim))
        elif input_dim != len(active_dims):
            raise ValueError("Input size and the length of active dimensionals should be equal.")
        self.input_dim = input_dim
        self.active_dims = active_dims

    def forward(self, X, Z=None, diag=False):
        raise NotImplementedError

    def _slice_input(self, X):
        if X.dim() == 2:
            return X[:, self.active_dims]
        if X.dim() == 1:
            return X
        raise ValueError("Input X must be either 1 or 2 dimensional.")


class Isotropy(Kernel):
    def __init__(self, input_dim, variance=None, lengthscale=None, active_dims=None):
        super().__init__(input_dim, active_dims)
        variance = torch.tensor(1.0) if variance is None else variance
        self.variance = PyroParam(variance, constraints.positive)
        lengthscale = torch.tensor(1.0) if lengthscale is None else lengthscale
        self.lengthscale = PyroParam(lengthscale, constraints.positive)

    def _square_scaled_dist(self, X, Z=None):
        if Z is None:
            Z = X
        X = self._slice_input(X)
        Z = self._slice_input(Z)
        if X.size(1) != Z.size(1):
            raise ValueError("Inputs must have the same number of features.")
        scaled_X = X / self.lengthscale
        scaled_Z = Z / self.lengthscale
        X2 = (scaled_X ** 2).sum(1, keepdim=True)
        Z2 = (scaled_Z ** 2).sum(1, keepdim=True)
        XZ = scaled_X.matmul(scaled_Z.t())
        r2 = X2 - 2 * XZ + Z2.t()
        return r2.clamp(min=0)

    def _scaled_dist(self, X, Z=None):
        return _torch_sqrt(self._square_scaled_dist(X, Z))

    def _diag(self, X):
        return self.variance.expand(X.size(0))


class RBF(Isotropy):
    def forward(self, X, Z=None, diag=False):
        if diag:
            return self._diag(X)
        r2 = self._square_scaled_dist(X, Z)
        return self.variance * torch.exp(-0.5 * r2)


class Exponential(Isotropy):
    def forward(self, X, Z=None, diag=False):
        if diag:
            return self._diag(X)
        r = self._scaled_dist(X, Z)
        return self.variance * torch.exp(-r)


class Matern32(Isotropy):
    def forward(self, X, Z=None, diag=False):
        if diag:
            return self._diag(X)
        r = self._scaled_dist(X, Z)
        sqrt3_r = 3 ** 0.5 * r
        return self.variance * (1 + sqrt3_r) * torch.exp(-sqrt3_r)


class Matern52(Isotropy):
    def forward(self, X, Z=None, diag=False):
        if diag:
            return self._diag(X)
        r2 = self._square_scaled_dist(X, Z)
        r = _torch_sqrt(r2)
        sqrt5_r = 5 ** 0.5 * r
        return self.variance * (1 + sqrt5_r + (5 / 3) * r2) * torch.exp(-sqrt5_r)


class RationalQuadratic(Isotropy):
    def __init__(self, input_dim, variance=None, lengthscale=None, scale_mixture=None, active_dims=None):
        super().__init__(input_dim, variance, lengthscale, active_dims)
        if scale_mixture is None:
            scale_mixture = torch.tensor(1.0)
        self.scale_mixture = PyroParam(scale_mixture, constraints.positive)

    def forward(self, X, Z=None, diag=False):
        if diag:
            return self._diag(X)
        r2 = self._square_scaled_dist(X, Z)
        return self.variance * (1 + (0.5 / self.scale_mixture) * r2).pow(-self.scale_mixture)
