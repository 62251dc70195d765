"""Isotropic stationary kernels with the surface of ``pyro.contrib.gp.kernels`` that the reference uses
(``gdrf/train_script.py:93-99`` KERNEL_DICT: ``RBF``, ``Matern32``, ``Matern52``, ``Exponential``): constructor
``(input_dim, variance=None, lengthscale=None, active_dims=None)``, positive-constrained learnable
``variance`` / ``lengthscale`` (stored as ``<name>_unconstrained`` like ``PyroParam``), and
``__call__(X, Z=None, diag=False)``.

The objects only carry the hyper-parameters into the fused ELBO op; ``__call__`` evaluates small
matrices with torch (used for the constructor's ``u_scale_tril`` initialisation,
``sparse_gdrf.py:100-110``) and is not on the hot path.
"""
from __future__ import annotations

import torch
from torch import nn

from ._lib import KERNEL_IDS


class Isotropy(nn.Module):
    kind = ""

    def __init__(self, input_dim: int, variance=None, lengthscale=None, active_dims=None):
        super().__init__()
        self.input_dim = int(input_dim)
        if active_dims is not None and list(active_dims) != list(range(self.input_dim)):
            raise NotImplementedError("active_dims other than all dimensions is not accelerated")
        variance = torch.tensor(1.0) if variance is None else torch.as_tensor(variance, dtype=torch.float32)
        lengthscale = torch.tensor(1.0) if lengthscale is None else torch.as_tensor(lengthscale, dtype=torch.float32)
        self.variance_unconstrained = nn.Parameter(variance.detach().clone().float().log())
        self.lengthscale_unconstrained = nn.Parameter(lengthscale.detach().clone().float().log())

    @property
    def variance(self) -> torch.Tensor:       # constraints.positive -> ExpTransform
        return self.variance_unconstrained.exp()

    @property
    def lengthscale(self) -> torch.Tensor:
        return self.lengthscale_unconstrained.exp()

    @property
    def kernel_id(self) -> int:
        return KERNEL_IDS[self.kind]

    def _square_scaled_dist(self, X, Z):
        ls = self.lengthscale
        sX, sZ = X / ls, Z / ls
        X2 = (sX ** 2).sum(1, keepdim=True)
        Z2 = (sZ ** 2).sum(1, keepdim=True)
        return (X2 - 2 * sX.matmul(sZ.t()) + Z2.t()).clamp(min=0)

    def forward(self, X, Z=None, diag=False):
        if diag:
            return self.variance.expand(X.size(0))
        if X.dim() == 1:
            X = X.unsqueeze(1)
        Z = X if Z is None else (Z.unsqueeze(1) if Z.dim() == 1 else Z)
        r2 = self._square_scaled_dist(X, Z)
        if self.kind == "rbf":
            return self.variance * torch.exp(-0.5 * r2)
        r = (r2 + 1e-12).sqrt()
        if self.kind == "exponential":
            return self.variance * torch.exp(-r)
        if self.kind == "matern32":
            s = (3 ** 0.5) * r
            return self.variance * (1 + s) * torch.exp(-s)
        if self.kind == "matern52":
            s = (5 ** 0.5) * r
            return self.variance * (1 + s + (5.0 / 3.0) * r2) * torch.exp(-s)
        return self.variance * (1 + (0.5 / self.scale_mixture) * r2).pow(-self.scale_mixture)


class RBF(Isotropy):
    kind = "rbf"


class Matern32(Isotropy):
    kind = "matern32"


class Matern52(Isotropy):
    kind = "matern52"


class Exponential(Isotropy):
    kind = "exponential"


class RationalQuadratic(Isotropy):
    """variance * (1 + r2 / (2 scale_mixture))^(-scale_mixture); ``scale_mixture`` is a third positive parameter."""
    kind = "rationalquadratic"

    def __init__(self, input_dim: int, variance=None, lengthscale=None, scale_mixture=None, active_dims=None):
        super().__init__(input_dim, variance, lengthscale, active_dims)
        sm = torch.tensor(1.0) if scale_mixture is None else torch.as_tensor(scale_mixture, dtype=torch.float32)
        self.scale_mixture_unconstrained = nn.Parameter(sm.detach().clone().float().log())

    @property
    def scale_mixture(self) -> torch.Tensor:
        return self.scale_mixture_unconstrained.exp()


# train_script.py:93-99
KERNEL_DICT = {"rbf": RBF, "matern32": Matern32, "matern52": Matern52, "exponential": Exponential,
               "rationalquadratic": RationalQuadratic}


def kernel_kind(kernel) -> str:
    """Maps a kernel object (ours, or a pyro.contrib.gp one) to the fused op's kernel id."""
    name = type(kernel).__name__.lower()
    if name in KERNEL_IDS:
        return name
    raise NotImplementedError(f"kernel {type(kernel).__name__} is not accelerated "
                              "(rbf, matern32, matern52, exponential, rationalquadratic are)")
