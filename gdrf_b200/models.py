"""Drop-in for ``gdrf.models.SparseMultinomialGDRF`` (reference ``gdrf/models/sparse_gdrf.py:16-129,
322-409`` and its bases ``abstract_gdrf.py`` / ``topic_model.py``) with the ELBO evaluated by the fused
sm_100a op in :mod:`gdrf_b200.elbo`.

Same constructor keywords, parameter names (``u_loc``, ``u_scale_tril``, ``noise``, ``_inducing_points``,
``_word_topic_matrix_map``, ``_kernel.variance``, ``_kernel.lengthscale`` -- each stored
``<name>_unconstrained`` the way ``pyro.nn.PyroParam`` stores them), methods (``model``, ``guide``,
``log_topic_probs``, ``topic_probs``, ``word_probs``, ``perplexity``, ``ml_topics``, ``ml_words``,
``scale``, ``artifacts``) and error behaviour.  Only ``whiten=True`` with the default zero mean and
softmax link is accelerated; anything else raises ``NotImplementedError`` (there is no fallback path).

When Pyro is importable, ``model`` registers the module's parameters (``pyro.module``) and contributes the whole
ELBO as one ``pyro.factor``; ``guide`` has nothing to sample, so ``pyro.infer.SVI(model=scale(m.model),
guide=scale(m.guide), ...)`` (``train_script.py:365-371``) sees the identical loss and steps the same parameters
(``tests/test_gpu_streaming.py::test_pyro_adapter_path_under_the_shim`` drives exactly that wiring).  Without Pyro,
:class:`gdrf_b200.svi.SVI` provides the same ``step(xs=, ws=, subsample=)`` call.
"""
from __future__ import annotations

import warnings
from typing import Callable, List, Optional, Tuple, Union

import torch
from torch import nn

from . import _lib
from .elbo import GDRFElbo, MarginalMoments, marginal_mean, marginal_moments, perplexity_from_mean
from .kernels import kernel_kind

try:  # optional: the reference's inference driver
    import pyro  # type: ignore
    _HAVE_PYRO = True
except Exception:  # pragma: no cover - pyro is absent in the build image
    pyro = None
    _HAVE_PYRO = False


def validate_dirichlet_param(b: torch.Tensor, K: int, V: int, device="cuda") -> torch.Tensor:
    """gdrf/models/utils.py:6-24."""
    assert (b <= 0).sum().item() == 0, "b must be positive"
    if b.dim() == 0:
        return torch.ones(K, V, device=device) * b.to(device)
    if b.dim() == 1:
        if b.shape[0] == K:
            return b.repeat(V, 1).T.to(device)
        if b.shape[0] == V:
            return b.repeat(K, 1).to(device)
        raise ValueError("parameter b must have length K or V if 1D")
    if b.dim() == 2:
        assert b.shape == torch.Size((K, V)), "b should be KxV if 2D"
        return b.to(device)
    raise ValueError("invalid b parameter- you passed %s" % (b,))


def host_jittercholesky(Kff: torch.Tensor, N: int, jitter: float, maxjitter: int) -> torch.Tensor:
    """gdrf/models/utils.py:27-40, used once by the constructor to initialise u_scale_tril."""
    Kff = Kff.clone()
    njitter = 0
    while njitter < maxjitter:
        try:
            Kff.view(-1)[:: N + 1] += jitter * (10 ** njitter)
            return torch.linalg.cholesky(Kff)
        except RuntimeError:
            njitter += 1
    raise RuntimeError("reached max jitter, covariance is unstable")


class SparseMultinomialGDRF(nn.Module):
    def __init__(self, num_observation_categories: int, num_topic_categories: int,
                 world: List[Tuple[float, float]], kernel, dirichlet_param: Union[float, torch.Tensor],
                 n_points: Union[int, List[int]], fixed_inducing_points: bool = False,
                 inducing_init: str = "random", mean_function: Callable = None, link_function: Callable = None,
                 noise: Optional[float] = None, device: str = "cpu", whiten: bool = True, jitter: float = 1e-8,
                 maxjitter: int = 5, randomize_wt_matrix: bool = False, randomize_metric=None,
                 randomize_iters: int = 100, reference_double_scale: bool = False, **kwargs):
        super().__init__()
        if mean_function is not None or link_function is not None:
            raise NotImplementedError("only the default zero mean and softmax link are accelerated")
        if not whiten:
            raise NotImplementedError("only whiten=True is accelerated")
        self._V = int(num_observation_categories)
        self._K = int(num_topic_categories)
        self.device = torch.device(device)
        self._world = world
        self._lower_bounds = torch.tensor([b[0] for b in world], dtype=torch.float32, device=self.device)
        self._upper_bounds = torch.tensor([b[1] for b in world], dtype=torch.float32, device=self.device)
        self._delta_bounds = self._upper_bounds - self._lower_bounds
        self._n_dims = len(world)
        self._unit_world = all(float(b[0]) == 0.0 and float(b[1]) == 1.0 for b in world)
        self._warned_world = False
        # opt-in (not a reference keyword): reproduce the reference guide's second scale() for a non-unit world
        # (sparse_gdrf.py:380) -- see elbo()
        self._reference_double_scale = bool(reference_double_scale)
        self._kernel = kernel.to(self.device)
        self._kernel_kind = kernel_kind(kernel)
        if isinstance(dirichlet_param, float):
            dirichlet_param = torch.tensor(dirichlet_param)
        self._dirichlet_param = validate_dirichlet_param(torch.as_tensor(dirichlet_param, dtype=torch.float32),
                                                         self._K, self._V, device=self.device)
        self._n_points = [n_points for _ in world] if isinstance(n_points, int) else list(n_points)
        self._fixed_inducing_points = fixed_inducing_points
        if inducing_init == "random":
            points = [torch.sort(torch.rand(self._n_points[i]))[0].to(self.device) * self._delta_bounds[i]
                      + self._lower_bounds[i] for i in range(self.dims)]
        elif inducing_init == "grid":
            points = [torch.arange(b[0], b[1] + (b[1] - b[0]) / (n - 1) - 1e-10, (b[1] - b[0]) / (n - 1))
                      for b, n in zip(world, self._n_points)]
        else:
            raise ValueError(f"inducing_init argument {inducing_init} not valid. Only 'random' and 'grid' "
                             "are currently supported")
        inducing = torch.stack([x.flatten() for x in torch.meshgrid(*points, indexing="ij")]).T.to(self.device)
        scaled = self.scale(inducing.float())
        if fixed_inducing_points:
            self.register_buffer("_inducing_points_fixed", scaled)
        else:   # interval(0, 1) per dimension -> sigmoid
            fi = torch.finfo(scaled.dtype)      # the clamp of torch's SigmoidTransform inverse (grid end points are 0 and 1)
            sc = scaled.clamp(min=fi.tiny, max=1.0 - fi.eps)
            self._inducing_points_unconstrained = nn.Parameter(torch.log(sc) - torch.log1p(-sc))
        self._jitter = float(jitter)
        self._maxjitter = int(maxjitter)
        self._whiten = whiten
        self.M = scaled.size(-2)
        self.D = scaled.size(-1)
        self.u_loc_unconstrained = nn.Parameter(torch.zeros(self._K, self.M, device=self.device))
        with torch.no_grad():
            # the reference factorises k(self._inducing_points): the constrained round trip, not the raw grid (:101-106)
            L = host_jittercholesky(self._kernel(self._inducing_points.detach()).contiguous(), self.M, self._jitter,
                                    self._maxjitter)
        S0 = L.float().repeat(self._K, 1, 1)
        # lower_cholesky: strictly-lower entries free, diagonal through exp
        unc = S0.tril(-1) + S0.diagonal(dim1=-2, dim2=-1).log().diag_embed()
        self.u_scale_tril_unconstrained = nn.Parameter(unc)
        noise = torch.tensor(1.0) if noise is None else torch.as_tensor(noise, dtype=torch.float32)
        self.noise_unconstrained = nn.Parameter(noise.to(self.device).log())
        # make_wt_matrix (abstract_gdrf.py:57-84): softmax over the *topic* axis of beta, then stored through
        # the stacked-simplex constraint (unconstrained = log p, constrained = row softmax)
        wt = torch.softmax(self._dirichlet_param, dim=-2)
        best = -1 if randomize_metric is None else randomize_metric(wt, self)
        if randomize_wt_matrix:     # the reference never updates `best`: the LAST candidate that beats the initial score wins
            for _ in range(1 if randomize_metric is None else randomize_iters):
                possible = torch.softmax(torch.randn_like(wt), dim=-2)
                score = 0 if randomize_metric is None else randomize_metric(possible, self)
                if score > best:
                    wt = possible
        self._word_topic_matrix_map_unconstrained = nn.Parameter(wt.log())
        self._eps_generator: Optional[torch.Generator] = None
        self.num_particles = 1
        if "xs" in kwargs and "ws" in kwargs and self.device.type == "cuda":
            with torch.no_grad():       # the reference runs self.model once at the end of __init__ (:123)
                self.elbo(kwargs["xs"].to(self.device), kwargs["ws"].to(self.device))

    # ------------------------------------------------------------------ constrained views
    @property
    def K(self): return self._K

    @property
    def V(self): return self._V

    @property
    def dims(self): return self._n_dims

    @property
    def u_loc(self): return self.u_loc_unconstrained

    @property
    def u_scale_tril(self):
        u = self.u_scale_tril_unconstrained
        return u.tril(-1) + u.diagonal(dim1=-2, dim2=-1).exp().diag_embed()

    @property
    def noise(self): return self.noise_unconstrained.exp()

    @property
    def _inducing_points(self):
        if self._fixed_inducing_points:
            return self._inducing_points_fixed
        return torch.sigmoid(self._inducing_points_unconstrained)

    @property
    def _word_topic_matrix_map(self):
        return torch.softmax(self._word_topic_matrix_map_unconstrained, dim=-1)

    @property
    def word_topic_matrix(self): return self._word_topic_matrix_map

    @property
    def _scale_mixture(self):
        """RationalQuadratic's third hyper-parameter (None for the other kernels)."""
        return getattr(self._kernel, "scale_mixture", None)

    @property
    def kernel_lengthscale(self): return self._kernel.lengthscale.detach().cpu().numpy()

    @property
    def kernel_variance(self): return self._kernel.variance.detach().cpu().numpy()

    # ------------------------------------------------------------------ input scaling (topic_model.py:168-198)
    def scale(self, input: torch.Tensor) -> torch.Tensor:
        return (input - self._lower_bounds) / self._delta_bounds

    def _check_bounds(self, input: torch.Tensor, epsilon: float = 1e-8) -> bool:
        inp = input.to(self.device)
        return bool(input.shape[-1] == self._n_dims and ((inp - self._lower_bounds > -epsilon)
                                                         & (inp - self._upper_bounds < epsilon)).all())

    def _scaled(self, xs: torch.Tensor) -> torch.Tensor:
        # the reference asserts the bounds on every call (topic_model.py:177: one device->host sync each, three per
        # step); a tensor that was already checked and has not been written since is not checked again
        # (a plain tuple, not a weak reference: the reference checkpoints the whole module with torch.save,
        # train_script.py:490-500, so every attribute has to pickle)
        key = (id(xs), xs.data_ptr(), tuple(xs.shape), xs._version)
        if getattr(self, "_bounds_checked", None) != key:
            assert self._check_bounds(xs)
            self._bounds_checked = key
        return self.scale(xs.to(self.device).float())

    def _check_Xnew_shape(self, Xnew: torch.Tensor):
        Z = self._inducing_points
        if Xnew.dim() != Z.dim():
            raise ValueError("Inducing points and test data should have the same number of dimensions, "
                             "but got {} and {}.".format(Z.dim(), Xnew.dim()))
        if Z.shape[1:] != Xnew.shape[1:]:
            raise ValueError("Inducing points and test data should have the same shape of features, "
                             "but got {} and {}.".format(Z.shape[1:], Xnew.shape[1:]))

    # ------------------------------------------------------------------ the hot path
    def seed_eps(self, seed: int) -> None:
        """Seeds the generator of the guide's self-drawn ``eps``.  Under torch.distributed every rank draws for its own
        observations, so the rank is folded into the seed (identical streams would correlate the shards)."""
        import torch.distributed as dist
        rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        self._eps_generator = torch.Generator(device=self.device).manual_seed(int(seed) + 1000003 * rank)

    def elbo(self, xs: torch.Tensor, ws: torch.Tensor, eps: Optional[torch.Tensor] = None,
             n_global: Optional[int] = None, n_offset: int = 0, include_prior: bool = True,
             flags: int = _lib.FLAG_CHOL_FP32_STATUS, chunk_rows: int = 0, scaled: bool = False) -> torch.Tensor:
        """ELBO / N of model+guide for (xs, ws) -- differentiable w.r.t. every parameter.  ``eps`` are the
        guide's standard-normal draws for the ``mu`` site ([K, N]); drawn on the device when omitted."""
        xs = xs.to(self.device)
        self._check_Xnew_shape(xs)
        x = xs.float() if scaled else self._scaled(xs)
        double_scale = self._reference_double_scale and not scaled and not self._unit_world
        if not scaled and not self._unit_world and not self._warned_world and not double_scale:
            # sparse_gdrf.py:380: the reference's guide applies scale() a second time on top of scale_decorator, so
            # for a world other than [0,1]^D its guide and model condition on different inputs.  train() always
            # hands over unit-cube data (train_script.py:263-271); here xs is scaled once for both unless the module
            # was built with reference_double_scale=True.
            warnings.warn("world is not the unit cube: the reference's guide rescales xs twice "
                          "(sparse_gdrf.py:380); gdrf_b200 scales once for model and guide "
                          "(reference_double_scale=True reproduces the reference)")
            self._warned_world = True
        N = x.shape[0]
        if eps is None:     # drawn here for exactly these observations: the window into eps starts at 0
            eps = torch.randn(self._K, N, device=self.device, generator=self._eps_generator)
            n_offset = 0
        n_global = N if n_global is None else int(n_global)
        ws = ws.to(self.device)

        def one(e, x=x):
            return GDRFElbo.apply(x, ws, self._inducing_points, self._kernel.variance, self._kernel.lengthscale,
                                  self.u_loc, self.u_scale_tril, self.noise, self._word_topic_matrix_map,
                                  self._dirichlet_param, e, _lib.KERNEL_IDS[self._kernel_kind], self._jitter,
                                  self._maxjitter, n_global, n_offset, include_prior, flags, chunk_rows,
                                  self._scale_mixture)

        # [P, K, N] (or num_particles > 1): Trace_ELBO(num_particles=P, vectorize_particles=True) averages the particles'
        # ELBOs (train_script.py:330-335); the op shares the contractions between them
        if eps.dim() == 2 and self.num_particles > 1:
            extra = torch.randn(self.num_particles - 1, self._K, eps.shape[-1], device=self.device,
                                generator=self._eps_generator)
            eps = torch.cat([eps.unsqueeze(0), extra])
        if not double_scale:
            return one(eps)
        # The reference on a non-unit world: the guide draws mu from the marginal at scale(scale(xs)) and scores it
        # there (lq); the model scores the same mu under the marginal at scale(xs) (lp_mu); the likelihood and the prior
        # only see mu and phi.  So ELBO_ref = ELBO(all terms at the guide's inputs) + lp_mu(mu | model's moments)
        # - lp_mu(mu | guide's moments): the fused op evaluates the first term, and the correction is an elementwise
        # function of the two sets of marginal moments (fp64, [K, N]), differentiated through gdrf_moments_vjp.
        x_guide = self.scale(x)
        kid = _lib.KERNEL_IDS[self._kernel_kind]
        base = one(eps, x_guide)
        gp_args = (self._inducing_points, self._kernel.variance, self._kernel.lengthscale, self.u_loc, self.u_scale_tril,
                   kid, self._jitter, self._maxjitter, flags, chunk_rows, self._scale_mixture, True)
        fl_g, fv_g = MarginalMoments.apply(x_guide, *gp_args)
        fl_m, fv_m = MarginalMoments.apply(x, *gp_args)
        e = eps[..., n_offset:n_offset + N].double()
        mu = fl_g + fv_g * e
        noise = self.noise.double()

        def lp(fl, fv):        # Normal(fl, fv + noise).log_prob(mu) without its constant; the variance is the scale
            sp = fv + noise
            return (-torch.log(sp) - 0.5 * ((mu - fl) / sp) ** 2).sum(dim=(-2, -1))

        delta = (lp(fl_m, fv_m) - lp(fl_g, fv_g)).mean()          # mean over the particles
        return base + (delta / float(n_global)).to(base.dtype)

    def model(self, xs, ws, subsample=False):
        """sparse_gdrf.py:322-373.  Under Pyro: one factor carrying N * (ELBO / N); the enclosing
        ``poutine.scale(1/N)`` (train_script.py:365) restores the reference's loss."""
        if _HAVE_PYRO:
            pyro.module("gdrf_b200", self)     # param sites: SVI's optimiser only steps what the trace registered
            e = self.elbo(xs, ws)
            pyro.factor("gdrf_elbo", e * xs.shape[0])
            return ws
        self._last_elbo = self.elbo(xs, ws)
        return ws

    def guide(self, xs, ws, subsample=False):
        """sparse_gdrf.py:375-409.  Every guide site is reparameterised and folded into ``model``'s factor."""
        return None

    # ------------------------------------------------------------------ evaluation path (abstract_gdrf.py:113-139)
    def log_topic_probs(self, xs):
        xs = xs.to(self.device)
        self._check_Xnew_shape(xs)
        return marginal_mean(self._scaled(xs), self._inducing_points, self._kernel.variance,
                             self._kernel.lengthscale, self.u_loc, self._kernel_kind, self._jitter, self._maxjitter,
                             scale_mixture=self._scale_mixture)

    def topic_probs(self, xs):
        return torch.softmax(self.log_topic_probs(xs), -2).T

    def word_probs(self, xs):
        return self.topic_probs(xs) @ self.word_topic_matrix

    def ml_topics(self, xs):
        return torch.argmax(self.log_topic_probs(xs), dim=-2)

    def ml_words(self, xs):
        return torch.argmax(self.word_probs(xs), dim=-2)

    def perplexity(self, x, w):
        return perplexity_from_mean(self.log_topic_probs(x), w.to(self.device), self.word_topic_matrix)

    def forward(self, Xnew, full_cov=False):
        """sparse_gdrf.py:277-319: (loc, var) of the GP marginal at Xnew (the mean function is zero)."""
        if full_cov:
            raise NotImplementedError("full_cov=True is not accelerated")
        Xnew = Xnew.to(self.device)
        self._check_Xnew_shape(Xnew)
        x = self._scaled(Xnew)
        params = (self._inducing_points, self._kernel.variance, self._kernel.lengthscale, self.u_loc, self.u_scale_tril)
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            # differentiable like the reference's (torch autograd through gp.util.conditional): gdrf_moments_vjp
            return MarginalMoments.apply(x, *params, _lib.KERNEL_IDS[self._kernel_kind], self._jitter, self._maxjitter,
                                         _lib.FLAG_CHOL_FP32_STATUS, 0, self._scale_mixture, False)
        with torch.no_grad():
            return marginal_moments(x, *params, self._kernel_kind, self._jitter, self._maxjitter,
                                    scale_mixture=self._scale_mixture)

    def artifacts(self, xs, ws, all: bool = False):
        ret = {"kernel variance": self.kernel_variance, "kernel lengthscale": self.kernel_lengthscale}
        if not self._fixed_inducing_points:
            ret["inducing_points"] = self._inducing_points.detach().cpu().numpy()
        return ret
