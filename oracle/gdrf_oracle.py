"""CPU oracle for the sparse multinomial GDRF ELBO and its gradient.

TEST INFRASTRUCTURE ONLY.  Nothing under ``gdrf_b200/`` imports this module; only
``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference``
legs of ``bench.py`` may call it, and there only as the checker / reported baseline.

PARITY STATUS -- pinned against the reference's own model code, UNPINNED for pyro-ppl itself.
The reference (san-soucie/gdrf) ships no golden vectors or asserting tests for this path
(``tests/test_gdrf.py:8-22`` asserts nothing) and part of its arithmetic lives in the
un-vendored dependency ``pyro-ppl 1.8.0`` on ``torch 1.9.1`` (``poetry.lock:1169-1170,1528-1529``),
which is not importable here.  Two things anchor this file:

  1. ``oracle/make_ref_fixtures.py`` EXECUTES the reference's unmodified ``gdrf/models/*.py``
     (constructor, model, guide, jittercholesky, scale_decorator, make_wt_matrix, evaluation
     methods, a 3-step SVI run) with ``pyro`` resolved to ``oracle/pyro_shim`` -- a restatement
     of only the pyro primitives the path touches -- and commits what it computed as
     ``tests/golden/ref_*.npz``; ``tests/test_reference_fixtures.py`` holds this oracle (and the
     CUDA drop-in) to those numbers.
  2. What the shim restates rather than executes (``pyro.contrib.gp`` kernels and ``conditional``,
     ``Trace_ELBO``, ``PyroParam``) is PARITY UNPINNED: written from the published behaviour of
     pyro-ppl 1.8, cross-checked only against code that is present -- ``torch.distributions``
     (Normal / Dirichlet / Multinomial, the classes Pyro wraps), torch's constraint registry and
     ``torch.linalg`` (``tests/test_oracle.py``).

This file restates, op for op in plain PyTorch,

  * ``gdrf/models/sparse_gdrf.py:322-409``  SparseMultinomialGDRF.model / .guide
  * ``gdrf/models/utils.py:27-40``          jittercholesky (cumulative in-place jitter)
  * ``gdrf/models/abstract_gdrf.py:17-22``  zero mean, softmax link over the topic axis
  * ``gdrf/models/abstract_gdrf.py:113-139`` topic_probs / word_probs / perplexity
  * ``gdrf/train_script.py:365-371``        poutine.scale(1/N) around model and guide
  * pyro.contrib.gp.kernels.{Isotropy,RBF,Matern32,Matern52,Exponential,RationalQuadratic}  (published algorithm)
  * pyro.contrib.gp.util.conditional(whiten=True, full_cov=False) (published algorithm)
  * pyro.infer.Trace_ELBO with fully reparameterised guide sites: loss = -(log p - log q)
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional, Tuple

import torch

KERNEL_IDS = {"rbf": 0, "matern32": 1, "matern52": 2, "exponential": 3, "rationalquadratic": 4}


# --------------------------------------------------------------------------------------
# pyro.contrib.gp.kernels.Isotropy restated (expansion form, clamp, sqrt(r2 + 1e-12))
# --------------------------------------------------------------------------------------
def square_scaled_dist(X: torch.Tensor, Z: torch.Tensor, lengthscale: torch.Tensor) -> torch.Tensor:
    sX = X / lengthscale
    sZ = Z / lengthscale
    X2 = (sX ** 2).sum(1, keepdim=True)
    Z2 = (sZ ** 2).sum(1, keepdim=True)
    XZ = sX.matmul(sZ.t())
    r2 = X2 - 2 * XZ + Z2.t()
    return r2.clamp(min=0)


def kernel_matrix(kind: str, X: torch.Tensor, Z: torch.Tensor, variance: torch.Tensor,
                  lengthscale: torch.Tensor, scale_mixture: Optional[torch.Tensor] = None) -> torch.Tensor:
    r2 = square_scaled_dist(X, Z, lengthscale)
    if kind == "rationalquadratic":      # pyro.contrib.gp.kernels.RationalQuadratic
        return variance * (1 + (0.5 / scale_mixture) * r2).pow(-scale_mixture)
    if kind == "rbf":
        return variance * torch.exp(-0.5 * r2)
    r = (r2 + 1e-12).sqrt()
    if kind == "exponential":            # pyro.contrib.gp.kernels.Exponential: variance * exp(-r)
        return variance * torch.exp(-r)
    if kind == "matern32":
        s = (3 ** 0.5) * r
        return variance * (1 + s) * torch.exp(-s)
    if kind == "matern52":
        s = (5 ** 0.5) * r
        return variance * (1 + s + (5.0 / 3.0) * r2) * torch.exp(-s)
    raise ValueError(kind)


# --------------------------------------------------------------------------------------
# gdrf/models/utils.py:27-40
# --------------------------------------------------------------------------------------
def jittercholesky(Kff: torch.Tensor, N: int, jitter: float, maxjitter: int,
                   noise: float = 0.0) -> Tuple[torch.Tensor, int]:
    """Returns (Lff, njitter).  The diagonal update is cumulative and in place, exactly as
    in the reference; it goes through a clone so autograd keeps flowing into ``Kff``."""
    njitter = 0
    Lff = None
    Kff = Kff.clone()
    Kff.view(-1)[:: N + 1] += noise
    while njitter < maxjitter:
        try:
            Kff = Kff.clone()
            Kff.view(-1)[:: N + 1] += jitter * (10 ** njitter)
            Lff = torch.linalg.cholesky(Kff)
            break
        except RuntimeError:
            njitter += 1
    if njitter >= maxjitter:
        raise RuntimeError("reached max jitter, covariance is unstable")
    return Lff, njitter


def effective_jitter(jitter: float, njitter: int) -> float:
    """Total diagonal loading after ``njitter`` failed attempts (utils.py:33 is cumulative)."""
    return sum(jitter * (10 ** i) for i in range(njitter + 1))


# --------------------------------------------------------------------------------------
# pyro.contrib.gp.util.conditional, whiten=True, full_cov=False, Lff given
# --------------------------------------------------------------------------------------
def conditional_whitened(kind, Xnew, X, variance, lengthscale, f_loc, f_scale_tril, Lff, scale_mixture=None):
    N = X.size(0)
    M = Xnew.size(0)
    latent_shape = f_loc.shape[:-1]
    Kfs = kernel_matrix(kind, X, Xnew, variance, lengthscale, scale_mixture)          # [N_ind, N_obs]
    f_loc_2D = f_loc.permute(-1, *range(len(latent_shape))).reshape(N, -1)
    S = f_scale_tril.permute(-2, -1, *range(len(latent_shape)))
    S_2D = S.reshape(N, -1)
    W = torch.linalg.solve_triangular(Lff, Kfs, upper=False).t()       # [N_obs, N_ind]
    loc = W.matmul(f_loc_2D).t().reshape(latent_shape + (M,))
    Kssdiag = variance.expand(M)
    Qssdiag = W.pow(2).sum(dim=-1)
    var = (Kssdiag - Qssdiag).clamp(min=0)
    W_S = W.matmul(S_2D).reshape((M,) + S.shape[1:])
    W_S = W_S.permute(list(range(2, W_S.dim())) + [0, 1])               # [K, N_obs, N_ind]
    var = var + W_S.pow(2).sum(dim=-1)
    return loc, var


EPS32 = 1.1920928955078125e-07   # torch.finfo(torch.float32).eps: the reference runs in fp32 (train_script.py:265)


def multinomial_log_prob(probs: torch.Tensor, value: torch.Tensor) -> torch.Tensor:
    """torch.distributions.Multinomial(probs=probs, validate_args=False).log_prob(value), restated
    (Categorical normalises probs; probs_to_logits clamps to [eps, 1 - eps]; log_prob adds the lgamma terms) with
    the clamp pinned to the *fp32* eps the reference sees, so that the fp64 evaluation of the oracle is the same
    function in higher precision (tests/test_oracle.py checks it against torch's class in fp32)."""
    p = probs / probs.sum(-1, keepdim=True)
    logits = torch.log(p.clamp(min=EPS32, max=1 - EPS32))
    v = value.to(probs.dtype)
    log_factorial_n = torch.lgamma(v.sum(-1) + 1)
    log_factorial_xs = torch.lgamma(v + 1).sum(-1)
    logits = logits.masked_fill((v == 0) & (logits == -float("inf")), 0)
    return log_factorial_n - log_factorial_xs + (logits * v).sum(-1)


@dataclass
class OracleInputs:
    xs: torch.Tensor            # [N, D]   already scaled to the unit cube
    ws: torch.Tensor            # [N, V]   int32 counts
    Z: torch.Tensor             # [M, D]
    variance: torch.Tensor      # []
    lengthscale: torch.Tensor   # [1] or [D]
    u_loc: torch.Tensor         # [K, M]
    u_scale_tril: torch.Tensor  # [K, M, M] lower
    noise: torch.Tensor         # []
    phi: torch.Tensor           # [K, V]  rows on the simplex
    beta: torch.Tensor          # [K, V]
    eps: torch.Tensor           # [K, N]  fixed standard-normal draws ("fixed posterior samples")
    kernel: str = "rbf"
    jitter: float = 1e-8
    maxjitter: int = 5
    n_global: Optional[int] = None   # the 1/N of poutine.scale; defaults to len(xs)
    scale_mixture: Optional[torch.Tensor] = None   # RationalQuadratic only
    # what the GUIDE conditions on when it differs from the model's xs: the reference's guide applies scale() a second
    # time (sparse_gdrf.py:380), so on a world other than the unit cube it sees scale(scale(xs)) while the model sees
    # scale(xs).  None: the same inputs (the unit-cube data of train(), train_script.py:263-271).
    xs_guide: Optional[torch.Tensor] = None

    def to(self, dtype: torch.dtype) -> "OracleInputs":
        f = lambda t: t.detach().to(dtype)
        return OracleInputs(f(self.xs), self.ws, f(self.Z), f(self.variance), f(self.lengthscale),
                            f(self.u_loc), f(self.u_scale_tril), f(self.noise), f(self.phi),
                            f(self.beta), f(self.eps), self.kernel, self.jitter, self.maxjitter,
                            self.n_global, None if self.scale_mixture is None else f(self.scale_mixture),
                            None if self.xs_guide is None else f(self.xs_guide))


GRAD_NAMES = ("Z", "variance", "lengthscale", "u_loc", "u_scale_tril", "noise", "phi")


def param_names(inp: "OracleInputs"):
    return GRAD_NAMES + (("scale_mixture",) if inp.scale_mixture is not None else ())


def _one_conditional(inp: OracleInputs, p: Dict[str, torch.Tensor], force_njitter=None, xs=None):
    xs = inp.xs if xs is None else xs
    sm = p.get("scale_mixture", inp.scale_mixture)
    Kuu = kernel_matrix(inp.kernel, p["Z"], p["Z"], p["variance"], p["lengthscale"], sm).contiguous()
    M = Kuu.size(0)
    if force_njitter is None:
        Luu, nj = jittercholesky(Kuu, M, inp.jitter, inp.maxjitter)
    else:
        nj = force_njitter
        Kj = Kuu + effective_jitter(inp.jitter, nj) * torch.eye(M, dtype=Kuu.dtype)
        Luu = torch.linalg.cholesky(Kj)
    f_loc, f_var = conditional_whitened(inp.kernel, xs, p["Z"], p["variance"], p["lengthscale"],
                                        p["u_loc"], p["u_scale_tril"], Luu, **({} if sm is None else {"scale_mixture": sm}))
    # zero mean function (abstract_gdrf.py:17-18) broadcast-added
    f_loc = f_loc + torch.zeros(xs.shape[:-1], dtype=f_loc.dtype)
    return f_loc, f_var, nj


def elbo_terms(inp: OracleInputs, params: Optional[Dict[str, torch.Tensor]] = None,
               twice: bool = True, force_njitter: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """ELBO pieces exactly as model/guide produce them.

    ``twice=True`` evaluates the conditional separately for guide and model, as the reference
    does (sparse_gdrf.py:334-344 and :384-394); the numbers are identical either way.
    ``force_njitter`` pins the escalation level (used to compare fp64 against an fp32 run that
    needed more escalations)."""
    p = params if params is not None else {k: getattr(inp, k) for k in param_names(inp)}
    N = inp.xs.size(0)
    # ---- guide (sparse_gdrf.py:375-409) ----
    f_loc_g, f_var_g, nj = _one_conditional(inp, p, force_njitter, xs=inp.xs_guide)
    mu = f_loc_g + f_var_g * inp.eps                      # Normal(f_loc, f_var).rsample()
    q_mu = torch.distributions.Normal(f_loc_g, f_var_g)
    lq = q_mu.log_prob(mu).sum()
    # Delta(phi).to_event(1): log-prob 0, value = the parameter
    phi = p["phi"]
    # ---- model (sparse_gdrf.py:323-373) replayed against the guide's mu / phi ----
    if twice or inp.xs_guide is not None:
        f_loc_m, f_var_m, _ = _one_conditional(inp, p, nj if force_njitter is None else force_njitter)
    else:
        f_loc_m, f_var_m = f_loc_g, f_var_g
    lp_mu = torch.distributions.Normal(f_loc_m, f_var_m + p["noise"]).log_prob(mu).sum()
    lp_phi = torch.distributions.Dirichlet(inp.beta, validate_args=False).log_prob(phi).sum()
    topic_probs = torch.softmax(mu, -2).transpose(-2, -1)
    probs = torch.matmul(topic_probs, phi)
    ll = multinomial_log_prob(probs, inp.ws).sum()
    n_scale = inp.n_global if inp.n_global is not None else N
    elbo = lp_mu + lp_phi + ll - lq
    return {"lp_mu": lp_mu, "lp_phi": lp_phi, "ll": ll, "lq": lq, "elbo": elbo,
            "elbo_over_n": elbo / n_scale, "loss": -elbo / n_scale, "njitter": nj,
            "f_loc": f_loc_g.detach(), "f_var": f_var_g.detach(), "mu": mu.detach()}


def loss_and_grads(inp: OracleInputs, twice: bool = True, force_njitter: Optional[int] = None,
                   include_prior: bool = True):
    """loss = -ELBO/N and d loss / d (constrained parameter) for every name in GRAD_NAMES."""
    names = param_names(inp)
    params = {k: getattr(inp, k).detach().clone().requires_grad_(True) for k in names}
    out = elbo_terms(inp, params, twice=twice, force_njitter=force_njitter)
    loss = out["loss"]
    if not include_prior:
        n_scale = inp.n_global if inp.n_global is not None else inp.xs.size(0)
        loss = loss + out["lp_phi"] / n_scale
    grads = torch.autograd.grad(loss, [params[k] for k in names], allow_unused=True)
    g = {k: (torch.zeros_like(params[k]) if gi is None else gi.detach()) for k, gi in zip(names, grads)}
    g["u_scale_tril"] = g["u_scale_tril"].tril()
    return {k: (v.detach() if torch.is_tensor(v) else v) for k, v in out.items()}, g


# --------------------------------------------------------------------------------------
# evaluation path (abstract_gdrf.py:113-139, sparse_gdrf.py:161-186)
# --------------------------------------------------------------------------------------
def log_topic_probs(inp: OracleInputs) -> torch.Tensor:
    p = {k: getattr(inp, k) for k in param_names(inp)}
    f_loc, _, _ = _one_conditional(inp, p)
    return f_loc


def perplexity(inp: OracleInputs) -> torch.Tensor:
    tp = torch.softmax(log_topic_probs(inp), -2).T
    wp = tp @ inp.phi
    w = inp.ws
    return ((w * wp.log()).sum() / -w.sum()).exp()


# --------------------------------------------------------------------------------------
# constraint maps used by PyroParam (torch.distributions.constraint_registry.transform_to)
# --------------------------------------------------------------------------------------
def positive(u):            # ExpTransform
    return u.exp()


def lower_cholesky(u):      # LowerCholeskyTransform
    return u.tril(-1) + u.diagonal(dim1=-2, dim2=-1).exp().diag_embed()


def unit_interval(u):       # SigmoidTransform (interval(0,1) -> affine is identity)
    return torch.sigmoid(u)


def simplex_rows(u):        # SoftmaxTransform per row
    return torch.softmax(u, -1)


# --------------------------------------------------------------------------------------
# synthetic problems
# --------------------------------------------------------------------------------------
def grid_points(n_per_dim) -> torch.Tensor:
    axes = [torch.linspace(0.0, 1.0, n) if n > 1 else torch.tensor([0.5]) for n in n_per_dim]
    mesh = torch.meshgrid(*axes, indexing="ij")
    return torch.stack([m.flatten() for m in mesh]).T.contiguous().float()


def make_problem(N: int, D: int, K: int, V: int, grid, kernel: str = "rbf", seed: int = 0,
                 jitter: float = 1e-4, maxjitter: int = 15, ls_factor: float = 0.75,
                 variance: float = 25.0, ard: bool = False, count_scale: int = 1) -> OracleInputs:
    """Synthetic inputs of SURVEY.md 8(d): uniform xs, tensor-product inducing grid, Dirichlet-mixture
    counts, perturbed-Cholesky u_scale_tril so every gradient is exercised."""
    g = torch.Generator().manual_seed(1234 + seed)
    xs = torch.rand(N, D, generator=g)
    Z = grid_points(grid)
    M = Z.size(0)
    spacing = min(1.0 / (n - 1) for n in grid if n > 1)
    ls = torch.full((D if ard else 1,), ls_factor * spacing)
    if ard:
        ls = ls * torch.linspace(1.0, 1.3, D)
    var = torch.tensor(float(variance))
    g2 = torch.Generator().manual_seed(4321 + seed)
    with torch.random.fork_rng():                       # Dirichlet.sample() draws from the global RNG
        torch.manual_seed(4321 + seed)
        theta_star = (torch.distributions.Dirichlet(torch.full((K,), 0.3)).sample((N,))
                      if K > 1 else torch.ones(N, 1))
        phi_star = torch.distributions.Dirichlet(torch.full((V,), 0.1)).sample((K,))
    probs = theta_star @ phi_star
    counts = torch.randint(V * count_scale, 10 * V * count_scale, (N,), generator=g2)
    ws = torch.zeros(N, V, dtype=torch.int32)
    idx = torch.multinomial(probs, int(counts.max()), replacement=True, generator=g2)
    mask = torch.arange(idx.size(1))[None, :] < counts[:, None]
    ws.scatter_add_(1, idx, mask.to(torch.int32))
    u_loc = 0.5 * torch.randn(K, M, generator=torch.Generator().manual_seed(7 + seed))
    sm = torch.tensor(1.7) if kernel == "rationalquadratic" else None
    Kuu = kernel_matrix(kernel, Z.double(), Z.double(), var.double(), ls.double(), None if sm is None else sm.double())
    L0 = torch.linalg.cholesky(Kuu + jitter * torch.eye(M, dtype=torch.float64)).float()
    S = L0.expand(K, M, M) + 0.05 * torch.randn(K, M, M, generator=torch.Generator().manual_seed(8 + seed)).tril()
    S = S.tril().contiguous()
    d = S.diagonal(dim1=-2, dim2=-1)
    d.copy_(d.abs().clamp(min=1e-3))
    phi = torch.softmax(torch.randn(K, V, generator=torch.Generator().manual_seed(9 + seed)), -1)
    beta = torch.full((K, V), 0.01)
    eps = torch.randn(K, N, generator=torch.Generator().manual_seed(2024 + seed))
    return OracleInputs(xs=xs, ws=ws, Z=Z, variance=var, lengthscale=ls, u_loc=u_loc, u_scale_tril=S,
                        noise=torch.tensor(1.0), phi=phi, beta=beta, eps=eps, kernel=kernel,
                        jitter=jitter, maxjitter=maxjitter, scale_mixture=sm)


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """Norm-wise relative error ||a-b|| / ||b|| in float64."""
    a = a.double().flatten()
    b = b.double().flatten()
    den = b.norm().item()
    return (a - b).norm().item() / (den if den > 0 else 1.0)
