"""Known-answer checks that pin the CPU oracle against code that exists here:
torch.distributions / torch.linalg (the classes Pyro wraps), scipy, closed forms, finite
differences, and the committed golden vectors.  CPU only."""
import math

import numpy as np
import pytest
import torch

from oracle import gdrf_oracle as O
from tests.helpers import GOLDEN_CASES, load_golden

torch.set_default_dtype(torch.float32)


def _direct_kernel(kind, X, Z, var, ls):
    d = (X[:, None, :] - Z[None, :, :]) / ls
    r2 = (d * d).sum(-1)
    r = r2.sqrt()
    if kind == "rbf":
        return var * torch.exp(-0.5 * r2)
    if kind == "exponential":
        return var * torch.exp(-r)
    if kind == "matern32":
        return var * (1 + math.sqrt(3) * r) * torch.exp(-math.sqrt(3) * r)
    return var * (1 + math.sqrt(5) * r + 5.0 / 3.0 * r2) * torch.exp(-math.sqrt(5) * r)


@pytest.mark.parametrize("kind", ["rbf", "matern32", "matern52", "exponential"])
def test_kernels_match_closed_form(kind):
    g = torch.Generator().manual_seed(0)
    X = torch.rand(50, 3, generator=g, dtype=torch.float64)
    Z = torch.rand(20, 3, generator=g, dtype=torch.float64)
    var = torch.tensor(2.5, dtype=torch.float64)
    ls = torch.tensor([0.3, 0.2, 0.5], dtype=torch.float64)
    a = O.kernel_matrix(kind, X, Z, var, ls)
    b = _direct_kernel(kind, X, Z, var, ls)
    assert torch.allclose(a, b, rtol=1e-5, atol=1e-6)   # sqrt(r2+1e-12) vs sqrt(r2)


def test_conditional_matches_dense_formulas():
    inp = O.make_problem(N=60, D=2, K=3, V=7, grid=[4, 4]).to(torch.float64)
    Kuu = O.kernel_matrix("rbf", inp.Z, inp.Z, inp.variance, inp.lengthscale)
    Kuu = Kuu + inp.jitter * torch.eye(16, dtype=torch.float64)
    L = torch.linalg.cholesky(Kuu)
    f_loc, f_var = O.conditional_whitened("rbf", inp.xs, inp.Z, inp.variance, inp.lengthscale,
                                          inp.u_loc, inp.u_scale_tril, L)
    Kxz = O.kernel_matrix("rbf", inp.xs, inp.Z, inp.variance, inp.lengthscale)
    Linv = torch.linalg.inv(L)
    for k in range(3):
        loc = Kxz @ Linv.T @ inp.u_loc[k]
        A = Linv.T @ inp.u_scale_tril[k]
        var = inp.variance - ((Kxz @ torch.linalg.inv(Kuu)) * Kxz).sum(-1) + ((Kxz @ A) ** 2).sum(-1)
        assert torch.allclose(f_loc[k], loc, rtol=1e-9, atol=1e-9)
        assert torch.allclose(f_var[k], var, rtol=1e-8, atol=1e-9)


def test_multinomial_and_dirichlet_terms_match_manual_and_scipy():
    from scipy.special import gammaln
    from scipy.stats import dirichlet
    inp = O.make_problem(N=40, D=2, K=3, V=9, grid=[3, 3]).to(torch.float64)
    out = O.elbo_terms(inp)
    mu = out["mu"]
    theta = torch.softmax(mu, 0).T
    p = theta @ inp.phi
    p = p / p.sum(-1, keepdim=True)
    w = inp.ws.double()
    ll = (torch.lgamma(w.sum(-1) + 1) - torch.lgamma(w + 1).sum(-1) + (w * p.log()).sum(-1)).sum()
    assert abs(ll.item() - out["ll"].item()) < 1e-8 * abs(ll.item())
    lp = sum(dirichlet.logpdf(inp.phi[k].numpy() / inp.phi[k].numpy().sum(), inp.beta[k].numpy())
             for k in range(3))
    assert abs(lp - out["lp_phi"].item()) < 1e-8 * abs(lp)
    assert np.isfinite(gammaln(1.0))


def test_multinomial_restatement_equals_torch_class_in_fp32():
    g = torch.Generator().manual_seed(3)
    probs = torch.rand(40, 9, generator=g) + 1e-3
    probs[0] = torch.tensor([1.0] + [0.0] * 8)          # exercises both clamp ends
    ws = torch.randint(0, 7, (40, 9), generator=g).int()
    ws[0, 1:] = 0
    ref = torch.distributions.Multinomial(probs=probs, validate_args=False).log_prob(ws)
    assert torch.allclose(O.multinomial_log_prob(probs, ws), ref, rtol=1e-6, atol=1e-5)
    assert O.EPS32 == torch.finfo(torch.float32).eps


def test_elbo_matches_closed_form():
    inp = O.make_problem(N=80, D=1, K=2, V=5, grid=[6], kernel="matern32").to(torch.float64)
    out = O.elbo_terms(inp)
    fv, e, nz = out["f_var"], inp.eps, inp.noise
    gp = (-(fv + nz).log() - (fv * e) ** 2 / (2 * (fv + nz) ** 2) + fv.log() + 0.5 * e ** 2).sum()
    assert abs((out["lp_mu"] - out["lq"]).item() - gp.item()) < 1e-9 * max(1.0, abs(gp.item()))
    assert abs(out["elbo"].item() - (gp + out["ll"] + out["lp_phi"]).item()) < 1e-9 * abs(out["elbo"].item())
    assert abs(out["loss"].item() + out["elbo"].item() / 80) < 1e-12


def test_twice_equals_once():
    inp = O.make_problem(N=50, D=2, K=3, V=6, grid=[3, 3])
    a, ga = O.loss_and_grads(inp, twice=True)
    b, gb = O.loss_and_grads(inp, twice=False)
    assert a["loss"].item() == pytest.approx(b["loss"].item(), rel=1e-6)
    for k in ga:
        assert O.rel_err(ga[k], gb[k]) < 1e-3 or ga[k].norm() < 1e-6


def test_gradients_match_finite_differences():
    inp = O.make_problem(N=30, D=2, K=2, V=5, grid=[3, 3]).to(torch.float64)
    _, g = O.loss_and_grads(inp)
    h = 1e-6
    for name in O.GRAD_NAMES:
        base = getattr(inp, name)
        flat_idx = [0, base.numel() // 2, base.numel() - 1]
        for i in flat_idx:
            if name == "u_scale_tril":
                r, c = np.unravel_index(i, base.shape)[-2:]
                if c > r:
                    continue
            vals = []
            for s in (+1, -1):
                p = base.clone()
                p.view(-1)[i] += s * h
                kw = {k: getattr(inp, k) for k in O.GRAD_NAMES}
                kw[name] = p
                vals.append(O.elbo_terms(inp, kw)["loss"].item())
            fd = (vals[0] - vals[1]) / (2 * h)
            an = g[name].reshape(-1)[i].item()
            assert abs(fd - an) <= 2e-5 * max(1.0, abs(an)), (name, i, fd, an)


def test_jittercholesky_is_cumulative_and_raises():
    A = torch.tensor([[1.0, 1.0], [1.0, 1.0 - 1e-3]], dtype=torch.float64)
    L, nj = O.jittercholesky(A, 2, 1e-6, 8)
    assert nj == 3                                           # 1e-6+1e-5+1e-4 fails, +1e-3 succeeds
    assert O.effective_jitter(1e-6, nj) == pytest.approx(1.111e-3)
    assert torch.allclose(L @ L.T, A + O.effective_jitter(1e-6, nj) * torch.eye(2, dtype=torch.float64))
    with pytest.raises(RuntimeError, match="reached max jitter"):
        O.jittercholesky(-torch.eye(3, dtype=torch.float64), 3, 1e-8, 5)


def test_observation_shards_sum_to_whole():
    inp = O.make_problem(N=101, D=2, K=3, V=8, grid=[3, 3]).to(torch.float64)
    full, gfull = O.loss_and_grads(inp)
    tot, gtot = 0.0, None
    for r, sl in enumerate((slice(0, 40), slice(40, 101))):
        sh = O.OracleInputs(inp.xs[sl], inp.ws[sl], inp.Z, inp.variance, inp.lengthscale, inp.u_loc,
                            inp.u_scale_tril, inp.noise, inp.phi, inp.beta, inp.eps[:, sl], inp.kernel,
                            inp.jitter, inp.maxjitter, n_global=101)
        o, g = O.loss_and_grads(sh, include_prior=(r == 0))
        lossr = o["loss"] + (0 if r == 0 else o["lp_phi"] / 101)
        tot += lossr.item()
        gtot = g if gtot is None else {k: gtot[k] + g[k] for k in g}
    assert tot == pytest.approx(full["loss"].item(), rel=1e-12)
    for k in gfull:
        assert O.rel_err(gtot[k], gfull[k]) < 1e-10


def test_constraint_maps_match_torch_registry():
    from torch.distributions import constraints, transform_to
    u = torch.randn(3, 4, 4, dtype=torch.float64)
    assert torch.allclose(O.lower_cholesky(u), transform_to(constraints.lower_cholesky)(u))
    assert torch.allclose(O.positive(u), transform_to(constraints.positive)(u))
    assert torch.allclose(O.unit_interval(u), transform_to(constraints.interval(0.0, 1.0))(u))
    v = torch.randn(5, 7, dtype=torch.float64)
    assert torch.allclose(O.simplex_rows(v), transform_to(constraints.simplex)(v))


@pytest.mark.parametrize("name", GOLDEN_CASES + ["c1_artificial2d"])
def test_oracle_reproduces_golden(name):
    inp, d = load_golden(name)
    nj = int(d["njitter"])
    o32, g32 = O.loss_and_grads(inp)
    assert int(o32["njitter"]) == nj
    assert o32["loss"].item() == pytest.approx(float(d["f32_loss"]), rel=2e-6)
    o64, g64 = O.loss_and_grads(inp.to(torch.float64), force_njitter=nj)
    assert o64["loss"].item() == pytest.approx(float(d["f64_loss"]), rel=1e-10)
    for k in ("lp_mu", "lp_phi", "ll", "lq"):
        assert o64[k].item() == pytest.approx(float(d[f"f64_{k}"]), rel=1e-9)
    for k in O.GRAD_NAMES:
        key = f"f64_grad_{k}"
        if key in d.files and nj == 0:
            assert O.rel_err(g64[k], torch.from_numpy(d[key])) < 1e-7, k
