"""Runs the reference's OWN, unmodified model code (``/root/reference/gdrf/models/*.py``) and writes what it
computes as fixtures under tests/golden/ref_*.npz.      TEST INFRASTRUCTURE ONLY.

pyro-ppl is absent from this container, so ``import pyro`` resolves to oracle/pyro_shim (a restatement of the few
pyro primitives the path touches -- see its docstring).  Everything the reference itself implements is executed
from its source: the ``SparseMultinomialGDRF`` constructor (grid inducing points, PyroParam constraints,
``make_wt_matrix``, ``validate_dirichlet_param``), ``scale_decorator`` / ``scale_context``, ``jittercholesky``,
``model`` / ``guide`` and the evaluation methods; torch supplies the distributions, constraint transforms and
linear algebra.  The loss is ``Trace_ELBO`` over ``poutine.scale(model, 1/N)``, ``poutine.scale(guide, 1/N)`` as
in ``gdrf/train_script.py:365-371``; gradients are with respect to the *unconstrained* parameters (what
``SVI.step`` differentiates).

``gdrf/__init__.py`` imports the CLI (fire, holoviews, wandb ... absent), so the package object is created by hand
and only ``gdrf.models`` is imported.

    python oracle/make_ref_fixtures.py        # needs /root/reference; the fixtures are committed
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("GDRF_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def import_reference_models():
    sys.path.insert(0, os.path.join(HERE, "pyro_shim"))
    import pyro  # noqa: F401  (the shim)
    pkg = types.ModuleType("gdrf")
    pkg.__path__ = [os.path.join(REF, "gdrf")]
    sys.modules["gdrf"] = pkg
    import gdrf.models as models
    return models


def make_data(N, D, V, seed):
    g = torch.Generator().manual_seed(seed)
    xs = torch.rand(N, D, generator=g)
    # two smooth "communities" so that the likelihood gradient has structure
    mix = torch.sigmoid(6.0 * (xs[:, :1] - 0.5))
    pa = torch.softmax(torch.randn(V, generator=g), 0)
    pb = torch.softmax(torch.randn(V, generator=g), 0)
    probs = mix * pa + (1 - mix) * pb
    counts = torch.randint(20, 60, (N,), generator=g)
    ws = torch.stack([torch.multinomial(probs[i], int(counts[i]), replacement=True, generator=g).bincount(minlength=V)
                      for i in range(N)]).to(torch.int32)
    return xs.float(), ws


CASES = {
    # name: (kernel class name, D, n_points, K, V, N, fixed_inducing_points, ARD lengthscale, jitter, maxjitter)
    "ref_rbf2d": ("RBF", 2, 5, 3, 12, 240, False, False, 1e-4, 15),
    "ref_m32_1d_fixed": ("Matern32", 1, 14, 4, 9, 200, True, False, 1e-5, 15),
    "ref_m52_3d_ard": ("Matern52", 3, 3, 2, 15, 160, False, True, 1e-4, 15),
    "ref_exp2d": ("Exponential", 2, 4, 3, 10, 150, False, False, 1e-6, 5),
    "ref_rq2d": ("RationalQuadratic", 2, 4, 3, 11, 180, False, False, 1e-5, 15),
    # an 11th entry is the world: on anything but the unit cube the reference's guide conditions on scale(scale(xs))
    # (sparse_gdrf.py:380) and its model on scale(xs) -- the drop-in's reference_double_scale=True reproduces this
    "ref_world2d": ("RBF", 2, 5, 3, 12, 220, False, False, 1e-4, 15, [(-1.0, 1.0), (0.0, 2.0)]),
}


def run_case(models, name, spec, svi_steps=3):
    import pyro
    import pyro.contrib.gp as gp
    from pyro import poutine
    from pyro.infer import SVI, Trace_ELBO
    kname, D, n_points, K, V, N, fixed, ard, jitter, maxjitter = spec[:10]
    world = spec[10] if len(spec) > 10 else [(0.0, 1.0)] * D
    seed = sum(map(ord, name))
    torch.manual_seed(seed)
    pyro.clear_param_store()
    xs, ws = make_data(N, D, V, seed)
    lo = torch.tensor([b[0] for b in world])
    xs = lo + xs * (torch.tensor([b[1] for b in world]) - lo)       # observations anywhere in the world
    ls = torch.tensor([0.35, 0.5, 0.7][:D]) if ard else torch.tensor(0.4)
    kernel = getattr(gp.kernels, kname)(D, variance=torch.tensor(1.3), lengthscale=ls)
    m = models.SparseMultinomialGDRF(
        num_observation_categories=V, num_topic_categories=K, world=world, kernel=kernel,
        dirichlet_param=0.1, n_points=n_points, fixed_inducing_points=fixed, inducing_init="grid", device="cpu",
        jitter=jitter, maxjitter=maxjitter, xs=xs, ws=ws)
    ctor_init = {k: v.detach().clone() for k, v in m.named_parameters()}      # what the reference's constructor set
    # move every parameter off its initial value (u_loc = 0, S = chol(Kuu), phi uniform are degenerate points)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        for pname, p in m.named_parameters():
            scale = {"u_loc_unconstrained": 0.5, "u_scale_tril_unconstrained": 0.05,
                     "_word_topic_matrix_map_unconstrained": 0.7, "_inducing_points_unconstrained": 0.0}.get(pname, 0.1)
            p.add_(scale * torch.randn(p.shape, generator=g))
            if pname == "_inducing_points_unconstrained":     # grid end points sit at logit(0), logit(1): pull them in
                p.clamp_(-3.0, 3.0)
    init = {k: v.detach().clone() for k, v in m.named_parameters()}

    scale = 1.0 / N
    loss_fn = Trace_ELBO()
    pyro.EPS_LOG.clear()
    torch.manual_seed(seed + 2)
    loss = loss_fn.differentiable_loss(poutine.scale(m.model, scale=scale), poutine.scale(m.guide, scale=scale),
                                       xs=xs, ws=ws, subsample=False)
    assert len(pyro.EPS_LOG) == 1, "exactly one reparameterised draw (the guide's mu) is expected"
    eps = pyro.EPS_LOG[0].clone()
    loss.backward()
    model_trace, guide_trace = loss_fn.last_traces
    out = {"loss": np.float64(loss.item()), "eps": eps.numpy(), "xs": xs.numpy(), "ws": ws.numpy(),
           "mu": guide_trace.nodes["mu"]["value"].detach().numpy(),
           "lq": np.float64(guide_trace.nodes["mu"]["fn"].log_prob(guide_trace.nodes["mu"]["value"]).sum().item()),
           "lp_mu": np.float64(model_trace.nodes["mu"]["fn"].log_prob(model_trace.nodes["mu"]["value"]).sum().item()),
           "lp_phi": np.float64(model_trace.nodes["phi"]["fn"].log_prob(model_trace.nodes["phi"]["value"]).sum().item()),
           "ll": np.float64(model_trace.nodes["w"]["fn"].log_prob(model_trace.nodes["w"]["value"]).sum().item()),
           "beta": m._dirichlet_param.numpy(),
           "world": np.array(world, dtype=np.float64),
           "spec": np.array([kname, D, n_points, K, V, N, int(fixed), int(ard), jitter, maxjitter], dtype=object).astype(str)}
    if fixed:
        out["Z_fixed"] = m._inducing_points.detach().numpy()
    for k, v in m.named_parameters():
        out["init/" + k] = ctor_init[k].numpy()
        out["param/" + k] = init[k].numpy()
        out["grad/" + k] = v.grad.detach().numpy()
        v.grad = None
    # evaluation path, from the reference's own methods
    with torch.no_grad():
        out["log_topic_probs"] = m.log_topic_probs(xs).numpy()
        out["perplexity"] = np.float64(m.perplexity(xs, ws).item())
        out["word_probs"] = m.word_probs(xs).numpy()
    # a short SVI run exactly as train_script.py:365-371,467 wires it (Adam lr 0.01)
    svi = SVI(poutine.scale(m.model, scale=scale), poutine.scale(m.guide, scale=scale),
              pyro.optim.Adam({"lr": 0.01}), Trace_ELBO())
    losses, eps_steps = [], []
    for _ in range(svi_steps):
        pyro.EPS_LOG.clear()
        losses.append(svi.step(xs=xs, ws=ws, subsample=False))
        eps_steps.append(pyro.EPS_LOG[0].numpy())
    out["svi_losses"] = np.array(losses)
    out["svi_eps"] = np.stack(eps_steps)
    for k, v in m.named_parameters():
        out["svi_param/" + k] = v.detach().numpy()
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    print(f"{name}: loss {loss.item():.6f}  lq {out['lq']:.3f} lp_mu {out['lp_mu']:.3f} ll {out['ll']:.3f} "
          f"lp_phi {out['lp_phi']:.3f}  svi {losses}")


if __name__ == "__main__":
    models = import_reference_models()
    only = sys.argv[1:]                      # python oracle/make_ref_fixtures.py [case ...]
    for name, spec in CASES.items():
        if not only or name in only:
            run_case(models, name, spec)
