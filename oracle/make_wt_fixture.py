"""Golden vectors for the randomised word-topic initialisation (abstract_gdrf.py:57-84), produced by the reference's
own constructor run under oracle/pyro_shim (TEST INFRASTRUCTURE; see oracle/make_ref_fixtures.py for the mechanism).

    python -m oracle.make_wt_fixture      ->  tests/golden/ref_wt_init.npz
"""
import os

import numpy as np
import torch

from oracle.make_ref_fixtures import import_reference_models, make_data

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref_wt_init.npz")


def metric(wt, model):      # deterministic, no randomness consumed: prefers peaked columns
    return float(wt.max(dim=-2).values.sum())


def main():
    models = import_reference_models()
    import pyro
    import pyro.contrib.gp as gp
    K, V, D, N = 4, 9, 2, 60
    xs, ws = make_data(N, D, V, 5)
    out = {}
    for name, kw in (("plain", dict(randomize_wt_matrix=True)),
                     ("metric", dict(randomize_wt_matrix=True, randomize_metric=metric, randomize_iters=7)),
                     ("metric_off", dict(randomize_wt_matrix=False, randomize_metric=metric))):
        torch.manual_seed(1234)
        pyro.clear_param_store()
        kernel = gp.kernels.RBF(D, variance=torch.tensor(1.3), lengthscale=torch.tensor(0.4))
        m = models.SparseMultinomialGDRF(
            num_observation_categories=V, num_topic_categories=K, world=[(0.0, 1.0)] * D, kernel=kernel,
            dirichlet_param=0.1, n_points=4, inducing_init="grid", device="cpu", jitter=1e-4, maxjitter=15,
            xs=xs, ws=ws, **kw)
        out[name] = dict(m.named_parameters())["_word_topic_matrix_map_unconstrained"].detach().numpy()
    # inducing_init="random" on a non-unit world: sorted uniform draws per dimension, scaled into the unit cube, then the
    # interval(0, 1) round trip and u_scale_tril = chol(k(Z, Z) + jitter) of THOSE points (sparse_gdrf.py:54-110)
    torch.manual_seed(4321)
    pyro.clear_param_store()
    kernel = gp.kernels.Matern32(D, variance=torch.tensor(2.0), lengthscale=torch.tensor(0.5))
    xs2 = xs * torch.tensor([2.0, 3.0]) + torch.tensor([1.0, -1.0])
    m = models.SparseMultinomialGDRF(
        num_observation_categories=V, num_topic_categories=K, world=[(1.0, 3.0), (-1.0, 2.0)], kernel=kernel,
        dirichlet_param=0.1, n_points=[3, 4], inducing_init="random", device="cpu", jitter=1e-4, maxjitter=15,
        xs=xs2, ws=ws)
    for k, v in m.named_parameters():
        out["random_init/" + k] = v.detach().numpy()
    np.savez(OUT, **out)
    print("wrote", OUT, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
