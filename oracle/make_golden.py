"""Generates tests/golden/*.npz from the CPU oracle (and, for C1, from the reference's shipped
dataset /root/reference/data/data_2d_artificial.csv, read exactly as gdrf/train_script.py:251-273
reads it).  Run from the repo root in the build container:

    python -m oracle.make_golden

The GPU box has no /root/reference, so the inputs travel inside the fixtures.  Each fixture
holds the inputs, the fp64 oracle's loss / ELBO terms / gradients (the gate) and the fp32
oracle's (the distance that is reported next to it).
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

from .gdrf_oracle import (GRAD_NAMES, OracleInputs, grid_points, jittercholesky, kernel_matrix,
                          loss_and_grads, make_problem, perplexity)

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # name: make_problem kwargs
    "rbf2d": dict(N=1500, D=2, K=4, V=50, grid=[8, 8], kernel="rbf", seed=0),
    "m32_1d": dict(N=1000, D=1, K=3, V=20, grid=[40], kernel="matern32", seed=1),
    "m52_3d_ard": dict(N=1200, D=3, K=5, V=33, grid=[4, 3, 3], kernel="matern52", seed=2, ard=True),
    "ragged": dict(N=777, D=2, K=7, V=101, grid=[5, 5], kernel="rbf", seed=3),
    "wide": dict(N=300, D=2, K=33, V=260, grid=[8, 7], kernel="rbf", seed=4),
}


def _pack(inp: OracleInputs, store_S: bool = True):
    d = dict(xs=inp.xs.numpy(), ws=inp.ws.numpy().astype(np.int32), Z=inp.Z.numpy(),
             variance=inp.variance.numpy(), lengthscale=inp.lengthscale.numpy(),
             u_loc=inp.u_loc.numpy(), noise=inp.noise.numpy(), phi=inp.phi.numpy(),
             beta=inp.beta.numpy(), eps=inp.eps.numpy(), kernel=np.array(inp.kernel),
             jitter=np.array(inp.jitter), maxjitter=np.array(inp.maxjitter))
    if store_S:
        d["u_scale_tril"] = inp.u_scale_tril.numpy()
    return d


def _outputs(inp: OracleInputs):
    o32, g32 = loss_and_grads(inp)
    nj = int(o32["njitter"])
    o64, g64 = loss_and_grads(inp.to(torch.float64), force_njitter=nj)
    d = {"njitter": np.array(nj)}
    for tag, o, g in (("f64", o64, g64), ("f32", o32, g32)):
        for k in ("lp_mu", "lp_phi", "ll", "lq", "elbo", "loss"):
            d[f"{tag}_{k}"] = o[k].double().numpy()
        for k in GRAD_NAMES:
            d[f"{tag}_grad_{k}"] = g[k].double().numpy() if tag == "f64" else g[k].numpy()
    d["f64_f_loc"] = o64["f_loc"].numpy()
    d["f64_f_var"] = o64["f_var"].numpy()
    d["f64_perplexity"] = perplexity(inp.to(torch.float64)).numpy() if nj == 0 else np.array(np.nan)
    return d


def c1_inputs() -> OracleInputs:
    """Config C1: the reference's own CPU-runnable case at train() defaults
    (train_script.py:102-145: K from data/cfg.yaml:4, 25 inducing points per dim, RBF l=0.1, var=25,
    jitter 1e-8, maxjitter 15, beta 0.01).  Inducing grid uses inducing_init='grid' so the fixture is
    deterministic; u_scale_tril is the constructor's init (sparse_gdrf.py:100-110)."""
    import pandas as pd
    # (the reference also passes parse_dates=True; the index here is integer pixel coordinates, so
    #  it is a no-op and only produces a pandas warning)
    df = pd.read_csv("/root/reference/data/data_2d_artificial.csv", index_col=[0, 1],
                     header=0).fillna(0).astype(int)
    idx = np.array(df.index.to_list())
    idx = idx - idx.min(axis=-2, keepdims=True)
    idx = idx / idx.max(axis=-2, keepdims=True)
    xs = torch.from_numpy(idx).float()
    ws = torch.from_numpy(df.values).int()
    K, V = 5, ws.shape[1]
    Z = grid_points([25, 25])
    M = Z.shape[0]
    var, ls = torch.tensor(25.0), torch.tensor([0.1])
    L, _ = jittercholesky(kernel_matrix("rbf", Z, Z, var, ls), M, 1e-8, 15)
    S = L.expand(K, M, M).contiguous()
    eps = torch.randn(K, xs.shape[0], generator=torch.Generator().manual_seed(2024))
    return OracleInputs(xs=xs, ws=ws, Z=Z, variance=var, lengthscale=ls, u_loc=torch.zeros(K, M),
                        u_scale_tril=S, noise=torch.tensor(1.0), phi=torch.full((K, V), 1.0 / V),
                        beta=torch.full((K, V), 0.01), eps=eps, kernel="rbf", jitter=1e-8, maxjitter=15)


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    for name, kw in CASES.items():
        inp = make_problem(**kw)
        d = _pack(inp)
        d.update(_outputs(inp))
        np.savez_compressed(os.path.join(OUT, f"{name}.npz"), **d)
        print(name, "loss64", float(d["f64_loss"]), "loss32", float(d["f32_loss"]), "njitter", int(d["njitter"]))
    if os.path.exists("/root/reference/data/data_2d_artificial.csv"):
        inp = c1_inputs()
        d = _pack(inp, store_S=False)
        # S is the constructor init: one fp32 Cholesky factor shared by the K topics.  Its lower triangle is stored: the
        # fp32 factorisation of this Kuu (the one that needs five jitter escalations) is not reproducible across BLAS
        # thread counts -- rebuilt on another host it can even land on another jitter level
        M = inp.Z.shape[0]
        il = np.tril_indices(M)
        d["u_scale_tril_shared_tril"] = inp.u_scale_tril[0].numpy()[il]
        d["ws"] = d["ws"].astype(np.int16)
        d.update({k: v for k, v in _outputs(inp).items()
                  if not k.endswith("grad_u_scale_tril") and not k.endswith("f_loc") and not k.endswith("f_var")})
        np.savez_compressed(os.path.join(OUT, "c1_artificial2d.npz"), **d)
        print("c1", "loss64", float(d["f64_loss"]), "loss32", float(d["f32_loss"]), "njitter", int(d["njitter"]))
    else:
        print("reference dataset absent; c1 fixture not regenerated", file=sys.stderr)


if __name__ == "__main__":
    main()
