"""Full-size oracle fixtures: tests/golden/full_*.npz.

TEST INFRASTRUCTURE.  The fp64 and fp32 oracles (oracle/gdrf_oracle.py) are evaluated HERE, in observation chunks,
on problems too large to re-evaluate on the GPU box inside a test (BASELINE.json configs at N = 100 000 / 4 096), and
what they computed is committed: the four ELBO terms, every small gradient in full, and -- because d loss / d
u_scale_tril is K M^2 numbers (134 MB at C4) -- a seeded uniform sample of its lower-triangular entries, which gives
an unbiased estimate of the norm-wise relative error.  The inputs are not stored: ``make_problem`` regenerates them
from the seed (same torch build on the GPU box).

    python -m oracle.make_fullsize_fixtures [case ...]
"""
from __future__ import annotations

import os
import sys
import time

import numpy as np
import torch

from oracle import gdrf_oracle as O

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

CASES = {
    # name: (make_problem kwargs, oracle chunk rows)
    "full_c2_100k": (dict(N=100_000, D=1, K=8, V=174, grid=[1000], kernel="matern32", seed=72), 10_000),
    "full_c3_100k": (dict(N=100_000, D=2, K=16, V=128, grid=[16, 16], kernel="rbf", seed=61), 20_000),
    "full_c4_100k": (dict(N=100_000, D=3, K=32, V=512, grid=[16, 8, 8], kernel="rbf", seed=53), 2_000),
    "full_c5_4096": (dict(N=4_096, D=3, K=64, V=1024, grid=[16, 16, 8], kernel="matern52", seed=83), 512),
}
N_SAMPLE = 100_000


def sample_index(K: int, M: int, n: int = N_SAMPLE, seed: int = 99):
    """Seeded uniform sample of (k, i, j), j <= i, of the K x M x M lower triangles (with replacement)."""
    g = torch.Generator().manual_seed(seed)
    k = torch.randint(0, K, (n,), generator=g)
    i = torch.randint(0, M, (n,), generator=g)
    j = torch.randint(0, M, (n,), generator=g)
    lo, hi = torch.minimum(i, j), torch.maximum(i, j)
    return k, hi, lo


def chunked(inp: O.OracleInputs, rows: int, dtype):
    """Sum of the chunk ELBOs / gradients (observations are independent given the parameters; the Dirichlet prior is
    counted once)."""
    N = inp.xs.shape[0]
    terms = {k: 0.0 for k in ("lp_mu", "lq", "ll")}
    grads, lp_phi, nj = None, None, None
    for lo in range(0, N, rows):
        hi = min(N, lo + rows)
        sub = O.OracleInputs(**{**inp.__dict__, "xs": inp.xs[lo:hi], "ws": inp.ws[lo:hi], "eps": inp.eps[:, lo:hi],
                                "n_global": N})
        o, g = O.loss_and_grads(sub.to(dtype), twice=False, include_prior=(lo == 0))
        for k in terms:
            terms[k] += float(o[k].double().item())
        if lp_phi is None:
            lp_phi, nj = float(o["lp_phi"].double().item()), int(o["njitter"])
        grads = {k: v.double() for k, v in g.items()} if grads is None else {k: grads[k] + g[k].double() for k in g}
    terms["lp_phi"] = lp_phi
    return terms, grads, nj


def make(name: str) -> None:
    kw, rows = CASES[name]
    t0 = time.time()
    inp = O.make_problem(**kw)
    K, M = inp.u_loc.shape
    out = {"spec": np.array([repr(kw)]), "rows": np.array(rows)}
    kk, ii, jj = sample_index(K, M)
    for tag, dtype in (("f64", torch.float64), ("f32", torch.float32)):
        terms, g, nj = chunked(inp, rows, dtype)
        print(name, tag, f"{time.time() - t0:.0f}s", terms, flush=True)
        for k, v in terms.items():
            out[f"{tag}_{k}"] = np.array(v)
        out[f"{tag}_njitter"] = np.array(nj)
        for k, v in g.items():
            if k == "u_scale_tril":
                out[f"{tag}_grad_u_scale_tril_sample"] = v[kk, ii, jj].numpy().astype(np.float64 if tag == "f64" else np.float32)
                out[f"{tag}_grad_u_scale_tril_norm"] = np.array(v.norm().item())
            else:
                out[f"{tag}_grad_{k}"] = v.numpy().astype(np.float64 if tag == "f64" else np.float32)
    np.savez_compressed(os.path.join(GOLDEN, f"{name}.npz"), **out)
    print(name, "done", f"{time.time() - t0:.0f}s", flush=True)


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    for n in (sys.argv[1:] or list(CASES)):
        make(n)
