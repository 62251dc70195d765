"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import os

import numpy as np
import torch

from oracle.gdrf_oracle import OracleInputs, jittercholesky, kernel_matrix

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = ["rbf2d", "m32_1d", "m52_3d_ard", "ragged", "wide"]


def load_golden(name):
    d = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    t = lambda k: torch.from_numpy(np.asarray(d[k]))
    Z = t("Z").float()
    var, ls = t("variance").float(), t("lengthscale").float()
    kernel = str(d["kernel"])
    jitter, maxjitter = float(d["jitter"]), int(d["maxjitter"])
    if "u_scale_tril" in d.files:
        S = t("u_scale_tril").float()
    else:  # C1: constructor init, sparse_gdrf.py:100-110
        K = t("u_loc").shape[0]
        L, _ = jittercholesky(kernel_matrix(kernel, Z, Z, var, ls), Z.shape[0], jitter, maxjitter)
        S = L.expand(K, *L.shape).contiguous()
    inp = OracleInputs(xs=t("xs").float(), ws=t("ws").int(), Z=Z, variance=var, lengthscale=ls,
                       u_loc=t("u_loc").float(), u_scale_tril=S, noise=t("noise").float(),
                       phi=t("phi").float(), beta=t("beta").float(), eps=t("eps").float(),
                       kernel=kernel, jitter=jitter, maxjitter=maxjitter)
    return inp, d


REF_CASES = ["ref_rbf2d", "ref_m32_1d_fixed", "ref_m52_3d_ard", "ref_exp2d", "ref_rq2d"]
REF_KERNELS = {"RBF": "rbf", "Matern32": "matern32", "Matern52": "matern52", "Exponential": "exponential",
               "RationalQuadratic": "rationalquadratic"}


def load_ref_fixture(name):
    """tests/golden/ref_*.npz: what the reference's own model code computed (oracle/make_ref_fixtures.py).
    Returns (spec dict, unconstrained parameters, raw npz)."""
    d = np.load(os.path.join(GOLDEN, f"{name}.npz"), allow_pickle=False)
    s = [str(x) for x in d["spec"]]
    spec = dict(kernel=REF_KERNELS[s[0]], kernel_class=s[0], D=int(s[1]), n_points=int(s[2]), K=int(s[3]), V=int(s[4]),
                N=int(s[5]), fixed=bool(int(s[6])), ard=bool(int(s[7])), jitter=float(s[8]), maxjitter=int(s[9]))
    u = {k[len("param/"):]: torch.from_numpy(d[k]) for k in d.files if k.startswith("param/")}
    return spec, u, d


def ref_constrained(spec, u, d, dtype=torch.float64):
    """Unconstrained leaves (requires_grad) and the constrained values the PyroParam transforms give."""
    from oracle import gdrf_oracle as O
    leaves = {k: v.detach().to(dtype).clone().requires_grad_(True) for k, v in u.items()}
    Z = (torch.from_numpy(d["Z_fixed"]).to(dtype) if spec["fixed"]
         else O.unit_interval(leaves["_inducing_points_unconstrained"]))
    params = {"Z": Z, "variance": O.positive(leaves["_kernel.variance_unconstrained"]),
              "lengthscale": O.positive(leaves["_kernel.lengthscale_unconstrained"]).reshape(-1),
              "u_loc": leaves["u_loc_unconstrained"], "u_scale_tril": O.lower_cholesky(leaves["u_scale_tril_unconstrained"]),
              "noise": O.positive(leaves["noise_unconstrained"]),
              "phi": O.simplex_rows(leaves["_word_topic_matrix_map_unconstrained"])}
    if "_kernel.scale_mixture_unconstrained" in leaves:
        params["scale_mixture"] = O.positive(leaves["_kernel.scale_mixture_unconstrained"])
    return leaves, params


def ref_oracle_inputs(spec, params, d, eps, dtype=torch.float64):
    det = {k: v.detach() for k, v in params.items()}
    return OracleInputs(xs=torch.from_numpy(d["xs"]).to(dtype), ws=torch.from_numpy(d["ws"]).int(), Z=det["Z"],
                        variance=det["variance"], lengthscale=det["lengthscale"], u_loc=det["u_loc"],
                        u_scale_tril=det["u_scale_tril"], noise=det["noise"], phi=det["phi"],
                        beta=torch.from_numpy(d["beta"]).to(dtype), eps=torch.as_tensor(eps).to(dtype),
                        kernel=spec["kernel"], jitter=spec["jitter"], maxjitter=spec["maxjitter"],
                        scale_mixture=det.get("scale_mixture"))
