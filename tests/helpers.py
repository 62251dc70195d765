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
    else:  # C1: constructor init, sparse_gdrf.py:100-110 -- one fp32 factor shared by the K topics
        K, M = t("u_loc").shape
        if "u_scale_tril_shared_tril" in d.files:      # as computed when the fixture was made (not reproducible across
            L = torch.zeros(M, M)                      # hosts: an fp32 Cholesky at the edge of positive definiteness)
            il = np.tril_indices(M)
            L[il[0], il[1]] = t("u_scale_tril_shared_tril").float()
        else:
            L, _ = jittercholesky(kernel_matrix(kernel, Z, Z, var, ls), M, jitter, maxjitter)
        S = L.expand(K, *L.shape).contiguous()
    inp = OracleInputs(xs=t("xs").float(), ws=t("ws").int(), Z=Z, variance=var, lengthscale=ls,
                       u_loc=t("u_loc").float(), u_scale_tril=S, noise=t("noise").float(),
                       phi=t("phi").float(), beta=t("beta").float(), eps=t("eps").float(),
                       kernel=kernel, jitter=jitter, maxjitter=maxjitter)
    return inp, d


REF_CASES = ["ref_rbf2d", "ref_m32_1d_fixed", "ref_m52_3d_ard", "ref_exp2d", "ref_rq2d", "ref_world2d"]
REF_KERNELS = {"RBF": "rbf", "Matern32": "matern32", "Matern52": "matern52", "Exponential": "exponential",
               "RationalQuadratic": "rationalquadratic"}


def load_ref_fixture(name):
    """tests/golden/ref_*.npz: what the reference's own model code computed (oracle/make_ref_fixtures.py).
    Returns (spec dict, unconstrained parameters, raw npz)."""
    d = np.load(os.path.join(GOLDEN, f"{name}.npz"), allow_pickle=False)
    s = [str(x) for x in d["spec"]]
    spec = dict(kernel=REF_KERNELS[s[0]], kernel_class=s[0], D=int(s[1]), n_points=int(s[2]), K=int(s[3]), V=int(s[4]),
                N=int(s[5]), fixed=bool(int(s[6])), ard=bool(int(s[7])), jitter=float(s[8]), maxjitter=int(s[9]))
    # the world the reference model was built on ([0, 1]^D in the older fixtures, which do not store it); xs is in
    # world coordinates
    spec["world"] = ([tuple(map(float, b)) for b in d["world"]] if "world" in d.files else [(0.0, 1.0)] * spec["D"])
    spec["unit_world"] = all(b == (0.0, 1.0) for b in spec["world"])
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
    # scale_decorator (topic_model.py:122-144) maps xs into the unit cube for model and guide; the guide then applies
    # scale() once more (sparse_gdrf.py:380) -- a no-op only on the unit world
    lo = torch.tensor([b[0] for b in spec["world"]], dtype=dtype)
    delta = torch.tensor([b[1] - b[0] for b in spec["world"]], dtype=dtype)
    xs1 = (torch.from_numpy(d["xs"]).to(dtype) - lo) / delta
    xs_guide = None if spec["unit_world"] else (xs1 - lo) / delta
    return OracleInputs(xs=xs1, xs_guide=xs_guide, ws=torch.from_numpy(d["ws"]).int(), Z=det["Z"],
                        variance=det["variance"], lengthscale=det["lengthscale"], u_loc=det["u_loc"],
                        u_scale_tril=det["u_scale_tril"], noise=det["noise"], phi=det["phi"],
                        beta=torch.from_numpy(d["beta"]).to(dtype), eps=torch.as_tensor(eps).to(dtype),
                        kernel=spec["kernel"], jitter=spec["jitter"], maxjitter=spec["maxjitter"],
                        scale_mixture=det.get("scale_mixture"))


# ---------------------------------------------------------------------------------------------------------------
# full-size fixtures (oracle/make_fullsize_fixtures.py): inputs regenerated from the seed, oracle outputs committed
# ---------------------------------------------------------------------------------------------------------------
FULL_CASES = ["full_c2_100k", "full_c3_100k", "full_c4_100k", "full_c5_4096"]


def load_fullsize(name):
    """(OracleInputs regenerated by make_problem, npz of what the fp64 / fp32 oracles computed here)."""
    from oracle.make_fullsize_fixtures import CASES
    from oracle.gdrf_oracle import make_problem
    d = np.load(os.path.join(GOLDEN, f"{name}.npz"))
    kw, _ = CASES[name]
    assert repr(kw) == str(d["spec"][0]), "fixture was generated for another problem"
    return make_problem(**kw), d


def fullsize_errors(g, d, N):
    """Norm-wise relative errors per gradient tensor: (ours vs fp64, fp32 oracle vs fp64, ours vs fp32 oracle).
    ``g``: d ELBO / d constrained (what the C ABI returns); u_scale_tril is compared on the fixture's seeded sample of
    its lower-triangular entries."""
    from oracle.gdrf_oracle import GRAD_NAMES, rel_err
    from oracle.make_fullsize_fixtures import sample_index
    out = {}
    for k in GRAD_NAMES:
        ours = -g[k].detach().cpu().double() / N
        if k == "u_scale_tril":
            kk, ii, jj = sample_index(ours.shape[0], ours.shape[1])
            ours = ours[kk, ii, jj]
            r64 = torch.from_numpy(d["f64_grad_u_scale_tril_sample"]).double()
            r32 = torch.from_numpy(d["f32_grad_u_scale_tril_sample"]).double()
        else:
            r64 = torch.from_numpy(d[f"f64_grad_{k}"]).double().reshape(ours.shape)
            r32 = torch.from_numpy(d[f"f32_grad_{k}"]).double().reshape(ours.shape)
        out[k] = (rel_err(ours, r64), rel_err(r32, r64), rel_err(ours, r32))
    return out


# ---------------------------------------------------------------------------------------------------------------
# the parity gate
# ---------------------------------------------------------------------------------------------------------------
TOL = 1e-4      # north_star: "within a stated fp32 tolerance (<= 1e-4 relative)" of the reference's ELBO and gradients


def parity_ok(err64: float, err32: float, err_vs32: float, tol: float = TOL) -> bool:
    """One rule for every gradient tensor of every case.  err64: CUDA vs the fp64 oracle (the truth both are measured
    against); err32: fp32 oracle -- the reference's own arithmetic -- vs fp64; err_vs32: CUDA vs the fp32 oracle.
    Passes when the CUDA path is within 1e-4 of the truth, or, where the fp32 reference itself is further than 1e-4
    from the truth, when it is no further from the truth than the reference is (factor 1.0), or when it is within 1e-4
    of the fp32 reference (what north_star literally names)."""
    return err64 <= max(tol, err32) or err_vs32 <= tol


def assert_parity(rows, what=""):
    """rows: {tensor: (err64, err32, err_vs32)}; prints the three-way table and asserts the gate on every row."""
    print(what, {k: tuple(f"{x:.1e}" for x in v) for k, v in rows.items()})
    bad = {k: v for k, v in rows.items() if not parity_ok(*v)}
    assert not bad, (what, bad)
