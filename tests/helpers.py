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
