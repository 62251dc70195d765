"""Where the forward's error sits (development tool): marginal moments of the CUDA path, read back in fp64
(gdrf_marginal_moments_f64), against the fp64 oracle -- bias and spread of the relative error of f_var (whose value
the model uses as a *scale* of the guide's draw, sparse_gdrf.py:403-405) and of f_loc, per forward variant.

    python tests/numerics_probe.py [--out gpurun_out/numerics_probe.json]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oracle import gdrf_oracle as O  # noqa: E402

CASES = {
    "c3_shape": dict(N=8192, D=2, K=16, V=128, grid=[16, 16], kernel="rbf", seed=61),
    "c4_shape": dict(N=4096, D=3, K=32, V=512, grid=[16, 8, 8], kernel="rbf", seed=52),
    "c2_shape": dict(N=4096, D=1, K=8, V=174, grid=[1000], kernel="matern32", seed=71),
    "c5_shape": dict(N=1024, D=3, K=64, V=1024, grid=[16, 16, 8], kernel="matern52", seed=82),
}


def main():
    from gdrf_b200 import _lib
    from gdrf_b200.elbo import marginal_moments
    out_path = sys.argv[sys.argv.index("--out") + 1] if "--out" in sys.argv else None
    base = _lib.FLAG_CHOL_FP32_STATUS
    variants = {"default (corrections first)": base,
                "interleaved MMAs": base | _lib.FLAG_INTERLEAVED_MMAS,
                "segmented accumulation": base | _lib.FLAG_SEGMENTED_FWD,
                "bf16x6": base | _lib.FLAG_FWD_BF16,
                "plain-FMA checker, fp16 planes": base | _lib.FLAG_REF_G[1] | _lib.FLAG_REF_G[2]}
    res = []
    for name, kw in CASES.items():
        inp = O.make_problem(**kw)
        with torch.no_grad():
            o = O.elbo_terms(inp.to(torch.float64), twice=False)
            o32 = O.elbo_terms(inp, twice=False)
        fl64, fv64 = o["f_loc"], o["f_var"]
        c = lambda t: t.cuda()
        row = {"case": name, "fvar_mean": float(fv64.mean()), "floc_rms": float(fl64.pow(2).mean().sqrt())}
        e = (o32["f_var"].double() - fv64) / fv64
        row["fp32 oracle"] = {"fvar_bias": float(e.mean()), "fvar_std": float(e.std()), "fvar_max": float(e.abs().max()),
                              "floc_rel": O.rel_err(o32["f_loc"], fl64)}
        for vname, fl in variants.items():
            a, b = marginal_moments(c(inp.xs), c(inp.Z), c(inp.variance), c(inp.lengthscale), c(inp.u_loc),
                                    c(inp.u_scale_tril), inp.kernel, inp.jitter, inp.maxjitter, flags=fl,
                                    dtype=torch.float64)
            e = (b.cpu() - fv64) / fv64
            row[vname] = {"fvar_bias": float(e.mean()), "fvar_std": float(e.std()), "fvar_max": float(e.abs().max()),
                          "floc_rel": O.rel_err(a.cpu(), fl64)}
        print(json.dumps(row), flush=True)
        res.append(row)
    if out_path:
        json.dump(res, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
