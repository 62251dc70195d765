"""Three-way parity table (development / documentation tool; the asserting version is tests/test_gpu_parity.py).

For every case and gradient tensor: CUDA vs fp64 oracle, fp32 oracle vs fp64 oracle (the reference's own arithmetic),
CUDA vs fp32 oracle -- norm-wise relative errors.  north_star names the fp32 reference; the fp64 oracle is the truth
both are measured against.

    python tests/parity_table.py [--flags F] [--out gpurun_out/parity.json] [case ...]
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from oracle import gdrf_oracle as O  # noqa: E402
from tests.helpers import FULL_CASES, GOLDEN_CASES, fullsize_errors, load_fullsize, load_golden  # noqa: E402

LIVE = {
    "smoke_700": (dict(N=700, D=2, K=3, V=24, grid=[6, 6], kernel="rbf", seed=5), 700),
    "sweep_300": (dict(N=300, D=2, K=3, V=2, grid=[5, 7], kernel="matern52", seed=400), 300),
    "c2_shape_3000": (dict(N=3000, D=1, K=8, V=174, grid=[1000], kernel="matern32", seed=71), 3000),
    "c3_shape_20k": (dict(N=20000, D=2, K=16, V=128, grid=[16, 16], kernel="rbf", seed=61), 20000),
    "c4_shape_6000": (dict(N=6000, D=3, K=32, V=512, grid=[16, 8, 8], kernel="rbf", seed=52), 2000),
    "c5_shape_768": (dict(N=768, D=3, K=64, V=1024, grid=[16, 16, 8], kernel="matern52", seed=82), 768),
}


def run_cuda(inp, flags=None):
    from gdrf_b200 import _lib
    from gdrf_b200.elbo import elbo_value_and_grads
    dev = torch.device("cuda:0")
    c = lambda t: t.to(dev)
    fl = _lib.FLAG_CHOL_FP32_STATUS if flags is None else flags
    terms, g, nj = elbo_value_and_grads(c(inp.xs), c(inp.ws), c(inp.Z), c(inp.variance), c(inp.lengthscale),
                                        c(inp.u_loc), c(inp.u_scale_tril), c(inp.noise), c(inp.phi), c(inp.beta),
                                        c(inp.eps), kernel=inp.kernel, jitter=inp.jitter, maxjitter=inp.maxjitter,
                                        flags=fl, scale_mixture=None if inp.scale_mixture is None else c(inp.scale_mixture))
    torch.cuda.synchronize()
    return terms.cpu(), {k: v.cpu().double() for k, v in g.items()}, nj


def three_way(ours, r64, r32):
    return (O.rel_err(ours, r64), O.rel_err(r32, r64), O.rel_err(ours, r32))


def case_rows(name, flags):
    t0 = time.time()
    if name in FULL_CASES:
        inp, d = load_fullsize(name)
        N = inp.xs.shape[0]
        t, g, nj = run_cuda(inp, flags)
        rows = fullsize_errors(g, d, N)
        e64 = float(d["f64_lp_mu"] + d["f64_lp_phi"] + d["f64_ll"] - d["f64_lq"])
        e32 = float(d["f32_lp_mu"] + d["f32_lp_phi"] + d["f32_ll"] - d["f32_lq"])
    elif name in GOLDEN_CASES or name == "c1_artificial2d":
        inp, d = load_golden(name)
        N = inp.xs.shape[0]
        t, g, nj = run_cuda(inp, flags)
        rows = {}
        for k in O.GRAD_NAMES:
            if f"f64_grad_{k}" in d.files:
                rows[k] = three_way(-g[k] / N, torch.from_numpy(d[f"f64_grad_{k}"]), torch.from_numpy(d[f"f32_grad_{k}"]).double())
        e64 = -float(d["f64_loss"]) * N
        e32 = -float(d["f32_loss"]) * N if "f32_loss" in d.files else float("nan")
    else:
        kw, rows_per = LIVE[name]
        inp = O.make_problem(**kw)
        N = inp.xs.shape[0]
        from oracle.make_fullsize_fixtures import chunked
        t64, g64, _ = chunked(inp, rows_per, torch.float64)
        t32, g32, _ = chunked(inp, rows_per, torch.float32)
        t, g, nj = run_cuda(inp, flags)
        rows = {k: three_way(-g[k] / N, g64[k], g32[k]) for k in O.GRAD_NAMES}
        e64 = t64["lp_mu"] + t64["lp_phi"] + t64["ll"] - t64["lq"]
        e32 = t32["lp_mu"] + t32["lp_phi"] + t32["ll"] - t32["lq"]
    elbo = (t[0] + t[3] + t[2] - t[1]).item()
    rows["ELBO"] = (abs(elbo - e64) / abs(e64), abs(e32 - e64) / abs(e64), abs(elbo - e32) / abs(e32))
    return {"case": name, "N": N, "njitter": nj, "seconds": round(time.time() - t0, 1),
            "errors": {k: [float(f"{x:.3e}") for x in v] for k, v in rows.items()}}


def main():
    args = sys.argv[1:]
    flags, out = None, None
    while args and args[0].startswith("--"):
        if args[0] == "--flags":
            flags = int(args[1])
        elif args[0] == "--out":
            out = args[1]
        args = args[2:]
    names = args or (GOLDEN_CASES + ["c1_artificial2d"] + list(LIVE) + FULL_CASES)
    res = []
    for n in names:
        r = case_rows(n, flags)
        res.append(r)
        print(json.dumps(r), flush=True)
    print("\n| case | tensor | CUDA vs fp64 | fp32 oracle vs fp64 | CUDA vs fp32 oracle |\n|---|---|---|---|---|")
    for r in res:
        for k, (a, b, c) in r["errors"].items():
            print(f"| {r['case']} (N={r['N']}) | {k} | {a:.1e} | {b:.1e} | {c:.1e} |")
    if out:
        json.dump(res, open(out, "w"), indent=1)


if __name__ == "__main__":
    main()
