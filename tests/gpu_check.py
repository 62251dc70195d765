"""Exploratory GPU parity runner (development tool; the asserting version is tests/test_gpu_parity.py).

    python tests/gpu_check.py            # runs every case in its own subprocess (a trap cannot poison the rest)
    python tests/gpu_check.py case NAME FLAGS
"""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def problems():
    return {
        "rbf2d": ("golden", "rbf2d"),
        "m32_1d": ("golden", "m32_1d"),
        "m52_3d_ard": ("golden", "m52_3d_ard"),
        "ragged": ("golden", "ragged"),
        "wide": ("golden", "wide"),
        "mid512": ("make", dict(N=3000, D=2, K=3, V=40, grid=[20, 20], kernel="rbf", seed=11)),
        "mid768": ("make", dict(N=2500, D=2, K=2, V=30, grid=[25, 25], kernel="matern32", seed=12)),
    }


def run_case(name, flags, chunk_rows=0):
    import torch
    from oracle import gdrf_oracle as O
    from tests.helpers import load_golden
    import gdrf_b200
    from gdrf_b200.elbo import elbo_value_and_grads
    kind, arg = problems()[name]
    if kind == "golden":
        inp, d = load_golden(arg)
    else:
        inp = O.make_problem(**arg)
    t0 = time.time()
    o64, g64 = O.loss_and_grads(inp.to(torch.float64), twice=False)
    t_or = time.time() - t0
    o32, g32 = O.loss_and_grads(inp, twice=False)
    dev = torch.device("cuda:0")
    c = lambda t: t.to(dev)
    N = inp.xs.shape[0]
    terms, g, nj = elbo_value_and_grads(c(inp.xs), c(inp.ws), c(inp.Z), c(inp.variance), c(inp.lengthscale),
                                        c(inp.u_loc), c(inp.u_scale_tril), c(inp.noise), c(inp.phi), c(inp.beta),
                                        c(inp.eps), kernel=inp.kernel, jitter=inp.jitter, maxjitter=inp.maxjitter,
                                        flags=flags, chunk_rows=chunk_rows)
    torch.cuda.synchronize()
    t = terms.cpu()
    res = {"case": name, "flags": flags, "njitter": nj, "oracle_s": round(t_or, 2)}
    for i, k in enumerate(("lp_mu", "lq", "ll", "lp_phi")):
        ref = o64[k].item()
        res[k] = (t[i].item() - ref) / max(1.0, abs(ref))
    elbo = (t[0] + t[3] + t[2] - t[1]).item()
    res["elbo_rel"] = (elbo - o64["elbo"].item()) / abs(o64["elbo"].item())
    for k in O.GRAD_NAMES:
        ours = -g[k].cpu().double() / N          # d loss = -d ELBO / N
        res["g_" + k] = float(f"{O.rel_err(ours, g64[k]):.2e}")
        res["o32_" + k] = float(f"{O.rel_err(g32[k], g64[k]):.2e}")
    print("RESULT " + json.dumps(res))


def main():
    if len(sys.argv) >= 4 and sys.argv[1] == "case":
        run_case(sys.argv[2], int(sys.argv[3]), int(sys.argv[4]) if len(sys.argv) > 4 else 0)
        return
    from gdrf_b200 import _lib
    base = _lib.FLAG_CHOL_FP32_STATUS
    plan = []
    quick = "--quick" in sys.argv
    for name in ("rbf2d", "mid512"):
        plan.append((name, base | _lib.FLAG_REF_ALL, 0))
        plan.append((name, base, 0))
        if not quick:
            for i in range(1, 7):
                plan.append((name, base | (_lib.FLAG_REF_ALL & ~_lib.FLAG_REF_G[i]), 0))   # only Gi on tensor cores
    for name in ("m32_1d", "m52_3d_ard", "ragged", "wide", "mid768"):
        if not quick:
            plan.append((name, base | _lib.FLAG_REF_ALL, 0))
        plan.append((name, base, 0))
    plan.append(("mid512", base, 1024))    # multi-chunk streaming
    out = []
    for name, flags, chunk in plan:
        cmd = [sys.executable, os.path.abspath(__file__), "case", name, str(flags), str(chunk)]
        t0 = time.time()
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=240, cwd=ROOT)
            tail = (r.stdout + r.stderr).strip().splitlines()
            line = next((l for l in tail if l.startswith("RESULT ")), None)
            if line:
                print(line, f"# {time.time() - t0:.1f}s", flush=True)
            else:
                print(f"FAILED {name} flags={flags} chunk={chunk} rc={r.returncode}: " + " | ".join(tail[-6:]), flush=True)
        except subprocess.TimeoutExpired:
            print(f"TIMEOUT {name} flags={flags}", flush=True)


if __name__ == "__main__":
    main()
