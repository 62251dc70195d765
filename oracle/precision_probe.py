"""CPU probe (exploratory, not part of the product or the tests): how many 16-bit planes do the two backward
contractions need?  Emulates  dW = sum_k g2_k (T_k S_k^T)  and  dS_k = (g2_k W)^T T_k  with operands rounded to
1 or 2 fp16 / bf16 planes inside an otherwise fp64 evaluation and reports the relative error of every gradient."""
import sys
import torch
sys.path.insert(0, ".")
from oracle import gdrf_oracle as O


def planes(x, n, dt):
    hi = x.to(dt).to(x.dtype)
    if n == 1:
        return hi
    lo = (x - hi).to(dt).to(x.dtype)
    if n == 2:
        return hi + lo
    return x


CFG = {}


class RowNormTS(torch.autograd.Function):
    """tsq[k, n] = | W[n, :] S_k |^2"""

    @staticmethod
    def forward(ctx, W, S):            # W [N, M], S [K, M, M]
        T = torch.einsum("nm,kmj->knj", W, S)
        ctx.save_for_backward(W, S, T)
        return T.pow(2).sum(-1)

    @staticmethod
    def backward(ctx, g):              # g [K, N]
        W, S, T = ctx.saved_tensors
        g2 = 2.0 * g
        c = CFG
        Tq3 = planes(T, c["T3"], c["dt"])
        Sq = planes(S, c["S"], c["dt"])
        dW = torch.einsum("kn,knj,kmj->nm", g2, Tq3, Sq)
        WG = g2.unsqueeze(-1) * W.unsqueeze(0)                       # [K, N, M]
        if c.get("wg_norm"):
            sc = 2.0 ** torch.ceil(torch.log2(WG.abs().max()))
            WGq = planes(WG / sc, c["WG"], c["dt"]) * sc
        else:
            WGq = planes(WG, c["WG"], c["dt"])
        Tq6 = planes(T, c["T6"], c["dt"])
        dS = torch.einsum("knm,knj->kmj", WGq, Tq6)
        return dW, dS


def conditional(kind, Xnew, X, variance, lengthscale, f_loc, f_scale_tril, Lff, **kw):
    Kfs = O.kernel_matrix(kind, X, Xnew, variance, lengthscale)
    W = torch.linalg.solve_triangular(Lff, Kfs, upper=False).t()
    loc = (W @ f_loc.t()).t()
    var = (variance - W.pow(2).sum(-1)).clamp(min=0)
    return loc, var + RowNormTS.apply(W, f_scale_tril.tril())


def run(inp, cfg):
    CFG.clear(); CFG.update(cfg)
    old = O.conditional_whitened
    O.conditional_whitened = conditional
    try:
        out, g = O.loss_and_grads(inp.to(torch.float64), twice=False)
    finally:
        O.conditional_whitened = old
    return g


if __name__ == "__main__":
    torch.manual_seed(0)
    cases = [dict(N=6000, D=2, K=6, V=40, grid=[12, 12], seed=3),
             dict(N=3000, D=3, K=8, V=64, grid=[6, 6, 5], seed=5, kernel="matern32")]
    for kw in cases:
        inp = O.make_problem(**kw)
        _, gref = O.loss_and_grads(inp.to(torch.float64), twice=False)
        _, g32 = O.loss_and_grads(inp.to(torch.float32), twice=False)
        print(kw)
        print("   fp32 oracle          ", {k: f"{O.rel_err(g32[k], gref[k]):.1e}" for k in O.GRAD_NAMES})
        for name, cfg in [
            ("exact planes (check)", dict(T3=9, S=9, WG=9, T6=9, dt=torch.float16)),
            ("bf16 x2 all (shipped)", dict(T3=2, S=2, WG=2, T6=2, dt=torch.bfloat16)),
            ("fp16 x1 all", dict(T3=1, S=1, WG=1, T6=1, dt=torch.float16, wg_norm=True)),
            ("fp16 T1 S2 | WG2 T1", dict(T3=1, S=2, WG=2, T6=1, dt=torch.float16, wg_norm=True)),
            ("fp16 T1 S2 | WG1 T1", dict(T3=1, S=2, WG=1, T6=1, dt=torch.float16, wg_norm=True)),
            ("fp16 T1 S1 | WG2 T1", dict(T3=1, S=1, WG=2, T6=1, dt=torch.float16, wg_norm=True)),
            ("bf16 x1 all", dict(T3=1, S=1, WG=1, T6=1, dt=torch.bfloat16)),
        ]:
            g = run(inp, cfg)
            print(f"   {name:22s}", {k: f"{O.rel_err(g[k], gref[k]):.1e}" for k in O.GRAD_NAMES})
