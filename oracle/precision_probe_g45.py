"""CPU probe (exploratory): do G4 (dKxz = dWtot Linv) and G5 (C5 = dWtot^T W) need three bf16 planes per operand
(24 bits, 6 MMAs per product) or would two (16 bits, 3 MMAs) do?  Emulates the rounding inside an fp64 evaluation."""
import sys
import torch
sys.path.insert(0, ".")
from oracle import gdrf_oracle as O
from oracle.precision_probe import planes

CFG = {}


class Whiten(torch.autograd.Function):
    """W = Kxz Linv^T with Linv = L^-1"""

    @staticmethod
    def forward(ctx, Kxz, Linv):
        W = Kxz @ Linv.t()
        ctx.save_for_backward(Kxz, Linv, W)
        return W

    @staticmethod
    def backward(ctx, dW):
        Kxz, Linv, W = ctx.saved_tensors
        n4, n5, dt = CFG["g4"], CFG["g5"], torch.bfloat16
        dKxz = planes(dW, n4, dt) @ planes(Linv, n4, dt)
        C5 = planes(dW, n5, dt).t() @ planes(W, n5, dt)
        L = torch.linalg.inv(Linv)
        dLinv = C5 @ L.t()                 # dLinv = dW^T Kxz = dW^T W L^T
        return dKxz, dLinv


def conditional(kind, Xnew, X, variance, lengthscale, f_loc, f_scale_tril, Lff, **kw):
    Kfs = O.kernel_matrix(kind, X, Xnew, variance, lengthscale)
    Linv = torch.linalg.solve_triangular(Lff, torch.eye(Lff.size(0), dtype=Lff.dtype), upper=False)
    W = Whiten.apply(Kfs.t(), Linv)
    loc = (W @ f_loc.t()).t()
    var = (variance - W.pow(2).sum(-1)).clamp(min=0)
    T = torch.einsum("nm,kmj->knj", W, f_scale_tril.tril())
    return loc, var + T.pow(2).sum(-1)


def run(inp, cfg):
    CFG.clear(); CFG.update(cfg)
    old = O.conditional_whitened
    O.conditional_whitened = conditional
    try:
        _, g = O.loss_and_grads(inp.to(torch.float64), twice=False)
    finally:
        O.conditional_whitened = old
    return g


if __name__ == "__main__":
    cases = [dict(N=6000, D=2, K=6, V=40, grid=[12, 12], seed=3),
             dict(N=3000, D=3, K=8, V=64, grid=[6, 6, 5], seed=5, kernel="matern32"),
             dict(N=8000, D=2, K=4, V=30, grid=[16, 16], seed=7)]
    for kw in cases:
        inp = O.make_problem(**kw)
        _, gref = O.loss_and_grads(inp.to(torch.float64), twice=False)
        _, g32 = O.loss_and_grads(inp.to(torch.float32), twice=False)
        print(kw)
        names = ("Z", "variance", "lengthscale")
        print("   fp32 oracle   ", {k: f"{O.rel_err(g32[k], gref[k]):.1e}" for k in names})
        for name, cfg in [("exact", dict(g4=9, g5=9)), ("G4 x2, G5 x3", dict(g4=2, g5=9)),
                          ("G4 x3, G5 x2", dict(g4=9, g5=2)), ("both x2", dict(g4=2, g5=2)),
                          ("both x3 (shipped)", dict(g4=3, g5=3))]:
            g = run(inp, cfg)
            print(f"   {name:18s}", {k: f"{O.rel_err(g[k], gref[k]):.1e}" for k in names})
