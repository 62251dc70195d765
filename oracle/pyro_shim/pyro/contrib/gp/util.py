"""gp.util.conditional: mean and (co)variance of f(Xnew) given inducing values with N(f_loc, f_scale_tril) at X."""
import torch


def conditional(Xnew, X, kernel, f_loc, f_scale_tril=None, Lff=None, full_cov=False, whiten=False, jitter=1e-6):
    N = X.size(0)
    M = Xnew.size(0)
    latent_shape = f_loc.shape[:-1]

    if Lff is None:
        Kff = kernel(X).contiguous()
        Kff.view(-1)[:: N + 1] += jitter
        Lff = torch.linalg.cholesky(Kff)
    Kfs = kernel(X, Xnew)

    # latent dimensions go last so that one triangular solve serves every latent function
    f_loc_2D = f_loc.reshape(-1, N).t()
    S_2D = None
    if f_scale_tril is not None:
        S_2D = f_scale_tril.reshape(-1, N, N).permute(1, 2, 0).reshape(N, -1)

    if whiten:
        v_2D = f_loc_2D
        W = torch.linalg.solve_triangular(Lff, Kfs, upper=False).t()
    else:
        pack = torch.cat([f_loc_2D, Kfs] + ([S_2D] if S_2D is not None else []), dim=1)
        Lffinv_pack = torch.linalg.solve_triangular(Lff, pack, upper=False)
        v_2D = Lffinv_pack[:, : f_loc_2D.size(1)]
        W = Lffinv_pack[:, f_loc_2D.size(1): f_loc_2D.size(1) + M].t()
        if S_2D is not None:
            S_2D = Lffinv_pack[:, -S_2D.size(1):]

    loc = W.matmul(v_2D).t().reshape(latent_shape + (M,))

    if full_cov:
        Kss = kernel(Xnew)
        cov = Kss - W.matmul(W.t())
    else:
        Kssdiag = kernel(Xnew, diag=True)
        Qssdiag = W.pow(2).sum(dim=-1)
        # Kss - Qss is non-negative in theory only
        var = (Kssdiag - Qssdiag).clamp(min=0)

    if S_2D is not None:
        W_S = W.matmul(S_2D).reshape((M, N) + tuple(latent_shape))
        W_S = W_S.permute(list(range(2, W_S.dim())) + [0, 1])      # latent_shape + (M, N)
        if full_cov:
            cov = cov + W_S.matmul(W_S.transpose(-2, -1))
        else:
            var = var + W_S.pow(2).sum(dim=-1)

    if full_cov:
        return loc, cov.expand(latent_shape + (M, M))
    return loc, var.expand(latent_shape + (M,))
