"""The ELBO of the sparse multinomial GDRF as one custom autograd op over the C ABI.

``GDRFElbo.apply`` replaces everything ``pyro.infer.SVI.step`` evaluates for
``SparseMultinomialGDRF.model`` / ``.guide`` (reference ``gdrf/models/sparse_gdrf.py:322-409``): kernel
matrices, the jitter-escalating Cholesky (``gdrf/models/utils.py:27-40``), the whitened sparse-GP
conditional, the reparameterised draw of ``mu``, both Normal terms, the Dirichlet prior on ``phi``, the
topic softmax, the theta-phi mixture and the multinomial log-likelihood -- and their gradients.

Inputs are the *constrained* parameter values; autograd chains through the constraint transforms
outside the op.  ``eps`` may be ``[P, K, N]``: P draws of the guide per observation (``Trace_ELBO(num_particles=P,
vectorize_particles=True)``, ``train_script.py:330-335``); value and gradient are then the mean over the particles,
computed with ONE pass of the contractions (the marginal moments do not depend on the draw and the backward
contractions are linear in the per-observation weights) and P passes of the per-observation chain.  The value returned is ELBO / n_global for the observations handed in (this rank's
shard); the reference's loss (``poutine.scale(1/N)``, ``train_script.py:365``) is its negative.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional, Tuple

import torch

from . import _lib

_WORKSPACES: Dict[Tuple[int, int], torch.Tensor] = {}
DEFAULT_CHUNK_ROWS = 148 * 128     # the library's chunk (csrc/gdrf_capi.cu: DEFAULT_SMS * 128 observation rows)
_COPY_STREAMS: Dict[int, "torch.cuda.Stream"] = {}
_JITTER_HINTS: Dict[Tuple[int, int, int, int], int] = {}   # (device, M, D, kernel) -> jitter level of the last prologue


def _copy_stream(device: torch.device) -> "torch.cuda.Stream":
    """Host-to-device copy stream of the sub-shard pipeline, one per device."""
    key = device.index or 0
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(device)
    return _COPY_STREAMS[key]


_SIDE_STREAMS: Dict[int, "torch.cuda.Stream"] = {}


def _side_stream(device: torch.device) -> "torch.cuda.Stream":
    """Second compute stream (the speculative jitter probe runs on it next to the full prologue), one per device."""
    key = device.index or 0
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device)
    return _SIDE_STREAMS[key]


def _workspace(device: torch.device, nbytes: int) -> torch.Tensor:
    """One workspace per (device, stream): the C ABI is re-entrant per (workspace, stream) (include/gdrf_b200.h), so two
    modules driven on two streams of one device must not share scratch memory."""
    key = (device.index or 0, int(torch.cuda.current_stream(device).cuda_stream))
    ws = _WORKSPACES.get(key)
    if ws is None or ws.numel() < nbytes:
        _WORKSPACES.pop(key, None)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _WORKSPACES[key] = ws
    return ws


def release_workspaces() -> None:
    _WORKSPACES.clear()


def _f32(t: torch.Tensor, name: str, device) -> torch.Tensor:
    if t.device != device:
        raise ValueError(f"{name} is on {t.device}, expected {device}")
    return t.detach().to(torch.float32).contiguous()


def effective_jitter(jitter: float, njitter: int) -> float:
    """Diagonal loading after ``njitter`` failed attempts: the reference adds jitter*10^i cumulatively
    and in place (``gdrf/models/utils.py:33``)."""
    return float(sum(jitter * (10 ** i) for i in range(njitter + 1)))


def jitter_search(full, probe, speculate, maxjitter: int, hint: int, fp32_mode: bool, probe_max: int) -> int:
    """The level ``jittercholesky`` (gdrf/models/utils.py:27-40) stops at -- the first njitter < maxjitter whose
    factorisation succeeds -- found with as few launch chains and host read-backs as possible.  Pure host logic over three
    callables (``_Call.prologue`` binds them to the C ABI; tests/test_host.py drives them with tables):

    * ``full(nj)`` runs the full prologue at level nj and returns its status (0: factorised; > 0: failed);
    * ``probe(first, count)`` returns the fp32 pass / fail statuses of levels first .. first + count - 1 from ONE batched
      launch chain (count <= probe_max; only used when ``fp32_mode``: the reference's fp32 arithmetic decides);
    * ``speculate(hint)`` queues the probe of levels 0 .. hint - 1 next to the full prologue at ``hint`` and returns
      (status of the full prologue, statuses of the lower levels) from one read-back.

    ``hint`` is the level the previous evaluation of a model of this shape landed on (0: none).  Whatever the hint, the
    level returned is the reference's; RuntimeError with the reference's message when every level fails."""
    known = {}                               # level -> status already established (> 0: fails)
    if fp32_mode and 0 < hint < maxjitter and hint <= probe_max:
        st_full, lower = speculate(hint)
        known = {lvl: v for lvl, v in enumerate(lower)}
        if all(v > 0 for v in known.values()):
            if st_full == 0:
                return hint
            known[hint] = 1
    nj = 0
    while nj < maxjitter:
        if known.get(nj, 0) > 0:             # known to fail
            nj += 1
            continue
        if fp32_mode and nj > 0 and nj not in known:
            count = min(probe_max, maxjitter - nj)
            for i, v in enumerate(probe(nj, count)):
                known[nj + i] = v
            continue
        if full(nj) == 0:
            return nj
        known[nj] = 1
        nj += 1
    raise RuntimeError("reached max jitter, covariance is unstable")


class _Call:
    """Validated shapes + ctypes structs for one evaluation."""

    def __init__(self, xs, ws, Z, variance, lengthscale, u_loc, u_scale_tril, noise, phi, beta, eps,
                 kernel_id, n_offset, flags, chunk_rows, scale_mixture=None):
        if not xs.is_cuda:
            raise RuntimeError("gdrf_b200 runs on an sm_100a CUDA device only; there is no CPU path")
        dev = xs.device
        self.device = dev
        N, D = xs.shape
        M = Z.shape[0]
        K = u_loc.shape[0]
        V = ws.shape[1]
        if ws.shape[0] != N:
            raise ValueError(f"xs has {N} rows but ws has {ws.shape[0]}")
        if Z.shape[1] != D:
            raise ValueError("Inducing points and data should have the same number of dimensions, "
                             f"but got {Z.shape[1]} and {D}.")
        if u_loc.shape != (K, M) or (u_scale_tril is not None and u_scale_tril.shape != (K, M, M)):
            raise ValueError("u_loc must be [K, M] and u_scale_tril [K, M, M]")
        if phi.shape != (K, V) or beta.shape != (K, V):
            raise ValueError("phi and beta must be [K, V]")
        if eps.dim() not in (2, 3) or eps.shape[-2] != K or eps.shape[-1] < n_offset + N:
            raise ValueError("eps must be [K, >= n_offset + N] (or [particles, K, >= n_offset + N])")
        n_particles = int(eps.shape[0]) if eps.dim() == 3 else 1
        ls = lengthscale.reshape(-1)
        if ls.numel() not in (1, D):
            raise ValueError("lengthscale must have 1 or D entries")
        self.t = dict(
            xs=_f32(xs, "xs", dev), ws=ws.detach().to(torch.int32).contiguous(), eps=_f32(eps, "eps", dev),
            z=_f32(Z, "Z", dev), variance=_f32(variance.reshape(1), "variance", dev),
            lengthscale=_f32(ls, "lengthscale", dev), u_loc=_f32(u_loc, "u_loc", dev),
            u_scale_tril=None if u_scale_tril is None else _f32(u_scale_tril, "u_scale_tril", dev), noise=_f32(noise.reshape(1), "noise", dev),
            phi=_f32(phi, "phi", dev), beta=_f32(beta, "beta", dev),
            scale_mixture=None if scale_mixture is None else _f32(scale_mixture.reshape(1), "scale_mixture", dev))
        if int(kernel_id) == _lib.KERNEL_IDS["rationalquadratic"] and scale_mixture is None:
            raise ValueError("the RationalQuadratic kernel needs its scale_mixture parameter")
        if self.t["ws"].device != dev:
            raise ValueError("ws must live on the same device as xs")
        self.shape = _lib.Shape(n_local=N, n_offset=int(n_offset), n_eps=int(eps.shape[-1]), d=D, m=M, k=K, v=V,
                                kernel_id=int(kernel_id), ls_dim=int(ls.numel()), chunk_rows=int(chunk_rows),
                                flags=int(flags), n_particles=n_particles)
        self.inputs = _lib.Inputs(**{k: (None if v is None else v.data_ptr()) for k, v in self.t.items()})
        self.ws_bytes = _lib.workspace_bytes(self.shape)
        self.workspace = _workspace(dev, self.ws_bytes)
        self.stream = torch.cuda.current_stream(dev).cuda_stream

    def _full_prologue(self, jitter: float, nj: int, status: torch.Tensor) -> None:
        _lib.check(_lib.load().gdrf_prologue(ctypes.byref(self.shape), ctypes.byref(self.inputs), float(jitter), nj,
                                             self.workspace.data_ptr(), self.ws_bytes, self.stream, status.data_ptr()))

    def _probe(self, jitter: float, first: int, count: int, status: torch.Tensor, stream=None) -> None:
        """fp32 'does the reference's factorisation fail?' for levels first .. first+count-1 in one launch chain."""
        _lib.check(_lib.load().gdrf_jitter_probe(ctypes.byref(self.shape), ctypes.byref(self.inputs), float(jitter),
                                                 first, count, self.workspace.data_ptr(), self.ws_bytes,
                                                 self.stream if stream is None else stream, status.data_ptr()))

    def prologue(self, jitter: float, maxjitter: int) -> int:
        """jittercholesky (utils.py:27-40): the first level njitter < maxjitter whose factorisation succeeds.

        The reference walks the levels one by one, each a try/except around an fp32 ``torch.linalg.cholesky``.  Here a
        failing level 0 is followed by ONE batched probe of the remaining levels (``gdrf_jitter_probe``: the levels are
        independent and a factorisation is a latency-bound launch chain, so eight levels cost little more than one: 0.71 ms
        against 0.63 ms for a full prologue at M = 625, 1.33 against 0.96 ms at M = 1024) and the full prologue at the first
        level that passes.  A model that keeps landing on level L > 0 (C1 at the reference's
        defaults: L = 5 on every step) is met by speculation: the probe of levels 0 .. L-1 and the full prologue at L are
        queued back to back and read back together -- one host synchronisation per step; if the guess does not hold
        (a lower level passes now, or L fails) the general search runs.  The level returned is always the reference's."""
        maxjitter = int(maxjitter)
        status = torch.zeros(1 + _lib.PROBE_MAX, dtype=torch.int32, device=self.device)
        fp32_mode = bool(self.shape.flags & _lib.FLAG_CHOL_FP32_STATUS)
        key = (self.device.index or 0, self.shape.m, self.shape.d, self.shape.kernel_id)

        def full(nj):
            self._full_prologue(jitter, nj, status)
            st = int(status[0].item())       # host read-back (mirrors try/except)
            if st == -1:                     # factorised, but an operand may leave the fp16 range: repack as bf16 planes
                self.shape.flags |= _lib.FLAG_FWD_BF16
                self._full_prologue(jitter, nj, status)
                st = int(status[0].item())
            return st

        def probe(first, count):
            self._probe(jitter, first, count, status[1:])
            return status[1:1 + count].tolist()

        def speculate(hint):
            # the probe (its own scratch region of the workspace) runs on a side stream next to the full prologue: both
            # are latency-bound chains of small launches
            main = torch.cuda.current_stream(self.device)
            side = _side_stream(self.device)
            side.wait_stream(main)
            self._probe(jitter, 0, hint, status[1:], stream=side.cuda_stream)
            self._full_prologue(jitter, hint, status)
            main.wait_stream(side)
            st = status[:1 + hint].tolist()  # one read-back for the probe and the prologue
            if st[0] == -1 and all(v > 0 for v in st[1:]):
                st[0] = full(hint)           # the guess holds but the planes have to be repacked as bf16
            return st[0], st[1:]

        nj = jitter_search(full, probe, speculate, maxjitter, _JITTER_HINTS.get(key, 0), fp32_mode, _lib.PROBE_MAX)
        _JITTER_HINTS[key] = nj
        return nj

    def step(self, want_grad: bool, terms=None, grad=None, extra_flags: int = 0):
        lib = _lib.load()
        if terms is None:
            terms = torch.empty(4, dtype=torch.float64, device=self.device)
        flags = (self.shape.flags & ~(_lib.FLAG_WANT_GRAD | _lib.FLAG_TERMS_IN_GRAD)) | int(extra_flags)
        if want_grad:
            flags |= _lib.FLAG_WANT_GRAD
            ge = _lib.grad_elems(self.shape)
            if grad is None:      # 8 floats of tail: the terms ride behind the gradient (one all-reduce carries both)
                grad = torch.empty(ge + _lib.TERMS_TAIL, dtype=torch.float32, device=self.device)
            if grad.numel() >= ge + _lib.TERMS_TAIL:
                flags |= _lib.FLAG_TERMS_IN_GRAD
        self.shape.flags = flags
        out = _lib.Outputs(terms=terms.data_ptr(), grad=grad.data_ptr() if grad is not None else None)
        _lib.check(lib.gdrf_elbo_step(ctypes.byref(self.shape), ctypes.byref(self.inputs), ctypes.byref(out),
                                      self.workspace.data_ptr(), self.ws_bytes, self.stream))
        return terms, grad


def split_grad(flat: torch.Tensor, K: int, M: int, V: int, D: int, ls_dim: int):
    """Views into the flat gradient of include/gdrf_b200.h:gdrf_outputs (``scale_mixture`` is present when the
    buffer carries the RationalQuadratic kernel's extra entry; an 8-float tail holding the terms is ignored)."""
    o = 0
    out = {}
    base = K * M * M + K * M + K * V + M * D + 2 + ls_dim
    blocks = [("u_scale_tril", (K, M, M)), ("u_loc", (K, M)), ("phi", (K, V)), ("Z", (M, D)),
              ("variance", ()), ("lengthscale", (ls_dim,)), ("noise", ())]
    if flat.numel() in (base + 1, base + 1 + _lib.TERMS_TAIL):
        blocks.append(("scale_mixture", ()))
    for name, shape in blocks:
        n = 1
        for s in shape:
            n *= s
        out[name] = flat[o:o + n].view(shape)
        o += n
    return out


def flat_gradient(g) -> torch.Tensor:
    """The flat buffer behind the views returned by :func:`split_grad` / :func:`elbo_value_and_grads`, INCLUDING the
    8-float tail that carries the four ELBO terms as (hi, lo) fp32 pairs: one ``all_reduce`` of this tensor sums the
    parameter gradients and the loss of all ranks (one collective per step)."""
    base = g["u_scale_tril"]
    total = sum(v.numel() for v in g.values()) + _lib.TERMS_TAIL
    return base.reshape(-1).as_strided((total,), (1,))


def terms_from_flat(flat: torch.Tensor) -> torch.Tensor:
    """fp64 [lp_mu, lq, ll, lp_phi] from the tail of a (possibly all-reduced) flat gradient buffer."""
    t = flat[-_lib.TERMS_TAIL:].double().view(4, 2)
    return t[:, 0] + t[:, 1]


class GDRFElbo(torch.autograd.Function):
    """elbo_over_n = GDRFElbo.apply(xs, ws, Z, variance, lengthscale, u_loc, u_scale_tril, noise, phi, beta,
    eps, kernel_id, jitter, maxjitter, n_global, n_offset, include_prior, flags, chunk_rows, scale_mixture)"""

    last_terms: Optional[torch.Tensor] = None   # [lp_mu, lq, ll, lp_phi] of the most recent call (device, fp64)
    last_njitter: int = 0

    @staticmethod
    def forward(ctx, xs, ws, Z, variance, lengthscale, u_loc, u_scale_tril, noise, phi, beta, eps,
                kernel_id: int, jitter: float, maxjitter: int, n_global: int, n_offset: int = 0,
                include_prior: bool = True, flags: int = _lib.FLAG_CHOL_FP32_STATUS, chunk_rows: int = 0,
                scale_mixture=None):
        fl = int(flags) | (_lib.FLAG_INCLUDE_PRIOR if include_prior else 0)
        call = _Call(xs, ws, Z, variance, lengthscale, u_loc, u_scale_tril, noise, phi, beta, eps,
                     kernel_id, n_offset, fl, chunk_rows, scale_mixture)
        want_grad = any(ctx.needs_input_grad[i] for i in (2, 3, 4, 5, 6, 7, 8)) or \
            (scale_mixture is not None and ctx.needs_input_grad[19])
        ctx.sm_shape = None if scale_mixture is None else scale_mixture.shape
        GDRFElbo.last_njitter = call.prologue(jitter, maxjitter)
        terms, grad = call.step(want_grad)
        GDRFElbo.last_terms = terms
        ctx.n_global = int(n_global)
        ctx.dims = (u_loc.shape[0], Z.shape[0], ws.shape[1], xs.shape[1], call.shape.ls_dim)
        ctx.ls_shape = lengthscale.shape
        ctx.var_shape = variance.shape
        ctx.noise_shape = noise.shape
        ctx.dtypes = tuple(t.dtype for t in (Z, variance, lengthscale, u_loc, u_scale_tril, noise, phi))
        ctx.stream = call.stream
        if grad is not None:
            ctx.save_for_backward(grad)
        elbo = (terms[0] + terms[3] + terms[2] - terms[1]) / float(n_global)
        return elbo.to(torch.float32)

    @staticmethod
    def backward(ctx, grad_out):
        (flat,) = ctx.saved_tensors
        lib = _lib.load()
        flat = flat[:flat.numel() - _lib.TERMS_TAIL]
        dst = torch.empty_like(flat)
        go = grad_out.detach().to(torch.float32).contiguous()
        _lib.check(lib.gdrf_elbo_backward(flat.data_ptr(), flat.numel(), go.data_ptr(), 1.0 / ctx.n_global,
                                          dst.data_ptr(), torch.cuda.current_stream(flat.device).cuda_stream))
        K, M, V, D, ls_dim = ctx.dims
        g = split_grad(dst, K, M, V, D, ls_dim)
        dZ, dvar, dls, du, dS, dnoise, dphi = (g["Z"], g["variance"].reshape(ctx.var_shape),
                                               g["lengthscale"].reshape(ctx.ls_shape), g["u_loc"],
                                               g["u_scale_tril"], g["noise"].reshape(ctx.noise_shape), g["phi"])
        outs = [dZ, dvar, dls, du, dS, dnoise, dphi]
        outs = [o.to(dt) if ctx.needs_input_grad[i + 2] else None for i, (o, dt) in enumerate(zip(outs, ctx.dtypes))]
        dsm = None
        if ctx.sm_shape is not None and ctx.needs_input_grad[19]:
            dsm = g["scale_mixture"].reshape(ctx.sm_shape)
        return (None, None, *outs, None, None, None, None, None, None, None, None, None, None, dsm)


def elbo_value_and_grads(xs, ws, Z, variance, lengthscale, u_loc, u_scale_tril, noise, phi, beta, eps,
                         kernel: str = "rbf", jitter: float = 1e-8, maxjitter: int = 5, n_global=None,
                         n_offset: int = 0, include_prior: bool = True,
                         flags: int = _lib.FLAG_CHOL_FP32_STATUS, chunk_rows: int = 0, scale_mixture=None,
                         all_reduce: bool = False, group=None):
    """Direct (no autograd graph) evaluation: returns (terms fp64[4] = lp_mu, lq, ll, lp_phi;
    dict of d ELBO_sum / d constrained parameter; njitter).

    ``all_reduce=True`` (one process per GPU, torch.distributed initialised): terms and gradient are summed over the
    ranks of ``group``.  The step is split where ``GDRF_FLAG_PARTIAL`` splits it: as soon as the last chunk has added
    into d/d u_scale_tril (99.9 % of the bytes) its all-reduce starts on NCCL's stream, and the per-step epilogue
    (Cholesky adjoint, Kuu adjoint, prior, assembly of the small gradients and the terms) runs underneath it; a second,
    small all-reduce carries the rest and the loss."""
    import torch.distributed as dist
    fl = int(flags) | (_lib.FLAG_INCLUDE_PRIOR if include_prior else 0)
    call = _Call(xs, ws, Z, variance, lengthscale, u_loc, u_scale_tril, noise, phi, beta, eps,
                 _lib.KERNEL_IDS[kernel], n_offset, fl, chunk_rows, scale_mixture)
    nj = call.prologue(jitter, maxjitter)
    K, M = u_loc.shape
    reduce_ = all_reduce and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    if not reduce_:
        terms, grad = call.step(True)
    else:
        terms, grad = call.step(True, extra_flags=_lib.FLAG_PARTIAL)                 # chunks only: dS is complete
        big = dist.all_reduce(grad[:K * M * M], op=dist.ReduceOp.SUM, group=group, async_op=True)
        n_local = call.shape.n_local
        call.shape.n_local = 0                                                        # epilogue only
        call.shape.flags &= ~_lib.FLAG_PARTIAL
        call.step(True, terms=terms, grad=grad, extra_flags=_lib.FLAG_CONTINUE)
        call.shape.n_local = n_local
        dist.all_reduce(grad[K * M * M:], op=dist.ReduceOp.SUM, group=group)          # small gradients + the terms
        big.wait()
        terms = terms_from_flat(grad)
    g = split_grad(grad, K, Z.shape[0], ws.shape[1], xs.shape[1], call.shape.ls_dim)
    return terms, g, nj


def sub_shard_rows(N: int, n_sub: int) -> int:
    """Rows per sub-shard of :func:`elbo_value_and_grads_from_host` (the size its staging buffers need): N / n_sub
    rounded up to 256, and to whole chunks of 148 * 128 observations once that is at least half a chunk."""
    n_sub = max(1, min(int(n_sub), (N + 255) // 256))
    per = ((N + n_sub - 1) // n_sub + 255) // 256 * 256
    if 2 * per > DEFAULT_CHUNK_ROWS:
        per = (per + DEFAULT_CHUNK_ROWS - 1) // DEFAULT_CHUNK_ROWS * DEFAULT_CHUNK_ROWS
    return per


def elbo_value_and_grads_from_host(xs_host, ws_host, eps_host, Z, variance, lengthscale, u_loc, u_scale_tril, noise,
                                   phi, beta, kernel: str = "rbf", jitter: float = 1e-8, maxjitter: int = 5,
                                   n_global=None, n_offset: int = 0, include_prior: bool = True,
                                   flags: int = _lib.FLAG_CHOL_FP32_STATUS, n_sub: int = 8, staging=None,
                                   scale_mixture=None, all_reduce: bool = False, group=None):
    """Same result as :func:`elbo_value_and_grads`, with the observations (``xs_host`` [N, D] fp32, ``ws_host``
    [N, V] int32, ``eps_host`` [K, >= n_offset + N] fp32) living in pinned HOST memory.  The shard is cut into
    ``n_sub`` sub-shards; while sub-shard i is being evaluated on the compute stream, sub-shard i+1 is copied
    host-to-device on a second stream into the other half of a double buffer (GDRF_FLAG_CONTINUE /
    GDRF_FLAG_PARTIAL carry the accumulators across the calls).  Sub-shards of at least half a chunk are rounded up to
    whole chunks of 148 * 128 observations, so that only the last one ends in a short chunk; with one chunk per sub-shard
    the first copy (the only one nothing hides but the prologue) is 41 MB at the C4 shape.  Parameters are device tensors.
    ``all_reduce=True``: as in :func:`elbo_value_and_grads` (d/d u_scale_tril is reduced under the per-step epilogue).
    Returns (terms, grads dict, njitter)."""
    import torch.distributed as dist
    dev = Z.device
    N, D = xs_host.shape
    V = ws_host.shape[1]
    K, M = u_loc.shape
    fl = int(flags) | (_lib.FLAG_INCLUDE_PRIOR if include_prior else 0)
    per = sub_shard_rows(N, n_sub)
    bounds = [(lo, min(N, lo + per)) for lo in range(0, max(N, 1), per)]
    rows = per if N > 0 else 1
    if staging is not None and staging[0]["xs"].shape[0] < rows:
        raise ValueError(f"staging buffers hold {staging[0]['xs'].shape[0]} rows, the sub-shards have {rows}")
    if staging is None:
        staging = [dict(xs=torch.empty(rows, D, dtype=torch.float32, device=dev),
                        ws=torch.empty(rows, V, dtype=torch.int32, device=dev),
                        eps=torch.empty(K, rows, dtype=torch.float32, device=dev)) for _ in range(2)]
    compute = torch.cuda.current_stream(dev)
    copy_stream = _copy_stream(dev)
    def make_call(i):      # built while the previous sub-shard computes: the host stays ahead of the device
        lo, hi = bounds[i]
        st_ = staging[i % 2]
        return _Call(st_["xs"][:hi - lo], st_["ws"][:hi - lo], Z, variance, lengthscale, u_loc, u_scale_tril, noise, phi,
                     beta, st_["eps"], _lib.KERNEL_IDS[kernel], 0, fl, 0, scale_mixture)

    free_ev = [torch.cuda.Event(), torch.cuda.Event()]
    copied_ev = [torch.cuda.Event() for _ in bounds]

    def issue_copy(i):
        lo, hi = bounds[i]
        n = hi - lo
        st_ = staging[i % 2]
        with torch.cuda.stream(copy_stream):
            if i >= 2:
                copy_stream.wait_event(free_ev[i % 2])
            else:
                copy_stream.wait_stream(compute)
            st_["xs"][:n].copy_(xs_host[lo:hi], non_blocking=True)
            st_["ws"][:n].copy_(ws_host[lo:hi], non_blocking=True)
            for k in range(K):     # row by row: each source slice is contiguous pinned memory
                st_["eps"][k, :n].copy_(eps_host[k, n_offset + lo:n_offset + hi], non_blocking=True)
            copied_ev[i].record(copy_stream)

    issue_copy(0)
    call = make_call(0)
    nj = call.prologue(jitter, maxjitter)
    step_flags = call.shape.flags
    terms = torch.empty(4, dtype=torch.float64, device=dev)
    grad = torch.empty(_lib.grad_elems(call.shape) + _lib.TERMS_TAIL, dtype=torch.float32, device=dev)
    reduce_ = all_reduce and dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1
    for i in range(len(bounds)):
        if i + 1 < len(bounds):
            issue_copy(i + 1)
        compute.wait_event(copied_ev[i])
        extra = (_lib.FLAG_CONTINUE if i > 0 else 0) | (_lib.FLAG_PARTIAL if (i + 1 < len(bounds) or reduce_) else 0)
        call.shape.flags = step_flags & ~(_lib.FLAG_CONTINUE | _lib.FLAG_PARTIAL)
        call.step(True, terms=terms, grad=grad, extra_flags=extra)
        free_ev[i % 2].record(compute)
        last = call
        if i + 1 < len(bounds):
            call = make_call(i + 1)
    if reduce_:       # dS is complete: its all-reduce runs under the per-step epilogue (elbo_value_and_grads)
        big = dist.all_reduce(grad[:K * M * M], op=dist.ReduceOp.SUM, group=group, async_op=True)
        n_local = last.shape.n_local
        last.shape.n_local = 0
        last.shape.flags &= ~(_lib.FLAG_CONTINUE | _lib.FLAG_PARTIAL)
        last.step(True, terms=terms, grad=grad, extra_flags=_lib.FLAG_CONTINUE)
        last.shape.n_local = n_local
        dist.all_reduce(grad[K * M * M:], op=dist.ReduceOp.SUM, group=group)
        big.wait()
        terms = terms_from_flat(grad)
    g = split_grad(grad, K, M, V, D, last.shape.ls_dim)
    return terms, g, nj


def marginal_mean(xs, Z, variance, lengthscale, u_loc, kernel: str = "rbf", jitter: float = 1e-8,
                  maxjitter: int = 5, flags: int = _lib.FLAG_CHOL_FP32_STATUS, chunk_rows: int = 0,
                  scale_mixture=None) -> torch.Tensor:
    """f_loc [K, N] of ``log_topic_probs`` (sparse_gdrf.py:161-186)."""
    K, M = u_loc.shape
    N = xs.shape[0]
    dev = xs.device
    dummy_ws = torch.zeros(N, 1, dtype=torch.int32, device=dev)
    call = _Call(xs, dummy_ws, Z, variance, lengthscale, u_loc,
                 None, torch.ones((), device=dev), torch.ones(K, 1, device=dev),
                 torch.ones(K, 1, device=dev), torch.zeros(K, N, device=dev), _lib.KERNEL_IDS[kernel], 0, flags,
                 chunk_rows, scale_mixture)
    call.prologue(jitter, maxjitter)
    out = torch.empty(K, N, dtype=torch.float32, device=dev)
    _lib.check(_lib.load().gdrf_marginal_mean(ctypes.byref(call.shape), ctypes.byref(call.inputs), out.data_ptr(),
                                              call.workspace.data_ptr(), call.ws_bytes, call.stream))
    return out


def marginal_moments(xs, Z, variance, lengthscale, u_loc, u_scale_tril, kernel: str = "rbf", jitter: float = 1e-8,
                     maxjitter: int = 5, flags: int = _lib.FLAG_CHOL_FP32_STATUS, chunk_rows: int = 0,
                     scale_mixture=None, dtype=torch.float32):
    """(f_loc, f_var), each [K, N]: ``SparseGDRF.forward(Xnew, full_cov=False)`` (sparse_gdrf.py:277-319).
    ``dtype=torch.float64`` returns the moments as the library holds them (gdrf_marginal_moments_f64)."""
    K, M = u_loc.shape
    N = xs.shape[0]
    dev = xs.device
    dummy_ws = torch.zeros(N, 1, dtype=torch.int32, device=dev)
    call = _Call(xs, dummy_ws, Z, variance, lengthscale, u_loc, u_scale_tril, torch.ones((), device=dev),
                 torch.ones(K, 1, device=dev), torch.ones(K, 1, device=dev), torch.zeros(K, N, device=dev),
                 _lib.KERNEL_IDS[kernel], 0, flags, chunk_rows, scale_mixture)
    call.prologue(jitter, maxjitter)
    floc = torch.empty(K, N, dtype=dtype, device=dev)
    fvar = torch.empty(K, N, dtype=dtype, device=dev)
    fn = _lib.load().gdrf_marginal_moments_f64 if dtype == torch.float64 else _lib.load().gdrf_marginal_moments
    _lib.check(fn(ctypes.byref(call.shape), ctypes.byref(call.inputs), floc.data_ptr(), fvar.data_ptr(),
                  call.workspace.data_ptr(), call.ws_bytes, call.stream))
    return floc, fvar


class MarginalMoments(torch.autograd.Function):
    """(f_loc, f_var) = MarginalMoments.apply(xs, Z, variance, lengthscale, u_loc, u_scale_tril, kernel_id, jitter,
    maxjitter, flags, chunk_rows, scale_mixture, f64) -- the sparse-GP marginal of ``SparseGDRF.forward`` /
    ``gp.util.conditional(..., full_cov=False, whiten=True)`` (sparse_gdrf.py:277-319, :334-344), differentiable with
    respect to the constrained Z, variance, lengthscale, u_loc, u_scale_tril (and scale_mixture) the way torch autograd
    differentiates the reference's: the backward is ``gdrf_moments_vjp`` (the contractions of ``gdrf_elbo_step``'s
    backward, fed with the upstream gradients instead of the ELBO's per-observation weights)."""

    @staticmethod
    def forward(ctx, xs, Z, variance, lengthscale, u_loc, u_scale_tril, kernel_id: int, jitter: float, maxjitter: int,
                flags: int = _lib.FLAG_CHOL_FP32_STATUS, chunk_rows: int = 0, scale_mixture=None, f64: bool = False):
        K, M = u_loc.shape
        N = xs.shape[0]
        dev = xs.device
        call = _moments_call(xs, Z, variance, lengthscale, u_loc, u_scale_tril, kernel_id, flags, chunk_rows, scale_mixture)
        ctx.njitter = call.prologue(jitter, maxjitter)
        dtype = torch.float64 if f64 else torch.float32
        floc = torch.empty(K, N, dtype=dtype, device=dev)
        fvar = torch.empty(K, N, dtype=dtype, device=dev)
        fn = _lib.load().gdrf_marginal_moments_f64 if f64 else _lib.load().gdrf_marginal_moments
        _lib.check(fn(ctypes.byref(call.shape), ctypes.byref(call.inputs), floc.data_ptr(), fvar.data_ptr(),
                      call.workspace.data_ptr(), call.ws_bytes, call.stream))
        ctx.save_for_backward(xs, Z, variance, lengthscale, u_loc, u_scale_tril,
                              *( [scale_mixture] if scale_mixture is not None else []))
        ctx.has_sm = scale_mixture is not None
        ctx.args = (int(kernel_id), float(jitter), int(maxjitter), int(call.shape.flags), int(chunk_rows))
        return floc, fvar

    @staticmethod
    def backward(ctx, g_floc, g_fvar):
        saved = ctx.saved_tensors
        xs, Z, variance, lengthscale, u_loc, u_scale_tril = saved[:6]
        sm = saved[6] if ctx.has_sm else None
        kernel_id, jitter, maxjitter, flags, chunk_rows = ctx.args
        K, M = u_loc.shape
        call = _moments_call(xs, Z, variance, lengthscale, u_loc, u_scale_tril, kernel_id,
                             flags & ~_lib.FLAG_FWD_BF16, chunk_rows, sm)
        call.prologue(jitter, maxjitter)      # same parameters: lands on the level of the forward
        up_loc = torch.zeros(K, xs.shape[0], dtype=torch.float32, device=xs.device) if g_floc is None \
            else g_floc.detach().to(torch.float32).contiguous()
        up_var = None if g_fvar is None else g_fvar.detach().to(torch.float32).contiguous()
        grad = torch.empty(_lib.grad_elems(call.shape), dtype=torch.float32, device=xs.device)
        _lib.check(_lib.load().gdrf_moments_vjp(ctypes.byref(call.shape), ctypes.byref(call.inputs), up_loc.data_ptr(),
                                                None if up_var is None else up_var.data_ptr(), grad.data_ptr(),
                                                call.workspace.data_ptr(), call.ws_bytes, call.stream))
        g = split_grad(grad, K, M, 1, xs.shape[1], call.shape.ls_dim)
        need = ctx.needs_input_grad
        outs = [None,
                g["Z"].to(Z.dtype) if need[1] else None,
                g["variance"].reshape(variance.shape).to(variance.dtype) if need[2] else None,
                g["lengthscale"].reshape(lengthscale.shape).to(lengthscale.dtype) if need[3] else None,
                g["u_loc"].to(u_loc.dtype) if need[4] else None,
                g["u_scale_tril"].to(u_scale_tril.dtype) if need[5] else None,
                None, None, None, None, None,
                g["scale_mixture"].reshape(sm.shape).to(sm.dtype) if (sm is not None and need[11]) else None,
                None]
        return tuple(outs)


def _moments_call(xs, Z, variance, lengthscale, u_loc, u_scale_tril, kernel_id, flags, chunk_rows, scale_mixture):
    """A _Call for the moments-only entry points: the observation-side inputs they do not read are one-column dummies."""
    K = u_loc.shape[0]
    N = xs.shape[0]
    dev = xs.device
    return _Call(xs, torch.zeros(N, 1, dtype=torch.int32, device=dev), Z, variance, lengthscale, u_loc, u_scale_tril,
                 torch.ones((), device=dev), torch.ones(K, 1, device=dev), torch.ones(K, 1, device=dev),
                 torch.empty(K, N, device=dev),
                 kernel_id, 0, flags, chunk_rows, scale_mixture)


def marginal_moments_diff(xs, Z, variance, lengthscale, u_loc, u_scale_tril, kernel: str = "rbf", jitter: float = 1e-8,
                          maxjitter: int = 5, flags: int = _lib.FLAG_CHOL_FP32_STATUS, chunk_rows: int = 0,
                          scale_mixture=None, dtype=torch.float32):
    """:func:`marginal_moments` as a differentiable op (see :class:`MarginalMoments`)."""
    return MarginalMoments.apply(xs, Z, variance, lengthscale, u_loc, u_scale_tril, _lib.KERNEL_IDS[kernel], jitter,
                                 maxjitter, flags, chunk_rows, scale_mixture, dtype == torch.float64)


def perplexity_from_mean(floc: torch.Tensor, ws: torch.Tensor, phi: torch.Tensor) -> torch.Tensor:
    """exp(-sum w log(word_probs) / sum w)  (abstract_gdrf.py:137-139) without the N x V matrix."""
    K, N = floc.shape
    V = ws.shape[1]
    dev = floc.device
    shape = _lib.Shape(n_local=N, n_offset=0, n_eps=N, d=1, m=1, k=K, v=V, kernel_id=0, ls_dim=1, chunk_rows=0, flags=0)
    wsc = ws.detach().to(torch.int32).contiguous()
    phic = phi.detach().to(torch.float32).contiguous()
    inputs = _lib.Inputs(ws=wsc.data_ptr(), phi=phic.data_ptr())
    out = torch.zeros(2, dtype=torch.float64, device=dev)
    fl = floc.detach().to(torch.float32).contiguous()
    _lib.check(_lib.load().gdrf_perplexity_terms(ctypes.byref(shape), ctypes.byref(inputs), fl.data_ptr(),
                                                 out.data_ptr(), torch.cuda.current_stream(dev).cuda_stream))
    return torch.exp(-out[0] / out[1]).to(torch.float32)
