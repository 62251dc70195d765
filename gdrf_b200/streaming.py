"""Streaming / mini-batch inference around the fused ELBO op.

The reference's only large-N strategy (``gdrf/train_script.py:394-465``): every sub-epoch it builds a recency
distribution over the observations seen so far, draws a multiset of row indices from it with ``np.random.choice``, and
calls ``svi.step(xs=xs[selection, ...], ws=ws[selection, ...], subsample=False)`` under the *fixed* ``1 / len(xs)``
scale of ``train_script.py:365``.  Here the host logic is restated with the same names and the same use of the numpy
generator (so a seeded run draws the same rows), the data set stays resident in HBM and the fancy-index is one
HBM-bound gather kernel behind the C ABI (``gdrf_gather_rows``); the step is the same fused ELBO + gradient with
``n_global = len(xs)``.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence, Union

import numpy as np
import torch

from . import _lib
from .svi import shard_bounds

STREAMING_MODES = ("uniform", "now", "exp", "uniform_now", "exp_now", "uniform_exp")


def streaming_probabilities(streaming_inference: str, n_stream: int, streaming_exp: float = 1.0,
                            streaming_weight: float = 0.1) -> list:
    """train_script.py:404-441: weights over the ``n_stream`` observations seen so far (oldest first), normalised.
    Python floats and Python ``sum`` in the reference's order, so the probabilities handed to ``np.random.choice`` are
    bit-identical to the reference's."""
    if streaming_inference == "uniform":
        p = [1.0 for _ in range(n_stream)]
    elif streaming_inference == "now":
        p = [0.0 for _ in range(n_stream)]
        p[-1] = 1.0
    elif streaming_inference == "exp":
        p = [np.exp(-streaming_exp * (n_stream - i)) for i in range(n_stream)]
    elif streaming_inference == "uniform_now":
        p = [streaming_weight / n_stream for _ in range(n_stream)]
        p[-1] += 1 - streaming_weight
    elif streaming_inference == "exp_now":
        p = [np.exp(-streaming_exp * (n_stream - i)) for i in range(n_stream)]
        tot = sum(p)
        p = [streaming_weight * q / tot for q in p]
        p[-1] += 1 - streaming_weight
    elif streaming_inference == "uniform_exp":
        p = [np.exp(-streaming_exp * (n_stream - i)) for i in range(n_stream)]
        tot = sum(p)
        p = [(1 - streaming_weight) * q / tot for q in p]
        p = [q + streaming_weight / (n_stream + 1) for q in p]
    else:
        raise ValueError(
            "streaming_inference should be one of 'uniform', 'now', 'exp', 'uniform_now, 'exp_now', or 'uniform_exp'; "
            "you passed %s" % (streaming_inference,))
    tot = sum(p)
    return [q / tot for q in p]


def streaming_window(epoch: int, n_data: int, epochs: int, streaming_truncate: int = -1,
                     streaming_batch: bool = False) -> int:
    """train_script.py:396-402: how many observations are visible at this epoch."""
    effective_epoch = epoch * n_data // epochs if streaming_batch else epoch
    return min(effective_epoch + 1, streaming_truncate) if streaming_truncate > 0 else effective_epoch + 1


def streaming_selection(epoch: int, n_data: int, epochs: int, streaming_inference: str, streaming_size: int = 1,
                        streaming_truncate: int = -1, streaming_exp: float = 1.0, streaming_weight: float = 0.1,
                        streaming_batch: bool = False, rng=None) -> np.ndarray:
    """train_script.py:396-452: the row indices of one sub-epoch, as a 1-D int64 array (``streaming_size <= 1`` gives
    one row, the reference's ``unsqueeze(dim=0)``).  ``rng``: anything with numpy's ``choice`` (default: the global
    ``np.random`` the reference uses)."""
    rng = np.random if rng is None else rng
    n_stream = streaming_window(epoch, n_data, epochs, streaming_truncate, streaming_batch)
    p = streaming_probabilities(streaming_inference, n_stream, streaming_exp, streaming_weight)
    selection = rng.choice(n_stream, size=streaming_size if streaming_size > 1 else None, p=p)
    selection = np.atleast_1d(np.asarray(selection, dtype=np.int64))
    if streaming_truncate > 0:
        # the window slides with the epoch counter; the reference offsets by `epoch`, not the effective epoch
        # (:447-451), so with streaming_batch the indices can be negative and wrap from the end -- reproduced as is
        selection = selection + (epoch + 1 - n_stream)
    return selection


class StreamingData:
    """Device-resident data set + the gather kernel.  ``xs``: [N, D] float32 already scaled to the unit cube (as
    ``train_script.py:263-271`` hands it over), ``ws``: [N, V] int32."""

    def __init__(self, xs: torch.Tensor, ws: torch.Tensor, device: Union[str, torch.device] = "cuda"):
        dev = torch.device(device)
        if xs.dim() != 2 or ws.dim() != 2 or xs.shape[0] != ws.shape[0]:
            raise ValueError("xs must be [N, D] and ws [N, V] with the same N")
        self.xs = xs.to(dev, torch.float32).contiguous()
        self.ws = ws.to(dev, torch.int32).contiguous()
        self._status = None

    def __len__(self) -> int:
        return int(self.xs.shape[0])

    def gather(self, selection: Union[np.ndarray, Sequence[int], torch.Tensor]):
        """(xs[selection, ...], ws[selection, ...]) on the device.  Host index arrays are validated on the host (the
        reference's ``IndexError``); device index tensors are validated by the kernel's status word."""
        n = len(self)
        on_device = isinstance(selection, torch.Tensor) and selection.is_cuda
        if on_device:
            idx = selection.to(torch.int64).reshape(-1).contiguous()
        else:
            sel = np.atleast_1d(np.asarray(selection.cpu() if isinstance(selection, torch.Tensor) else selection,
                                           dtype=np.int64)).reshape(-1)
            bad = np.nonzero((sel < -n) | (sel >= n))[0]
            if bad.size:
                raise IndexError(f"index {int(sel[bad[0]])} is out of bounds for dimension 0 with size {n}")
            idx = torch.from_numpy(sel).to(self.xs.device, non_blocking=True)
        if self.xs.device.type != "cuda":
            raise RuntimeError("gdrf_b200 has no CPU path: StreamingData.gather needs a CUDA device")
        n_sel = int(idx.numel())
        xs_out = torch.empty(n_sel, self.xs.shape[1], dtype=torch.float32, device=self.xs.device)
        ws_out = torch.empty(n_sel, self.ws.shape[1], dtype=torch.int32, device=self.xs.device)
        status = None
        if on_device:
            status = torch.zeros(1, dtype=torch.int32, device=self.xs.device)
        st = torch.cuda.current_stream(self.xs.device).cuda_stream
        _lib.check(_lib.load().gdrf_gather_rows(
            self.xs.data_ptr(), self.ws.data_ptr(), idx.data_ptr(), n_sel, n, int(self.xs.shape[1]),
            int(self.ws.shape[1]), xs_out.data_ptr(), ws_out.data_ptr(),
            status.data_ptr() if status is not None else None, ctypes.c_void_p(st)))
        if status is not None:
            s = int(status.item())
            if s:
                raise IndexError(f"selection[{s - 1}] is out of bounds for dimension 0 with size {n}")
        return xs_out, ws_out


def streaming_epoch(svi, data: StreamingData, epoch: int, epochs: int, streaming_inference: str,
                    streaming_size: int = 1, streaming_subepochs: int = 1, streaming_truncate: int = -1,
                    streaming_exp: float = 1.0, streaming_weight: float = 0.1, streaming_batch: bool = False,
                    rng=None, eps_fn=None, shard=None) -> float:
    """One epoch of the reference's streaming loop (train_script.py:394-460): ``streaming_subepochs`` steps, each on a
    freshly drawn multiset of rows, every loss scaled by the full data set's ``1 / len(xs)``.  ``shard=(rank, world)``
    splits each drawn multiset across the ranks of a data-parallel job (``svi`` all-reduces the gradients).  ``svi``: a
    :class:`gdrf_b200.svi.SVI` / ``FusedSVI``.  ``eps_fn(n_rows)`` may supply the guide's draws (tests).  Returns the
    last loss like the reference's loop variable."""
    n_data = len(data)
    loss = float("nan")
    for _ in range(streaming_subepochs):
        selection = streaming_selection(epoch, n_data, epochs, streaming_inference, streaming_size, streaming_truncate,
                                        streaming_exp, streaming_weight, streaming_batch, rng)
        if shard is not None:     # observation-sharded data parallelism: every rank draws the SAME multiset (same seeded
            lo, hi = shard_bounds(len(selection), *shard)     # generator) and evaluates its contiguous slice of it
            selection = selection[lo:hi]
        xs_stream, ws_stream = data.gather(selection)
        eps = eps_fn(xs_stream.shape[0]) if eps_fn is not None else None
        loss = svi.step(xs=xs_stream, ws=ws_stream, subsample=False, eps=eps, n_global=n_data)
    return loss
