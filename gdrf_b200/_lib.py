"""ctypes binding of the C ABI in include/gdrf_b200.h.

The shared library is built in-tree (``gdrf_b200/libgdrf_b200.so``) by ``__graft_entry__.build()`` /
``gdrf_b200.build``.  There is no CPU or PyTorch fallback: if the library is missing, or the device is
not an sm_100a GPU, every entry point raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int32, c_int64, c_size_t, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libgdrf_b200.so")

KERNEL_IDS = {"rbf": 0, "matern32": 1, "matern52": 2, "exponential": 3, "rationalquadratic": 4}

FLAG_WANT_GRAD = 1
FLAG_INCLUDE_PRIOR = 2
FLAG_CHOL_FP32_STATUS = 4
FLAG_STATUS_ONLY = 8
FLAG_FWD_BF16 = 16
FLAG_SINGLE_CTA = 32
FLAG_CONTINUE = 64
FLAG_PARTIAL = 128
FLAG_REF_G = {i: 1 << (7 + i) for i in range(1, 7)}
FLAG_REF_ALL = 0x3F << 8
FLAG_FULL_WIDTH = 1 << 14
FLAG_INTERLEAVED_MMAS = 1 << 15
FLAG_SEGMENTED_FWD = 1 << 16
FLAG_TERMS_IN_GRAD = 1 << 17
FLAG_LIKELIHOOD_FMA = 1 << 18
TERMS_TAIL = 8


class Shape(ctypes.Structure):
    _fields_ = [("n_local", c_int64), ("n_offset", c_int64), ("n_eps", c_int64), ("d", c_int32),
                ("m", c_int32), ("k", c_int32), ("v", c_int32), ("kernel_id", c_int32),
                ("ls_dim", c_int32), ("chunk_rows", c_int32), ("flags", c_int32), ("n_particles", c_int32)]


class Inputs(ctypes.Structure):
    _fields_ = [(n, c_void_p) for n in ("xs", "ws", "eps", "z", "variance", "lengthscale", "u_loc",
                                        "u_scale_tril", "noise", "phi", "beta", "scale_mixture")]


class Outputs(ctypes.Structure):
    _fields_ = [("terms", c_void_p), ("grad", c_void_p)]


EXPORTS = ("gdrf_workspace_bytes", "gdrf_grad_elems", "gdrf_prologue", "gdrf_elbo_step",
           "gdrf_elbo_backward", "gdrf_marginal_mean", "gdrf_marginal_moments", "gdrf_perplexity_terms",
           "gdrf_marginal_moments_f64", "gdrf_jitter_probe", "gdrf_moments_vjp",
           "gdrf_constrain", "gdrf_adam_step", "gdrf_clipped_adam_step", "gdrf_gather_rows", "gdrf_last_error",
           "gdrf_build_info", "gdrf_launch_count", "gdrf_profile_enable", "gdrf_profile_read")

_lib = None


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: build it with `python -m gdrf_b200.build` "
                           "(gdrf_b200 has no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    P = ctypes.POINTER
    lib.gdrf_workspace_bytes.argtypes = [P(Shape), P(c_size_t)]
    lib.gdrf_grad_elems.argtypes = [P(Shape), P(c_int64)]
    lib.gdrf_prologue.argtypes = [P(Shape), P(Inputs), c_double, c_int, c_void_p, c_size_t, c_void_p, c_void_p]
    lib.gdrf_elbo_step.argtypes = [P(Shape), P(Inputs), P(Outputs), c_void_p, c_size_t, c_void_p]
    lib.gdrf_elbo_backward.argtypes = [c_void_p, c_int64, c_void_p, c_float, c_void_p, c_void_p]
    lib.gdrf_marginal_mean.argtypes = [P(Shape), P(Inputs), c_void_p, c_void_p, c_size_t, c_void_p]
    lib.gdrf_marginal_moments.argtypes = [P(Shape), P(Inputs), c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.gdrf_marginal_moments_f64.argtypes = [P(Shape), P(Inputs), c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    lib.gdrf_perplexity_terms.argtypes = [P(Shape), P(Inputs), c_void_p, c_void_p, c_void_p]
    lib.gdrf_constrain.argtypes = [P(Shape), c_void_p, c_void_p, c_int, c_void_p]
    lib.gdrf_adam_step.argtypes = [P(Shape), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                                   c_float, c_float, c_float, c_float, c_int, c_float, c_int, c_void_p]
    lib.gdrf_clipped_adam_step.argtypes = [P(Shape), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_float,
                                           c_float, c_float, c_float, c_float, c_float, c_int, c_float, c_int, c_void_p]
    lib.gdrf_gather_rows.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int64, ctypes.c_int32, ctypes.c_int32,
                                     c_void_p, c_void_p, c_void_p, c_void_p]
    lib.gdrf_jitter_probe.argtypes = [P(Shape), P(Inputs), c_double, c_int, c_int, c_void_p, c_size_t, c_void_p, c_void_p]
    lib.gdrf_moments_vjp.argtypes = [P(Shape), P(Inputs), c_void_p, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]
    for n in EXPORTS[:15]:
        getattr(lib, n).restype = c_int
    lib.gdrf_launch_count.restype = ctypes.c_longlong
    lib.gdrf_profile_enable.argtypes = [c_int]
    lib.gdrf_profile_enable.restype = c_int
    lib.gdrf_profile_read.argtypes = [P(c_double), P(ctypes.c_longlong)]
    lib.gdrf_profile_read.restype = c_int
    lib.gdrf_last_error.restype = c_char_p
    lib.gdrf_build_info.restype = c_char_p
    _lib = lib
    return lib


def check(code: int) -> None:
    if code != 0:
        raise RuntimeError(f"gdrf_b200: {load().gdrf_last_error().decode()} (code {code})")


def workspace_bytes(shape: Shape) -> int:
    out = c_size_t(0)
    check(load().gdrf_workspace_bytes(ctypes.byref(shape), ctypes.byref(out)))
    return int(out.value)


def grad_elems(shape: Shape) -> int:
    out = c_int64(0)
    check(load().gdrf_grad_elems(ctypes.byref(shape), ctypes.byref(out)))
    return int(out.value)


PROBE_MAX = 8      # GDRF_PROBE_MAX


PROFILE_KINDS = ("G1", "G2_fwd", "scale_w", "G3", "G4", "G5", "G6")


def profile_read():
    ms = (c_double * 7)()
    n = (ctypes.c_longlong * 7)()
    check(load().gdrf_profile_read(ms, n))
    return {k: (ms[i], int(n[i])) for i, k in enumerate(PROFILE_KINDS)}
