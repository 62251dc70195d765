"""bench.py -- ELBO + gradient throughput of the sparse multinomial GDRF (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config C4|C3|...] [--obs N]

One "step" = one evaluation of the ELBO and its full gradient over the whole synthetic data set
(SURVEY.md 8(d) inputs at BASELINE.json configs[3]: N=1M, D=3, K=32, V=512, M=1024=16x8x8, RBF), including
the M x M prologue and, for N > 1 ranks, the all-reduce of the parameter gradients.  Observations are
sharded by index across ranks (strong scaling: the data set is fixed, no data-path collective).
Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CONFIGS = {
    # name: N, D, K, V, grid, kernel      
    "C1": dict(N=4800, D=2, K=5, V=100, grid=[25, 25], kernel="rbf"),
    "C2": dict(N=100_000, D=1, K=8, V=174, grid=[1000], kernel="matern32"),
    "C3": dict(N=100_000, D=2, K=16, V=128, grid=[16, 16], kernel="rbf"),
    "C4": dict(N=1_000_000, D=3, K=32, V=512, grid=[16, 8, 8], kernel="rbf"),
    "C5": dict(N=8_000_000, D=3, K=64, V=1024, grid=[16, 16, 8], kernel="matern52"),
}
METRIC = "ELBO+grad observations/sec (N=1M,K=32,V=512,M=1024) at 1/2/4/8 B200 vs CPU"
GEN_CHUNK = 15_625          # data generated in seeded chunks so that any sharding sees identical observations


def algorithmic_flops_per_obs(M, K, V):
    """each logical fp32 multiply-add counted once, S_k triangular, no credit for
    split-precision emulation or recompute."""
    return 3 * M * M * K + 3 * M * M + 6 * M * K + 6 * K * V


def params_for(cfg, device, dtype=None):
    """Deterministic parameters of SURVEY.md 8(d) (variance 25, lengthscale 0.75 x grid spacing, noise 1,
    u_loc ~ 0.5 N(0,1), S_k = chol(Kuu + jI) + 0.05 tril(N(0,1)), phi = row softmax(N(0,1)), beta 0.01)."""
    import torch
    from gdrf_b200.kernels import KERNEL_DICT
    grid, D, K, V = cfg["grid"], cfg["D"], cfg["K"], cfg["V"]
    axes = [torch.linspace(0.0, 1.0, n) if n > 1 else torch.tensor([0.5]) for n in grid]
    Z = torch.stack([m.flatten() for m in torch.meshgrid(*axes, indexing="ij")]).T.contiguous().float()
    M = Z.shape[0]
    spacing = min(1.0 / (n - 1) for n in grid if n > 1)
    ls = torch.tensor([0.75 * spacing])
    var = torch.tensor(25.0)
    jitter = 1e-4
    kern = KERNEL_DICT[cfg["kernel"]](D, variance=var.double(), lengthscale=ls.double()).double()
    with torch.no_grad():
        Kuu = kern(Z.double())
    L0 = torch.linalg.cholesky(Kuu + jitter * torch.eye(M, dtype=torch.float64)).float()
    S = L0.expand(K, M, M) + 0.05 * torch.randn(K, M, M, generator=torch.Generator().manual_seed(8)).tril()
    S = S.tril().contiguous()
    dg = S.diagonal(dim1=-2, dim2=-1)
    dg.copy_(dg.abs().clamp(min=1e-3))
    p = dict(Z=Z, variance=var, lengthscale=ls, noise=torch.tensor(1.0),
             u_loc=0.5 * torch.randn(K, M, generator=torch.Generator().manual_seed(7)),
             u_scale_tril=S,
             phi=torch.softmax(torch.randn(K, V, generator=torch.Generator().manual_seed(9)), -1),
             beta=torch.full((K, V), 0.01))
    return {k: v.to(device) for k, v in p.items()}, jitter, 15


def gen_chunk(cfg, chunk_id, rows, device):
    """xs ~ U[0,1)^D, counts ~ Multinomial(n_n, theta* phi*), n_n ~ U{V..10V-1}, eps ~ N(0,1); seeded by the
    global chunk index."""
    import torch
    D, K, V = cfg["D"], cfg["K"], cfg["V"]
    g = torch.Generator(device=device).manual_seed(1234 + 7919 * chunk_id)
    xs = torch.rand(rows, D, generator=g, device=device)
    gp = torch.Generator(device=device).manual_seed(4321)
    phi_star = torch._sample_dirichlet(torch.full((K, V), 0.1, device=device), generator=gp)
    theta_star = torch._sample_dirichlet(torch.full((rows, K), 0.3, device=device), generator=g)
    probs = theta_star @ phi_star
    counts = torch.randint(V, 10 * V, (rows,), generator=g, device=device)
    ws = torch.zeros(rows, V, dtype=torch.int32, device=device)
    sub = 4096
    for r0 in range(0, rows, sub):
        r1 = min(rows, r0 + sub)
        idx = torch.multinomial(probs[r0:r1], 10 * V, replacement=True, generator=g)
        mask = (torch.arange(10 * V, device=device)[None, :] < counts[r0:r1, None]).to(torch.int32)
        ws[r0:r1].scatter_add_(1, idx, mask)
    eps = torch.randn(K, rows, generator=g, device=device)
    return xs, ws, eps


def gen_shard(cfg, lo, hi, device):
    import torch
    assert lo % GEN_CHUNK == 0 or cfg["N"] < GEN_CHUNK
    xs_l, ws_l, eps_l = [], [], []
    n = lo
    while n < hi:
        rows = min(GEN_CHUNK, hi - n)
        x, w, e = gen_chunk(cfg, n // GEN_CHUNK, rows, device)
        xs_l.append(x); ws_l.append(w); eps_l.append(e)
        n += rows
    return torch.cat(xs_l), torch.cat(ws_l), torch.cat(eps_l, dim=1)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


PROFILE_SUMMARY = os.path.join(ROOT, "profiles", "r02b_gemm_ncu_summary.json")
PROFILE_FALLBACK = os.path.join(ROOT, "profiles", "r02_gemm_ncu_summary.json")


def roofline_block(prof, steps, n_local, ms_step_rank0, M, K, flags, _lib):
    """`roofline` of the JSON line from the CUDA-event records of the timed region (gdrf_profile_read: one record per
    launch, none dropped): algorithmic flops of the launches actually recorded / their summed duration."""
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peaks = json.load(open(peaks_path))
        peak_tf, peak_src = float(peaks.get("bf16_tflops_sustained", 1384.6)), "MEASURED_PEAKS.json bf16_tflops_sustained"
        burst_tf = float(peaks.get("bf16_tflops", peak_tf))
    else:
        peak_tf, peak_src = 1400.0, "fallback (B200_PROFILING.md sustained)"
        burst_tf = peak_tf
    # plausibility bound on issued MMA flops: 1.3 x the measured peak, taking the BURST figure -- a short kernel on a cool
    # box (small configurations, the 8-GPU shards) legitimately runs above the power-capped sustained number
    issued_limit = 1.3 * max(burst_tf, peak_tf) / peak_tf
    tri = 2.0 * (M * M / 2.0) * K                 # triangular-aware flops per observation of one W*S_k contraction
    alg = {"G1": 2.0 * M * M / 2, "G2_fwd": tri, "scale_w": 0.0, "G3": tri, "G4": 2.0 * M * M / 2, "G5": 2.0 * M * M,
           "G6": tri}
    Mp = (M + 255) // 256 * 256
    MBk = Mp // 64
    full_width = bool(flags & _lib.FLAG_FULL_WIDTH)
    # MMA work actually issued per logical multiply-add: products per split x whole 64-row k-blocks of 256-wide tiles of
    # the triangular operand (minus the column groups the narrow diagonal-block MMAs leave out) x padding of M to 256
    def issued_factor(kind):
        mma = {"G1": 6, "G2_fwd": 6 if (flags & _lib.FLAG_FWD_BF16) else 3}.get(kind, 3)
        if kind in ("G2_fwd", "G3", "G6"):
            narrow = 0 if full_width else {"G2_fwd": 6, "G3": 4}.get(kind, 0)
            pad = sum((MBk - 4 * j) * 4 - narrow for j in range(MBk // 4)) / (MBk * MBk / 2.0)
        elif kind in ("G1", "G4"):
            pad = sum((MBk - 4 * j) * 4 for j in range(MBk // 4)) / (MBk * MBk / 2.0)
        else:
            pad = 1.0
        return mma, pad * (float(Mp) / M) ** 2
    try:
        summ = json.load(open(PROFILE_SUMMARY if os.path.exists(PROFILE_SUMMARY) else PROFILE_FALLBACK))
    except Exception:
        summ = {}
    tags = {"G2_fwd": "G2<2>", "G3": "<G3>", "G6": "<G6>", "G1": "<G1T", "G4": "<G4T", "G5": "<G5T", "scale_w": "k_scale_w"}
    per_kernel = {}
    for kind, (ms, n) in prof.items():
        if n <= 0 or ms <= 0:
            continue
        entry = {"ms_per_step": ms / steps, "launches_per_step": n / steps, "avg_launch_ms": ms / n}
        if alg.get(kind, 0) > 0:
            # every observation of the shard passes through each contraction once per step
            tf = alg[kind] * n_local * steps / (ms * 1e-3) / 1e12
            mma, pad = issued_factor(kind)
            entry.update({"achieved_tflops": tf, "frac": tf / peak_tf, "issued_frac": tf * mma * pad / peak_tf,
                          "mma_per_product": mma, "tile_padding": pad})
        rec = next((v for k, v in summ.items() if tags.get(kind, "?") in k), None)
        if rec:
            entry["traffic"] = rec.get("dram_bytes")
            if rec.get("distinct_bytes"):
                entry["traffic_over_distinct"] = rec["dram_bytes"] / rec["distinct_bytes"]
            if rec.get("sm_mhz"):
                entry["ncu_sm_mhz"] = rec["sm_mhz"]
        per_kernel[kind] = entry
    doms = [k for k in per_kernel if "frac" in per_kernel[k]]
    if not doms:
        return None
    dom = max(doms, key=lambda k: per_kernel[k]["ms_per_step"])
    d = per_kernel[dom]
    total_ms = sum(e["ms_per_step"] for e in per_kernel.values())
    problems = []
    if total_ms > 1.02 * ms_step_rank0:
        problems.append(f"profiled kernels sum to {total_ms:.1f} ms/step > step {ms_step_rank0:.1f} ms")
    if any(e.get("issued_frac", 0) > issued_limit for e in per_kernel.values()):
        problems.append(f"issued_frac > {issued_limit:.2f} (1.3 x the measured burst bf16 peak)")
    if any(abs(e["launches_per_step"] - round(e["launches_per_step"])) > 1e-6 for e in per_kernel.values()):
        problems.append("launch records are not a whole number per step")
    if problems:      # a roofline that does not follow from the records is not printed
        return {"error": "; ".join(problems), "all_contractions_ms_per_step": {k: e["ms_per_step"] for k, e in per_kernel.items()}}
    return {"bound": "tensor", "kernel": f"gemm_tc2_kernel<{dom}>", "achieved": d["achieved_tflops"], "peak": peak_tf,
            "unit": "TFLOP/s", "frac": d["frac"], "traffic": d.get("traffic"),
            "issued": {"tflops": d["issued_frac"] * peak_tf, "frac": d["issued_frac"], "mma_per_product": d["mma_per_product"],
                       "tile_padding": d["tile_padding"],
                       "note": "issued 16-bit MMA flops / peak: the tensor-pipe utilisation this kernel runs at"},
            "traffic_note": "dram__bytes_read+write per launch, ncu --set full, one 18 944-observation chunk (profiles/)",
            "peak_source": peak_src, "avg_launch_ms": d["avg_launch_ms"], "launches": int(round(d["launches_per_step"] * steps)),
            "share_of_step": d["ms_per_step"] / ms_step_rank0,
            "note": "algorithmic flops: triangular-aware, each fp32 multiply-add counted once; the kernel issues 3 "
                    "16-bit MMAs per logical product (error-compensated split)",
            "all_contractions_ms_per_step": {k: e["ms_per_step"] for k, e in per_kernel.items()},
            "per_kernel": per_kernel}


def cpu_reference_rate(cfg, sample_rows, steps, warmup):
    """The reference's CPU path: the fp32 oracle executed op for op as the reference does (two conditionals,
    unfused ops, N x V probabilities materialised, torch autograd backward), all host threads, on a bounded
    sample of the workload (observations are independent given the parameters, so obs/s does not depend on N
    beyond the M x M prologue, which is included)."""
    import torch
    from oracle import gdrf_oracle as O     # the one place bench.py executes oracle/: the reported baseline
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    dev = torch.device("cpu")
    p, jitter, maxjitter = params_for(cfg, dev)
    xs, ws, eps = gen_chunk(cfg, 0, sample_rows, dev)
    inp = O.OracleInputs(xs=xs, ws=ws, Z=p["Z"], variance=p["variance"], lengthscale=p["lengthscale"],
                         u_loc=p["u_loc"], u_scale_tril=p["u_scale_tril"], noise=p["noise"], phi=p["phi"],
                         beta=p["beta"], eps=eps, kernel=cfg["kernel"], jitter=jitter, maxjitter=maxjitter,
                         n_global=cfg["N"])
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.loss_and_grads(inp, twice=True)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    best = min(times)
    return {"value": sample_rows / best, "unit": "observations/s", "cores": cores, "kind": "port",
            "sample": f"{sample_rows} observations of the workload, fp32 oracle as the reference executes it "
                      f"(2 conditionals + autograd), best of {steps} after {warmup} warm-up",
            "ms_per_step": 1e3 * statistics.mean(times)}


def run_reference(args, cfg_name, cfg):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rows = 2000 if cfg["K"] * cfg["grid"][0] > 200 else 8000
    rows = min(rows, cfg["N"])
    steps, warmup = max(1, min(args.steps, 3)), max(1, min(args.warmup, 1))
    r = cpu_reference_rate(cfg, rows, steps, warmup)
    line = {"metric": METRIC, "value": r["value"], "unit": "observations/s", "n_gpus": args.gpus, "steps": steps,
            "warmup": warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "impl": "reference",
            "config": {"workload": f"{cfg_name}: N={cfg['N']} D={cfg['D']} K={cfg['K']} V={cfg['V']} "
                                   f"M={_prod(cfg['grid'])} {cfg['kernel']}; bounded sample of {rows} observations/step"},
            "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
            "e2e": {"value": r["value"], "unit": "observations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def _prod(xs):
    p = 1
    for x in xs:
        p *= x
    return p


class Bench:
    """One configuration on this rank's shard: data, parameters, and timed passes through the PUBLIC surface of the
    package (gdrf_b200.elbo.elbo_value_and_grads / elbo_value_and_grads_from_host, gdrf_b200.SparseMultinomialGDRF +
    FusedSVI); the only private import is the instrumentation (launch counter, CUDA-event records)."""

    def __init__(self, name, cfg, dev, rank, world, flags_extra=0):
        import torch
        from gdrf_b200 import _lib
        from gdrf_b200.svi import shard_bounds
        self.torch, self._lib = torch, _lib
        self.name, self.cfg, self.dev, self.rank, self.world = name, cfg, dev, rank, world
        N = cfg["N"]
        lo, hi = shard_bounds(N, rank, world)
        if N >= GEN_CHUNK * world:      # keep shards on generation-chunk boundaries: any sharding sees the same data
            per = (N // GEN_CHUNK) // world * GEN_CHUNK
            lo, hi = rank * per, (N if rank == world - 1 else (rank + 1) * per)
        self.lo, self.hi, self.n_local = lo, hi, hi - lo
        self.xs, self.ws, self.eps = gen_shard(cfg, lo, hi, dev)
        self.prm, self.jitter, self.maxjitter = params_for(cfg, dev)
        self.M = _prod(cfg["grid"])
        self.flags = _lib.FLAG_CHOL_FP32_STATUS | int(flags_extra)
        self.chunk_rows = int(os.environ.get("GDRF_BENCH_CHUNK_ROWS", "0"))

    # ---- one ELBO + gradient evaluation over the shard, then the step's one collective ----
    def step(self, eps=None):
        import torch.distributed as dist
        from gdrf_b200.elbo import elbo_value_and_grads, flat_gradient, terms_from_flat
        p = self.prm
        terms, g, _ = elbo_value_and_grads(self.xs, self.ws, p["Z"], p["variance"], p["lengthscale"], p["u_loc"],
                                           p["u_scale_tril"], p["noise"], p["phi"], p["beta"],
                                           self.eps if eps is None else eps, kernel=self.cfg["kernel"],
                                           jitter=self.jitter, maxjitter=self.maxjitter, n_global=self.cfg["N"],
                                           include_prior=(self.rank == 0), flags=self.flags, chunk_rows=self.chunk_rows,
                                           all_reduce=True)      # gradient + terms summed over the ranks (NCCL)
        return terms, g

    def barrier(self):
        import torch.distributed as dist
        self.torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, steps, warmup):
        """W untimed + K timed calls of fn, CUDA events on the current stream, barrier + synchronize on both sides,
        max over ranks.  Returns (ms per step, this rank's own ms per step, last result)."""
        import torch.distributed as dist
        torch = self.torch
        out = None
        for _ in range(warmup):
            out = fn()
        self.barrier()
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        for a, b in ev:
            a.record()
            out = fn()
            b.record()
        self.barrier()
        mine = sum(a.elapsed_time(b) for a, b in ev)
        t = torch.tensor([mine], device=self.dev, dtype=torch.float64)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item() / steps, mine / steps, out

    def profiled(self, steps, warmup):
        """`timed(self.step)` with the library's per-launch CUDA-event records and launch counter switched on."""
        lib, _lib = self._lib.load(), self._lib
        for _ in range(warmup):
            self.step()
        self.barrier()
        lib.gdrf_profile_enable(1)
        _lib.profile_read()
        l0 = lib.gdrf_launch_count()
        ms, mine, out = self.timed(self.step, steps, 0)
        launches = lib.gdrf_launch_count() - l0
        prof = _lib.profile_read()
        lib.gdrf_profile_enable(0)
        return ms, mine, out, prof, int(launches)

    def e2e(self, steps):
        """Same metric through the host-buffer entry point: observations in pinned HOST memory, sub-shards copied
        H2D on a second stream while the previous one computes, the four terms read back to the host every step.  The
        134 MB gradient stays on the device (it feeds an on-device optimiser: FusedSVI)."""
        import torch.distributed as dist
        from gdrf_b200.elbo import DEFAULT_CHUNK_ROWS, elbo_value_and_grads_from_host, sub_shard_rows
        torch, p, cfg = self.torch, self.prm, self.cfg
        hx, hw, he = (t_.cpu().pin_memory() for t_ in (self.xs, self.ws, self.eps))
        h_terms = torch.empty(4, dtype=torch.float64).pin_memory()
        # one chunk per sub-shard: the first copy (41 MB at C4), the only one not hidden by compute, fits under the prologue
        n_sub = max(2, -(-self.n_local // DEFAULT_CHUNK_ROWS))
        per = sub_shard_rows(self.n_local, n_sub)
        D, K, V = cfg["D"], cfg["K"], cfg["V"]
        staging = [dict(xs=torch.empty(per, D, dtype=torch.float32, device=self.dev),
                        ws=torch.empty(per, V, dtype=torch.int32, device=self.dev),
                        eps=torch.empty(K, per, dtype=torch.float32, device=self.dev)) for _ in range(2)]

        def one():
            tm, g_, _ = elbo_value_and_grads_from_host(
                hx, hw, he, p["Z"], p["variance"], p["lengthscale"], p["u_loc"], p["u_scale_tril"], p["noise"],
                p["phi"], p["beta"], kernel=cfg["kernel"], jitter=self.jitter, maxjitter=self.maxjitter,
                n_global=cfg["N"], include_prior=(self.rank == 0), flags=self.flags, n_sub=n_sub, staging=staging,
                all_reduce=True)         # gradient + terms summed over the ranks (NCCL), dS under the per-step epilogue
            h_terms.copy_(tm, non_blocking=True)
            torch.cuda.current_stream().synchronize()
            return h_terms

        ms, _, ht = self.timed(one, steps, 1)
        h2d = hx.numel() * 4 + hw.numel() * 4 + he.numel() * 4
        return {"value": cfg["N"] / (ms * 1e-3), "unit": "observations/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": 32, "ms_per_step": ms,
                "loss": -float(ht[0] + ht[3] + ht[2] - ht[1]) / cfg["N"],      # the terms read back on the last step
                "note": "host buffers -> C ABI; the four ELBO terms come back to the host, the gradient stays on the "
                        "device for the on-device optimiser"}

    def svi_step(self, steps, warmup):
        """What the reference's caller drives every epoch (train_script.py:365-371,467): module -> constraint transforms
        -> ELBO + gradient -> all-reduce -> optimiser, through SparseMultinomialGDRF + FusedSVI only."""
        import gdrf_b200
        from gdrf_b200.kernels import KERNEL_DICT
        torch, cfg, p = self.torch, self.cfg, self.prm
        kern = KERNEL_DICT[cfg["kernel"]](cfg["D"], variance=p["variance"].cpu(), lengthscale=p["lengthscale"].cpu())
        m = gdrf_b200.SparseMultinomialGDRF(
            num_observation_categories=cfg["V"], num_topic_categories=cfg["K"], world=[(0.0, 1.0)] * cfg["D"],
            kernel=kern, dirichlet_param=0.01, n_points=list(cfg["grid"]), inducing_init="grid", device=str(self.dev),
            jitter=self.jitter, maxjitter=self.maxjitter, fixed_inducing_points=False)
        with torch.no_grad():
            m.u_loc_unconstrained.copy_(p["u_loc"])
            m._word_topic_matrix_map_unconstrained.copy_(p["phi"].log())
        m.seed_eps(2024)
        svi = gdrf_b200.FusedSVI(m, lr=1e-3)
        ms, _, loss = self.timed(lambda: svi.step(self.xs, self.ws, n_global=cfg["N"]), steps, warmup)
        return {"value": cfg["N"] / (ms * 1e-3), "unit": "observations/s", "ms_per_step": ms, "steps": steps,
                "loss_after": loss, "path": "SparseMultinomialGDRF + FusedSVI.step (constrain -> ELBO+grad -> "
                                            "all-reduce -> Adam), eps drawn on the device"}


def other_config_record(name, cfg, dev, peak_info):
    """One of the other BASELINE configurations on one GPU: value, dominant kernel and its roofline fraction."""
    b = Bench(name, cfg, dev, 0, 1)
    ms, mine, (terms, _), prof, launches = b.profiled(3, 3)
    roof = roofline_block(prof, 3, b.n_local, mine, b.M, cfg["K"], b.flags, b._lib) or {}
    t = terms.cpu()
    rec = {"workload": f"{name}: N={cfg['N']} D={cfg['D']} K={cfg['K']} V={cfg['V']} M={b.M} {cfg['kernel']}",
           "value": cfg["N"] / (ms * 1e-3), "unit": "observations/s", "ms_per_step": ms,
           "loss": -float(t[0] + t[3] + t[2] - t[1]) / cfg["N"], "gpu_launches_per_step": launches // 3,
           "dominant_kernel": roof.get("kernel"), "roofline_frac": roof.get("frac"),
           "issued_frac": (roof.get("issued") or {}).get("frac"),
           "alg_tflops_step": algorithmic_flops_per_obs(b.M, cfg["K"], cfg["V"]) * cfg["N"] / (ms * 1e-3) / 1e12}
    if "error" in roof:
        rec["roofline_error"] = roof["error"]
    del b
    return rec


def c1_record(dev):
    """BASELINE configs[0]: the shipped data/data_2d_artificial.csv (tests/golden/c1_artificial2d.npz holds it as
    train_script.py:251-273 reads it) at train()'s defaults: K=5, 25 x 25 inducing grid, RBF, jitter 1e-8 with the
    reference's escalation (lands on level 5)."""
    import numpy as np
    import torch
    from gdrf_b200.elbo import elbo_value_and_grads
    from gdrf_b200.models import host_jittercholesky
    from gdrf_b200.kernels import KERNEL_DICT
    d = np.load(os.path.join(ROOT, "tests", "golden", "c1_artificial2d.npz"))
    t = lambda k: torch.from_numpy(np.asarray(d[k])).to(dev)
    Z, var, ls = t("Z").float(), t("variance").float(), t("lengthscale").float()
    K, M = d["u_loc"].shape
    jitter, maxjitter = float(d["jitter"]), int(d["maxjitter"])
    kern = KERNEL_DICT["rbf"](2, variance=var.cpu(), lengthscale=ls.cpu())
    with torch.no_grad():
        L = host_jittercholesky(kern(Z.cpu()).contiguous(), M, jitter, maxjitter)      # constructor init, sparse_gdrf.py:100-110
    S = L.float().expand(K, M, M).contiguous().to(dev)
    args = (t("xs").float(), t("ws").int(), Z, var, ls, t("u_loc").float(), S, t("noise").float(), t("phi").float(),
            t("beta").float(), t("eps").float())
    N = args[0].shape[0]

    def one():
        return elbo_value_and_grads(*args, kernel="rbf", jitter=jitter, maxjitter=maxjitter)
    for _ in range(3):
        one()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
    for a, b in ev:
        a.record()
        terms, _, nj = one()
        b.record()
    torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev) / len(ev)
    tt = terms.cpu()
    return {"workload": f"C1: data/data_2d_artificial.csv N={N} D=2 K={K} V={args[1].shape[1]} M={M} rbf, train() defaults",
            "value": N / (ms * 1e-3), "unit": "observations/s", "ms_per_step": ms, "njitter": int(nj),
            "loss": -float(tt[0] + tt[3] + tt[2] - tt[1]) / N,
            "note": "latency-bound: the reference's jitter escalation (utils.py:27-40) lands on level 5 -- one full prologue at level 0, "
                    "five fp32 status-only factorisations, one full prologue at level 5, a status read-back after each"}


def particles_record(dev):
    """Shared-contraction particles at the C2 shape (scripts/mvco.py:136 runs num_particles=10): one particle vs ten."""
    import torch
    cfg = dict(CONFIGS["C2"])
    b = Bench("C2", cfg, dev, 0, 1)
    e10 = torch.randn(10, cfg["K"], b.n_local, device=dev, generator=torch.Generator(device=dev).manual_seed(77))
    ms1, _, _ = b.timed(lambda: b.step(), 5, 3)
    ms10, _, _ = b.timed(lambda: b.step(eps=e10), 5, 3)
    del b
    return {"workload": "C2 shape, N=100000", "ms_per_step_1_particle": ms1, "ms_per_step_10_particles": ms10,
            "ratio": ms10 / ms1, "note": "one prologue and one pass of every contraction whatever the particle count; "
                                         "the per-observation chain runs once per particle"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C4", choices=sorted(CONFIGS))
    ap.add_argument("--obs", type=int, default=0, help="override the number of observations (debugging)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip svi_step / other_configs / particles")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.obs:
        cfg["N"] = args.obs
    if args.impl == "reference":
        run_reference(args, args.config, cfg)
        return

    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl ours needs a B200: gdrf_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warmup = max(3, args.warmup)
    steps = max(1, args.steps)
    # C5 is BASELINE's 8-GPU weak-scaling stress config (1 M observations per GPU)
    scaling = "weak" if args.config == "C5" else "strong"
    if args.config == "C5" and not args.obs:
        cfg["N"] = 1_000_000 * world
    N, D, K, V = cfg["N"], cfg["D"], cfg["K"], cfg["V"]
    b = Bench(args.config, cfg, dev, rank, world, int(os.environ.get("GDRF_BENCH_FLAGS", "0")))
    M = b.M

    # ---------------- device-resident throughput ("value") ----------------
    for _ in range(warmup):
        b.step()
    sampler = ClockSampler(local_rank)
    sampler.start()
    t_wall = time.perf_counter()
    ms_per_step, ms_mine, (terms, _), prof, launches = b.profiled(steps, 0)
    t_wall = time.perf_counter() - t_wall
    clocks = sampler.stop()
    value = N / (ms_per_step * 1e-3)
    t = terms.cpu()
    loss = -float(t[0] + t[3] + t[2] - t[1]) / N

    e2e = None if args.no_e2e else b.e2e(steps)
    extras = not args.no_extras and not args.obs
    svi = b.svi_step(min(steps, 5), 2) if extras else None

    roofline = roofline_block(prof, steps, b.n_local, ms_mine, M, K, b.flags, b._lib) if rank == 0 else None
    n_local, flags = b.n_local, b.flags
    del b
    torch.cuda.empty_cache()

    # BASELINE configs[4]: the C5 stress shape, weak scaling at 1 M observations per GPU on 8 GPUs (all ranks take part)
    c5_weak = None
    if extras and world == 8 and args.config == "C4":
        from gdrf_b200.elbo import release_workspaces
        release_workspaces()
        torch.cuda.empty_cache()
        c5 = dict(CONFIGS["C5"])
        c5["N"] = 1_000_000 * world
        b5 = Bench("C5", c5, dev, rank, world)
        ms5, _, (t5, _) = b5.timed(b5.step, 2, 3)
        t5 = t5.cpu()
        c5_weak = {"workload": f"C5: N={c5['N']} D=3 K=64 V=1024 M=2048 matern52, 1 M observations per GPU x {world} GPUs",
                   "scaling": "weak", "value": c5["N"] / (ms5 * 1e-3), "unit": "observations/s", "ms_per_step": ms5,
                   "steps": 2, "warmup": 3, "loss": -float(t5[0] + t5[3] + t5[2] - t5[1]) / c5["N"],
                   "all_reduce_bytes": 4 * (64 * 2048 * 2048 + 64 * 2048 + 64 * 1024 + 2048 * 3 + 3 + 8)}
        del b5
        release_workspaces()
        torch.cuda.empty_cache()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    other, particles = None, None
    if extras and world == 1 and args.config == "C4":
        from gdrf_b200.elbo import release_workspaces
        other = {}
        for name, n_override in (("C2", None), ("C3", None), ("C5", 131072)):
            release_workspaces()
            torch.cuda.empty_cache()
            c = dict(CONFIGS[name])
            if n_override:
                c["N"] = n_override
            other[name] = other_config_record(name, c, dev, None)
            if n_override:
                other[name]["workload"] += " (shape of C5 at a bounded N; throughput does not depend on N beyond the M x M prologue)"
        other["C1"] = c1_record(dev)
        release_workspaces()
        torch.cuda.empty_cache()
        particles = particles_record(dev)
        release_workspaces()
        torch.cuda.empty_cache()

    cpu_baseline = None
    if not args.no_cpu_baseline:
        rows = 2000 if K * M > 8192 else 8000
        cpu_baseline = cpu_reference_rate(cfg, min(rows, N), 2, 1)
        cpu_baseline.pop("ms_per_step", None)

    # the full-size result through two code paths with different chunk boundaries (device-resident shard in chunks of
    # 18 944 vs host-streamed sub-shards; at N > 1 both are sums over the ranks): the ELBO is additive over observations
    checks = None
    if e2e is not None and e2e.get("loss") is not None:
        checks = {"e2e_vs_device_resident_loss_rel_diff": abs(e2e["loss"] - loss) / abs(loss),
                  "note": "same 1/N-scaled loss from the device-resident and the host-streamed evaluation of the full data set"}
    line = {"metric": METRIC, "value": value, "unit": "observations/s", "n_gpus": world, "steps": steps,
            "warmup": warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling,
            "vs_baseline": None,
            "dtype": "f32 (error-compensated 16-bit-plane tcgen05 products: fp16x3 forward and backward (22-bit operands), "
                     "bf16x6 whitening W = Kxz L^-T (24-bit); f32 accumulate; f64 per-observation chain and MxM prologue)",
            "data": "synthetic",
            "config": {"workload": f"{args.config}: N={N} D={D} K={K} V={V} M={M} {cfg['kernel']}, sharded by "
                                   f"observation over {world} GPU(s)", "l2": "inputs (ws) exceed L2 every step",
                       "parallelism": f"obs-shard x{world}", "jitter": 1e-4,
                       "api": "gdrf_b200.elbo.elbo_value_and_grads (public); one all-reduce of gradient + terms"},
            "loss": loss, "wall_s_timed": t_wall, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline, "cpu_baseline": cpu_baseline, "svi_step": svi, "other_configs": other,
            "particles": particles, "c5_weak_scaling": c5_weak, "checks": checks,
            "alg_tflops_step": algorithmic_flops_per_obs(M, K, V) * N / (ms_per_step * 1e-3) / 1e12}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
