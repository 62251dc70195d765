"""Times the per-step M x M prologue (gdrf_prologue) and the batched jitter probe with CUDA events:
    python tools/time_prologue.py            # C4 (M = 1024, K = 32) and C1 (M = 625, K = 5) shapes"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gdrf_b200 import _lib
from gdrf_b200.elbo import _Call


def run(M_grid, K, D, jitter, nj, reps=20):
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    pts = [torch.linspace(0, 1, n) for n in M_grid]
    Z = torch.stack([x.flatten() for x in torch.meshgrid(*pts, indexing="ij")]).T.contiguous().to(dev)
    M = Z.shape[0]
    N, V = 256, 4
    S = torch.eye(M).repeat(K, 1, 1).to(dev)
    call = _Call(torch.rand(N, D, generator=g).to(dev), torch.ones(N, V, dtype=torch.int32, device=dev), Z,
                 torch.tensor(25.0, device=dev), torch.tensor([0.75 / (M_grid[0] - 1)], device=dev),
                 torch.zeros(K, M, device=dev), S, torch.tensor(1.0, device=dev), torch.full((K, V), 1.0 / V, device=dev),
                 torch.ones(K, V, device=dev), torch.zeros(K, N, device=dev), 0, 0, _lib.FLAG_CHOL_FP32_STATUS, 0)
    status = torch.zeros(1 + _lib.PROBE_MAX, dtype=torch.int32, device=dev)
    out = {}
    for name, fn in (("prologue", lambda: call._full_prologue(jitter, nj, status)),
                     ("probe8", lambda: call._probe(jitter, 0, 8, status[1:]))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record()
        torch.cuda.synchronize()
        out[name] = round(a.elapsed_time(b) / reps, 4)
    out["status"] = status.tolist()
    return out


if __name__ == "__main__":
    print("C4 shape M=1024 K=32:", run([32, 32], 32, 2, 1e-4, 0))
    print("C1 shape M=625  K=5 :", run([25, 25], 5, 2, 1e-8, 5))
    print("C5 shape M=2048 K=8 :", run([2048], 8, 1, 1e-4, 0))
