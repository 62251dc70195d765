"""Development tool: repeats the C1 evaluation (cold jitter search and speculative search) and reports any run whose
ELBO terms or gradients differ from the first one by more than round-off.   python -m tests.stress_c1 [iterations [poison]]
poison: a byte value the whole workspace is filled with before every run (0xff: NaN in every float format, 0x7b: large finite
values) -- any read of scratch memory the step did not write itself then shows."""
import sys

import torch

from gdrf_b200 import elbo as E
from oracle import gdrf_oracle as O
from tests.helpers import load_golden
from tests.test_gpu_parity import _run


def main(iters, poison=None):
    inp, d = load_golden("c1_artificial2d")
    ref_t, ref_g, bad = None, None, 0
    for it in range(iters):
        if it % 2 == 0:
            E._JITTER_HINTS.clear()          # cold search on even iterations, speculation on odd ones
        if poison is not None and it > 0:    # every scratch byte the library did not write itself this step is garbage
            for ws in E._WORKSPACES.values():
                ws.fill_(poison)
        t, g, nj = _run(inp)
        if ref_t is None:
            ref_t, ref_g = t, g
            print("reference terms", [float(x) for x in t], "njitter", nj)
            continue
        dt = ((t - ref_t).abs() / ref_t.abs().clamp(min=1.0)).max().item()
        dg = max(O.rel_err(g[k], ref_g[k]) for k in g)
        if nj != 5 or not (dt < 1e-9) or not (dg < 1e-4):
            bad += 1
            print(f"iteration {it} ({'cold' if it % 2 == 0 else 'speculative'}): njitter {nj}, terms {[float(x) for x in t]}, "
                  f"max term rel diff {dt:.3e}, max gradient rel diff {dg:.3e}",
                  {k: f"{O.rel_err(g[k], ref_g[k]):.1e}" for k in g})
    print(f"{bad} deviating runs of {iters - 1}")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 300, int(sys.argv[2], 0) if len(sys.argv) > 2 else None)
