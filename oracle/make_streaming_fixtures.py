"""Golden vectors for the streaming sampler, produced by the reference's OWN source lines.

TEST INFRASTRUCTURE (never imported by the product).  ``gdrf/train_script.py`` cannot be imported here (wandb, pyro,
fire, holoviews are absent), so this script cuts the body of the streaming sub-epoch loop -- the statements between
``for subepoch in range(...)`` and ``xs_stream = xs[selection, ...]`` (train_script.py:396-452) -- out of the file
where it lies under /root/reference, dedents it and executes it unmodified in a namespace that supplies what the
enclosing function would (``np``, ``wandb.config``, ``epoch``, ``n_data``, ``epochs``, ``streaming_batch``,
``streaming_weight``), with the global numpy generator seeded.  The probabilities ``p`` and the drawn ``selection``
of every case are written to tests/golden/ref_streaming.json.

    python -m oracle.make_streaming_fixtures
"""
import json
import os
import textwrap
import types

import numpy as np

REF = "/root/reference/gdrf/train_script.py"
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref_streaming.json")


def reference_snippet() -> str:
    lines = open(REF).read().split("\n")
    start = next(i for i, l in enumerate(lines) if "for subepoch in range(wandb.config.streaming_subepochs)" in l) + 1
    end = next(i for i, l in enumerate(lines) if l.strip().startswith("xs_stream = xs[selection, ...]"))
    return textwrap.dedent("\n".join(lines[start:end]))


CASES = [
    # inference, epoch, n_data, epochs, size, truncate, exp, weight, batch_splits, seed
    ("uniform", 0, 50, 50, 1, -1, 1.0, 0.1, -1, 1),
    ("uniform", 17, 50, 50, 8, -1, 1.0, 0.1, -1, 2),
    ("now", 9, 50, 50, 4, -1, 1.0, 0.1, -1, 3),
    ("exp", 30, 64, 64, 16, -1, 0.25, 0.1, -1, 4),
    ("exp", 30, 64, 64, 1, 10, 0.5, 0.1, -1, 5),
    ("uniform_now", 12, 40, 40, 6, -1, 1.0, 0.3, -1, 6),
    ("exp_now", 25, 40, 40, 6, 12, 0.7, 0.25, -1, 7),
    ("uniform_exp", 39, 40, 40, 32, -1, 0.1, 0.4, -1, 8),
    ("uniform", 3, 1000, 10, 64, -1, 1.0, 0.1, 10, 9),       # streaming_batch: effective epoch = epoch * n_data // epochs
    ("exp", 7, 1000, 10, 64, 200, 0.01, 0.1, 10, 10),
    ("uniform_exp", 0, 5, 5, 3, 2, 2.0, 0.9, -1, 11),
]


def main():
    code = compile(reference_snippet(), REF, "exec")
    out = []
    for (inf, epoch, n_data, epochs, size, trunc, ex, wt, splits, seed) in CASES:
        cfg = types.SimpleNamespace(streaming_inference=inf, streaming_truncate=trunc, streaming_exp=ex,
                                    streaming_weight=wt, streaming_size=size)
        ns = dict(np=np, wandb=types.SimpleNamespace(config=cfg), epoch=epoch, n_data=n_data, epochs=epochs,
                  streaming_batch=splits > 0, streaming_weight=wt)
        np.random.seed(seed)
        exec(code, ns)
        sel = np.atleast_1d(np.asarray(ns["selection"])).astype(np.int64)
        out.append(dict(inference=inf, epoch=epoch, n_data=n_data, epochs=epochs, size=size, truncate=trunc, exp=ex,
                        weight=wt, batch=splits > 0, seed=seed, n_stream=int(ns["n_stream"]),
                        p=[float(q) for q in ns["p"]], selection=[int(s) for s in sel]))
    with open(OUT, "w") as f:
        json.dump(out, f)
    print("wrote", OUT, len(out), "cases")


if __name__ == "__main__":
    main()
