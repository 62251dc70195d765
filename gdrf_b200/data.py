"""CSV ingest exactly as the reference's training driver does it (``gdrf/train_script.py:251-273``): the first
``dimensions`` columns are the index (observation coordinates), every other column is a category count;
coordinates are min-max normalised per dimension, counts are ``fillna(0).astype(int)``."""
from __future__ import annotations

from typing import List, Tuple

import numpy as np
import torch


def load_counts_csv(path: str, dimensions: int, device="cpu") -> Tuple[torch.Tensor, torch.Tensor, List[Tuple[float, float]]]:
    """Returns (xs float32 [N, D] in [0, 1], ws int32 [N, V], world = [(min, max)] per dimension of xs)."""
    import pandas as pd
    # parse_dates=True as the reference passes it: a date-indexed 1-D file (the MVCO hourly series) becomes datetime64,
    # whose differences normalise to [0, 1] like any numeric index
    dataset = pd.read_csv(filepath_or_buffer=path, index_col=list(range(dimensions)), header=0,
                          parse_dates=True).fillna(0).astype(int)
    index = dataset.index
    index = index.values if dimensions == 1 else np.array(index.to_list())
    index = index - index.min(axis=-dimensions, keepdims=True)
    index = index / index.max(axis=-dimensions, keepdims=True)
    xs = torch.from_numpy(np.asarray(index, dtype=np.float64)).float().to(device)
    if dimensions == 1:
        xs = xs.unsqueeze(-1)
    ws = torch.from_numpy(dataset.values).int().to(device)
    min_xs = xs.min(dim=0).values.detach().cpu().numpy().tolist()
    max_xs = xs.max(dim=0).values.detach().cpu().numpy().tolist()
    return xs, ws, list(zip(min_xs, max_xs))
