"""Checkpoint interchange with the reference's training loop.

The reference saves ``{"epoch", "best_fitness", "model": deepcopy(model).half(), "optimizer", "wandb_id"}``
(``gdrf/train_script.py:490-499``) and resumes with ``csd = ckpt["model"].float().state_dict()``, an intersection on
matching keys *and shapes* (``gdrf/utils/general.py:455-461``) and ``load_state_dict(csd, strict=False)``
(``train_script.py:338-347``).  The drop-in keeps the reference's parameter names (``<name>_unconstrained``, the way
``PyroParam`` stores them), so the ``state_dict`` of either side loads into the other.  A pickled *module* of the
reference can only be opened where gdrf and pyro are importable; what is exchanged here is therefore the state dict --
taken from ``ckpt["model"]`` whether that is a module or already a dict.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Iterable, Mapping, Optional

import torch


def intersect_dicts(da: Mapping, db: Mapping, exclude: Iterable[str] = ()) -> dict:
    """gdrf/utils/general.py:455-461: entries of ``da`` whose key is in ``db`` with the same shape."""
    return {k: v for k, v in da.items() if k in db and not any(x in k for x in exclude) and v.shape == db[k].shape}


def _state_dict_of(obj) -> Mapping:
    if isinstance(obj, Mapping):
        return obj
    if hasattr(obj, "state_dict"):
        return obj.state_dict()
    raise TypeError(f"cannot take a state dict from {type(obj).__name__}")


def make_checkpoint(model: torch.nn.Module, epoch: int, best_fitness: float, optimizer_state=None,
                    wandb_id: Optional[str] = None) -> dict:
    """The reference's checkpoint dictionary with the model as an fp16 state dict (``deepcopy(model).half()``)."""
    sd = OrderedDict((k, v.detach().to("cpu", torch.float16) if v.is_floating_point() else v.detach().cpu())
                     for k, v in model.state_dict().items())
    return {"epoch": int(epoch), "best_fitness": float(best_fitness), "model": sd, "optimizer": optimizer_state,
            "wandb_id": wandb_id}


def load_checkpoint(model: torch.nn.Module, ckpt: Mapping, exclude: Iterable[str] = ()):
    """train_script.py:338-357: transfer every matching entry (as fp32), return ``(n_transferred, start_epoch,
    best_fitness)``; ``best_fitness`` is only taken over when the checkpoint carries optimiser state, as there."""
    csd = {k: (v.float() if torch.is_tensor(v) and v.is_floating_point() else v)
           for k, v in _state_dict_of(ckpt["model"]).items()}
    own = model.state_dict()
    csd = intersect_dicts(csd, own, exclude=exclude)
    model.load_state_dict(csd, strict=False)
    best = float(ckpt["best_fitness"]) if ckpt.get("optimizer") is not None else float("-inf")
    return len(csd), int(ckpt.get("epoch", -1)) + 1, best


def strip_optimizer(ckpt: dict) -> dict:
    """gdrf/utils/general.py:420-431 on the dictionary: drop optimiser state, epoch = -1, model in fp16."""
    out = dict(ckpt)
    for k in ("optimizer", "training_results", "wandb_id", "ema", "updates"):
        out[k] = None
    out["epoch"] = -1
    out["model"] = OrderedDict((k, v.half() if torch.is_tensor(v) and v.is_floating_point() else v)
                               for k, v in _state_dict_of(ckpt["model"]).items())
    return out
