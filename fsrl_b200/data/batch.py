"""A small attribute-dict with array slicing: the part of ``tianshou.data.Batch`` the
reference's hot path relies on (keyword construction, attribute access, ``update``, ``get``,
``pop``, integer/slice/mask indexing, ``len``)."""
from __future__ import annotations

from typing import Any

import numpy as np
import torch


class Batch:
    def __init__(self, *args, **kwargs):
        if args:
            if len(args) != 1:
                raise TypeError("Batch takes at most one positional argument")
            src = args[0]
            if isinstance(src, Batch):
                src = src.__dict__
            for k, v in dict(src).items():
                self.__dict__[k] = _wrap(v)
        for k, v in kwargs.items():
            self.__dict__[k] = _wrap(v)

    # mapping-ish
    def keys(self):
        return self.__dict__.keys()

    def items(self):
        return self.__dict__.items()

    def values(self):
        return self.__dict__.values()

    def get(self, k, d=None):
        return self.__dict__.get(k, d)

    def pop(self, k, *d):
        return self.__dict__.pop(k, *d)

    def update(self, *args, **kwargs):
        for a in args:
            for k, v in (a.items() if hasattr(a, "items") else a):
                self.__dict__[k] = _wrap(v)
        for k, v in kwargs.items():
            self.__dict__[k] = _wrap(v)

    def __contains__(self, k):
        return k in self.__dict__

    def is_empty(self):
        return len(self.__dict__) == 0

    def __getitem__(self, idx):
        if isinstance(idx, str):
            return self.__dict__[idx]
        out = Batch()
        for k, v in self.__dict__.items():
            if isinstance(v, Batch):
                out.__dict__[k] = v[idx] if not v.is_empty() else Batch()
            elif isinstance(v, (np.ndarray, torch.Tensor)):
                out.__dict__[k] = v[idx]
            else:
                out.__dict__[k] = v
        return out

    def __setitem__(self, idx, value):
        if isinstance(idx, str):
            self.__dict__[idx] = _wrap(value)
            return
        for k, v in value.items():
            self.__dict__[k][idx] = v

    def __len__(self):
        for v in self.__dict__.values():
            if isinstance(v, Batch):
                if not v.is_empty():
                    return len(v)
            elif isinstance(v, (np.ndarray, torch.Tensor)) and v.ndim > 0:
                return len(v)
        return 0

    def __repr__(self):
        return "Batch(" + ", ".join(f"{k}: {type(v).__name__}" for k, v in self.__dict__.items()) + ")"


def _wrap(v: Any):
    if isinstance(v, dict):
        return Batch(v)
    return v


def to_numpy(x):
    if isinstance(x, torch.Tensor):
        return x.detach().cpu().numpy()
    if isinstance(x, Batch):
        return Batch({k: to_numpy(v) for k, v in x.items()})
    if isinstance(x, (list, tuple)):
        return np.asarray([to_numpy(e) for e in x])
    return np.asarray(x)


def to_torch_as(x, y: torch.Tensor):
    return torch.as_tensor(x, dtype=y.dtype, device=y.device)
