"""Minimal stand-in for `jax.random` keys.

The reference threads PRNG keys through every `__call__` (e.g. resnet.py:343, vit.py:267). In the
inference forward pass they are dead (Dropout / DropPath return their input), so a key here is just
an integer seed that can be split; it is used only for parameter initialisation.
"""
from __future__ import annotations

from typing import List, Sequence, Union

import torch


class PRNGKey:
    __slots__ = ("seed",)

    def __init__(self, seed: int = 0):
        self.seed = int(seed) & 0x7FFFFFFFFFFFFFFF

    def __repr__(self):
        return f"PRNGKey({self.seed})"

    # jax keys are arrays of shape (2,); `keys.shape[0]`-style code paths are not needed here
    def generator(self) -> torch.Generator:
        g = torch.Generator(device="cpu")
        g.manual_seed(self.seed)
        return g


def _mix(a: int, b: int) -> int:
    # splitmix64-style mixing
    z = (a * 0x9E3779B97F4A7C15 + b + 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return z ^ (z >> 31)


def as_key(key) -> PRNGKey:
    if isinstance(key, PRNGKey):
        return key
    if key is None:
        return PRNGKey(0)
    if isinstance(key, int):
        return PRNGKey(key)
    # arrays / tensors / anything else: hash the flattened content
    try:
        import numpy as np

        arr = np.asarray(key).ravel()
        s = 0
        for v in arr.tolist()[:4]:
            s = _mix(s, int(v))
        return PRNGKey(s)
    except Exception:  # noqa: BLE001
        return PRNGKey(hash(key))


def split(key, num: int = 2) -> List[PRNGKey]:
    k = as_key(key)
    return [PRNGKey(_mix(k.seed, i + 1)) for i in range(num)]


def uniform(key, shape: Sequence[int], minval: float = 0.0, maxval: float = 1.0) -> torch.Tensor:
    g = as_key(key).generator()
    return torch.rand(tuple(shape), generator=g) * (maxval - minval) + minval


def normal(key, shape: Sequence[int]) -> torch.Tensor:
    return torch.randn(tuple(shape), generator=as_key(key).generator())


def truncated_normal(key, lower: float, upper: float, shape: Sequence[int]) -> torch.Tensor:
    t = torch.empty(tuple(shape))
    if lower >= upper:  # degenerate bounds (swin.py:303-312 passes lower=2, upper=2)
        return t.fill_(float(lower))
    torch.nn.init.trunc_normal_(t, mean=0.0, std=1.0, a=lower, b=upper, generator=as_key(key).generator())
    return t


KeyLike = Union[PRNGKey, int, None]
