"""Transformer MLP (reference: layers/mlps.py:12-66): fc1 -> act -> drop -> fc2 -> drop.
On the device: two tcgen05 GEMMs, the activation in the first one's epilogue."""
from typing import Callable, Tuple, Union

from .. import nn
from .. import random as jrandom


class MlpProjection(nn.Module):
    fc1: nn.Module
    act: Callable
    drop1: nn.Dropout
    fc2: nn.Module
    drop2: nn.Dropout

    def __init__(
        self,
        in_features: int,
        hidden_features: int = None,
        out_features: int = None,
        lin_layer=nn.Linear,
        act_layer: Callable = None,
        drop: Union[float, Tuple[float]] = 0.0,
        *,
        key=None,
    ):
        hidden = hidden_features or in_features
        out = out_features or in_features
        p1, p2 = drop if isinstance(drop, tuple) else (drop, drop)
        k1, k2 = jrandom.split(key, 2)
        self.fc1 = lin_layer(in_features, hidden, key=k1)
        self.act = act_layer
        self.drop1 = nn.Dropout(p1)
        self.fc2 = lin_layer(hidden, out, key=k2)
        self.drop2 = nn.Dropout(p2)

    def __call__(self, x, *, key=None):
        k1, k2 = jrandom.split(key, 2)
        h = self.act(self.fc1(x))
        h = self.drop1(h, key=k1)
        return self.drop2(self.fc2(h), key=k2)
