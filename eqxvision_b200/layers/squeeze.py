"""Squeeze-and-Excitation (reference: layers/squeeze.py:11-61):
x * scale_act(fc2(act(fc1(avgpool_1x1(x))))), both fc layers are 1x1 convs WITH bias."""
from typing import Callable

from .. import functional as F
from .. import nn
from .. import random as jrandom


class SqueezeExcitation(nn.Module):
    avgpool: nn.AdaptiveAvgPool2d
    fc1: nn.Conv2d
    fc2: nn.Conv2d
    activation: nn.Lambda
    scale_activation: nn.Lambda

    def __init__(
        self,
        input_channels: int,
        squeeze_channels: int,
        activation: Callable = None,
        scale_activation: Callable = None,
        *,
        key=None,
    ) -> None:
        k1, k2 = jrandom.split(key, 2)
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.fc1 = nn.Conv2d(input_channels, squeeze_channels, 1, key=k1)
        self.fc2 = nn.Conv2d(squeeze_channels, input_channels, 1, key=k2)
        self.activation = nn.Lambda(F.relu if activation is None else activation)
        self.scale_activation = nn.Lambda(F.sigmoid if scale_activation is None else scale_activation)

    def __call__(self, x, *, key=None):
        s = self.avgpool(x)
        s = self.activation(self.fc1(s), key=key)
        s = self.scale_activation(self.fc2(s), key=key)
        return x * s
