"""MobileNetV2 (reference: models/classification/mobilenetv2.py).

_InvertedResidual.conv = [expand 1x1 CNA]? (skipped when expand_ratio == 1), depthwise 3x3 CNA (stride 1|2),
project 1x1 Conv2d (no bias) -> BatchNorm; `x + conv(x)` when stride 1 and inp == oup.
Reference quirk kept on purpose: every activation is plain ReLU (mobilenetv2.py:54,67,176,200),
not torchvision's ReLU6.
Device lowering: 1x1 convs = tcgen05 GEMMs with BN + ReLU (+ residual) in the epilogue, depthwise 3x3 = channels-last
strip stencil (csrc/depthwise.cu), global pool + Linear head as for MobileNetV3.
"""
from typing import Any, Callable, List, Optional

from ... import functional as F
from ... import nn
from ... import random as jrandom
from ...layers import ConvNormActivation
from ...utils import _make_divisible, load_torch_weights


class _InvertedResidual(nn.Module):
    stride: int
    use_res_connect: int
    conv: nn.Sequential
    out_channels: int

    def __init__(self, inp: int, oup: int, stride: int, expand_ratio: int,
                 norm_layer: Optional[Callable[..., nn.Module]] = None, *, key=None) -> None:
        keys = jrandom.split(key, 3)
        self.stride = stride
        assert stride in [1, 2]
        norm_layer = nn.BatchNorm if norm_layer is None else norm_layer
        hidden_dim = int(round(inp * expand_ratio))
        self.use_res_connect = self.stride == 1 and inp == oup
        seq: List[nn.Module] = []
        if expand_ratio != 1:
            seq.append(ConvNormActivation(inp, hidden_dim, kernel_size=1, norm_layer=norm_layer,
                                          activation_layer=F.relu, key=keys[0]))
        seq.extend([
            ConvNormActivation(hidden_dim, hidden_dim, stride=stride, groups=hidden_dim, norm_layer=norm_layer,
                               activation_layer=F.relu, key=keys[1]),
            nn.Conv2d(hidden_dim, oup, 1, 1, 0, use_bias=False, key=keys[2]),
            norm_layer(oup, axis_name="batch"),
        ])
        self.conv = nn.Sequential(seq)
        self.out_channels = oup

    def __call__(self, x, *, key=None):
        if self.use_res_connect:
            return x + self.conv(x, key=key)
        return self.conv(x, key=key)


class MobileNetV2(nn.Module):
    """`torchvision.models.mobilenetv2` layout as ported by the reference (mobilenetv2.py:91-227)."""

    features: nn.Sequential
    classifier: nn.Sequential
    pool: nn.AdaptiveAvgPool2d

    def __init__(
        self,
        num_classes: int = 1000,
        width_mult: float = 1.0,
        inverted_residual_setting: Optional[List[List[int]]] = None,
        round_nearest: int = 8,
        block=None,
        norm_layer=None,
        dropout: float = 0.2,
        *,
        key=None,
    ) -> None:
        keys = jrandom.split(jrandom.PRNGKey(0) if key is None else key, 2)
        block = _InvertedResidual if block is None else block
        norm_layer = nn.BatchNorm if norm_layer is None else norm_layer
        if inverted_residual_setting is None:
            inverted_residual_setting = [  # t, c, n, s
                [1, 16, 1, 1], [6, 24, 2, 2], [6, 32, 3, 2], [6, 64, 4, 2], [6, 96, 3, 1], [6, 160, 3, 2], [6, 320, 1, 1]]
        if len(inverted_residual_setting) == 0 or len(inverted_residual_setting[0]) != 4:
            raise ValueError("inverted_residual_setting should be non-empty or a 4-element list, "
                             f"got {inverted_residual_setting}")
        input_channel = _make_divisible(32 * width_mult, round_nearest)
        last_channel = _make_divisible(1280 * max(1.0, width_mult), round_nearest)
        features: List[nn.Module] = [ConvNormActivation(3, input_channel, stride=2, norm_layer=norm_layer,
                                                        activation_layer=F.relu, key=keys[0])]
        for t, c, n, s in inverted_residual_setting:
            output_channel = _make_divisible(c * width_mult, round_nearest)
            for i in range(n):
                keys = jrandom.split(keys[1], 2)
                features.append(block(input_channel, output_channel, s if i == 0 else 1, expand_ratio=t,
                                      norm_layer=norm_layer, key=keys[0]))
                input_channel = output_channel
        keys = jrandom.split(keys[1], 2)
        features.append(ConvNormActivation(input_channel, last_channel, kernel_size=1, norm_layer=norm_layer,
                                           activation_layer=F.relu, key=keys[0]))
        self.features = nn.Sequential(features)
        self.classifier = nn.Sequential([nn.Dropout(p=dropout), nn.Linear(last_channel, num_classes, key=keys[1])])
        self.pool = nn.AdaptiveAvgPool2d((1, 1))

    def __call__(self, x, *, key=None):
        k = [None] * 3 if key is None else jrandom.split(key, 3)
        x = self.pool(self.features(x, key=k[0]), key=k[1])
        return self.classifier(F.ravel(x), key=k[2])


def mobilenet_v2(torch_weights: str = None, **kwargs: Any) -> MobileNetV2:
    """MobileNetV2 ("Inverted Residuals and Linear Bottlenecks", arXiv 1801.04381); mobilenetv2.py:230-244."""
    model = MobileNetV2(**kwargs)
    if torch_weights:
        model = load_torch_weights(model, torch_weights=torch_weights)
    return model
