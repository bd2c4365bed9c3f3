"""Lite R-ASPP (reference: models/segmentation/lraspp.py). Head (lraspp.py:71-116):
`x = cbr(high) * scale(high)` with cbr = conv1x1 -> BN -> ReLU and scale = global pool -> conv1x1 -> sigmoid,
bilinear resize to the low-level map, `low_classifier(low) + high_classifier(x)`; the model resizes the result to
the input resolution and returns `(None, out)` (lraspp.py:67-68). On the device: two fused 1x1 GEMMs, the pool,
one channel-gate kernel, the bilinear kernel, and the two classifier GEMMs fused through the residual epilogue."""
from typing import Callable, Optional

from ... import functional as F
from ... import nn
from ... import random as jrandom
from ...experimental import intermediate_layer_getter
from ...utils import load_torch_weights
from ..classification.mobilenetv3 import mobilenet_v3_large


class LRASPPHead(nn.Module):
    cbr: nn.Module
    scale: nn.Module
    low_classifier: nn.Module
    high_classifier: nn.Module

    def __init__(self, low_channels: int, high_channels: int, num_classes: int, inter_channels: int, key=None) -> None:
        keys = jrandom.split(jrandom.PRNGKey(0) if key is None else key, 4)
        self.cbr = nn.Sequential([
            nn.Conv2d(high_channels, inter_channels, 1, use_bias=False, key=keys[0]),
            nn.BatchNorm(inter_channels, axis_name="batch"),
            nn.Lambda(F.relu),
        ])
        self.scale = nn.Sequential([
            nn.AdaptiveAvgPool2d(1),
            nn.Conv2d(high_channels, inter_channels, 1, use_bias=False, key=keys[1]),
            nn.Lambda(F.sigmoid),
        ])
        self.low_classifier = nn.Conv2d(low_channels, num_classes, 1, key=keys[2])
        self.high_classifier = nn.Conv2d(inter_channels, num_classes, 1, key=keys[3])

    def __call__(self, x, *, key=None):
        low, high = x[0], x[1]
        y = self.cbr(high)
        s = self.scale(high)
        y = y * s
        y = F.resize_bilinear(y, low.shape[-2], low.shape[-1])
        return self.low_classifier(low) + self.high_classifier(y)


class LRASPP(nn.Module):
    """Lite R-ASPP network ("Searching for MobileNetV3"), lraspp.py:15-68."""
    backbone: nn.Module
    classifier: nn.Module

    def __init__(self, backbone: nn.Module, low_channels: int, high_channels: int, num_classes: int,
                 inter_channels: int = 128, key=None) -> None:
        self.backbone = backbone
        self.classifier = LRASPPHead(low_channels, high_channels, num_classes, inter_channels, key=key)

    def __call__(self, x, *, key=None):
        _, features = self.backbone(x, key=key)
        out = self.classifier(features)
        out = F.resize_bilinear(out, x.shape[-2], x.shape[-1])
        return None, out


def lraspp_mobilenet_v3_large(
    num_classes: Optional[int] = 21,
    backbone: nn.Module = None,
    intermediate_layers: Callable = None,
    torch_weights: str = None,
    *,
    key=None,
) -> LRASPP:
    """LRASPP with a (dilated) MobileNetV3-Large backbone (lraspp.py:119-175); the taps default to
    `backbone.features` indices [4, 16] (C2 and the last ConvNormActivation)."""
    key = jrandom.PRNGKey(0) if key is None else key
    if num_classes is None:
        num_classes = 21
    if backbone is None:
        backbone = mobilenet_v3_large(dilated=True)
    if intermediate_layers is None:
        intermediate_layers = lambda x: [4, 16]  # noqa: E731
    backbone = backbone.features
    num_channels = [backbone.layers[y].out_channels for y in intermediate_layers(backbone)]
    backbone = intermediate_layer_getter(backbone, intermediate_layers)
    model = LRASPP(backbone, num_channels[0], num_channels[1], num_classes=num_classes, key=key)
    if torch_weights:
        return load_torch_weights(model, torch_weights)
    return model
