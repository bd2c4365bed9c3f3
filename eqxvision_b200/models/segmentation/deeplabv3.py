"""DeepLabV3 (reference: models/segmentation/deeplabv3.py).

ASPP = five branches on the same 2048-channel map (1x1; 3x3 d12 / d24 / d36 with padding == dilation;
global-pool -> 1x1 -> BN -> ReLU -> bilinear broadcast), channel concat (1280), 1x1 projection + BN +
ReLU + Dropout(0.5); head = ASPP -> conv3x3+BN+ReLU -> conv1x1(bias).
Device lowering: all convs are the tcgen05 implicit GEMM with BN/ReLU folded; the dilated taps that
fall entirely into the zero padding are skipped per tile; the five branches store directly into
channel slices of one 1280-channel buffer (no concat pass).
"""
from typing import Callable, List, Optional

from ... import _trace as T
from ... import functional as F
from ... import nn
from ... import random as jrandom
from .fcn import FCNHead, _assemble
from ._utils import _SimpleSegmentationModel


class DeepLabV3(_SimpleSegmentationModel):
    """Ported from `torchvision.models.segmentation.deeplabv3`"""


def _conv_bn_relu(cin, cout, k, key, dilation=1):
    pad = 0 if k == 1 else dilation
    return [nn.Conv2d(cin, cout, k, padding=pad, dilation=dilation, use_bias=False, key=key),
            nn.BatchNorm(cout, axis_name="batch"), nn.Lambda(F.relu)]


class ASPPConv(nn.Sequential):
    def __init__(self, in_channels: int, out_channels: int, dilation: int, key=None) -> None:
        super().__init__(_conv_bn_relu(in_channels, out_channels, 3, key, dilation))


class ASPPPooling(nn.Sequential):
    def __init__(self, in_channels: int, out_channels: int, key=None) -> None:
        super().__init__([nn.AdaptiveAvgPool2d(1)] + _conv_bn_relu(in_channels, out_channels, 1, key))

    def __call__(self, x, *, key=None):
        h, w = x.shape[-2:]
        return F.resize_bilinear(super().__call__(x), h, w)


class ASPP(nn.Module):
    convs: nn.Module
    project: nn.Module

    def __init__(self, in_channels: int, atrous_rates: List[int], out_channels: int = 256, key=None) -> None:
        key = jrandom.PRNGKey(0) if key is None else key
        keys = jrandom.split(key, len(atrous_rates) + 3)
        branches = [nn.Sequential(_conv_bn_relu(in_channels, out_channels, 1, keys[0]))]
        branches += [ASPPConv(in_channels, out_channels, rate, key=keys[i + 1])
                     for i, rate in enumerate(tuple(atrous_rates))]
        branches.append(ASPPPooling(in_channels, out_channels, key=keys[-2]))
        self.convs = nn.Sequential(branches)
        self.project = nn.Sequential(
            _conv_bn_relu(len(self.convs) * out_channels, out_channels, 1, keys[-1]) + [nn.Dropout(0.5)])

    def __call__(self, x, *, key=None):
        return self.project(T.concat_channels([branch(x) for branch in self.convs.layers]), key=key)


class DeepLabHead(nn.Sequential):
    def __init__(self, in_channels: int, out_channels: int, key=None) -> None:
        k1, k2, k3 = jrandom.split(key, 3)
        super().__init__([
            ASPP(in_channels, [12, 24, 36], key=k1),
            nn.Conv2d(256, 256, 3, padding=1, use_bias=False, key=k2),
            nn.BatchNorm(256, axis_name="batch"),
            nn.Lambda(F.relu),
            nn.Conv2d(256, out_channels, 1, key=k3),
        ])


def deeplabv3(
    num_classes: Optional[int] = 21,
    backbone: nn.Module = None,
    intermediate_layers: Callable = None,
    classifier_module: nn.Module = None,
    classifier_in_channels: int = 2048,
    aux_classifier_module: nn.Module = None,
    aux_in_channels: int = 1024,
    silence_layers: Callable = None,
    torch_weights: str = None,
    *,
    key=None,
) -> DeepLabV3:
    """DeepLabV3-ResNet50 by default (deeplabv3.py:138-227). Sample call:

        net = deeplabv3(intermediate_layers=lambda x: [x.layer3, x.layer4], aux_in_channels=1024,
                        torch_weights=...)

    `intermediate_layers` is effectively required (it is called unconditionally, deeplabv3.py:195)."""
    return _assemble(DeepLabV3, backbone, intermediate_layers, silence_layers,
                     classifier_module or DeepLabHead, classifier_in_channels,
                     aux_classifier_module or FCNHead, aux_in_channels, num_classes, torch_weights, key)
