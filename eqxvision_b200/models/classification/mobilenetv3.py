"""MobileNetV3 large / small (reference: models/classification/mobilenetv3.py).

_InvertedResidual.block = [expand 1x1 CNA]?, depthwise kxk CNA (k in {3,5}; stride forced to 1 when
dilated, mobilenetv3.py:88), [SqueezeExcitation(hard-sigmoid gate, squeeze = make_divisible(exp//4, 8))]?,
project 1x1 CNA (no activation); `+ x` when stride 1 and in == out. BatchNorm eps 1e-3
(mobilenetv3.py:189). Classifier: Linear -> hard_swish -> Dropout -> Linear.
Device lowering as for EfficientNet (GEMM / depthwise stencil / SE), hard-swish and hard-sigmoid are
GEMM / stencil epilogues.
"""
from functools import partial
from typing import Any, Callable, List, Optional, Sequence

from ... import functional as F
from ... import nn
from ... import random as jrandom
from ...layers import ConvNormActivation
from ...layers import SqueezeExcitation as SElayer
from ...utils import _make_divisible, load_torch_weights


class _InvertedResidualConfig:
    """one row of Tables 1/2 of the MobileNetV3 paper"""

    def __init__(self, input_channels: int, kernel: int, expanded_channels: int, out_channels: int, use_se: bool,
                 activation: str, stride: int, dilation: int, width_mult: float):
        self.input_channels = self.adjust_channels(input_channels, width_mult)
        self.kernel = kernel
        self.expanded_channels = self.adjust_channels(expanded_channels, width_mult)
        self.out_channels = self.adjust_channels(out_channels, width_mult)
        self.use_se = use_se
        self.use_hs = activation == "HS"
        self.stride = stride
        self.dilation = dilation

    @staticmethod
    def adjust_channels(channels: int, width_mult: float):
        return _make_divisible(channels * width_mult, 8)


class _InvertedResidual(nn.Module):
    use_res_connect: int
    block: nn.Sequential
    out_channels: int

    def __init__(self, cnf: _InvertedResidualConfig, norm_layer: Callable[..., nn.Module],
                 se_layer: Callable[..., nn.Module] = partial(SElayer, scale_activation=F.hard_sigmoid),
                 *, key=None):
        keys = jrandom.split(key, 4)
        if not (1 <= cnf.stride <= 2):
            raise ValueError("illegal stride value")
        self.use_res_connect = cnf.stride == 1 and cnf.input_channels == cnf.out_channels
        act = F.hard_swish if cnf.use_hs else F.relu
        exp = cnf.expanded_channels
        seq: List[nn.Module] = []
        if exp != cnf.input_channels:
            seq.append(ConvNormActivation(cnf.input_channels, exp, kernel_size=1, norm_layer=norm_layer,
                                          activation_layer=act, key=keys[0]))
        seq.append(ConvNormActivation(exp, exp, kernel_size=cnf.kernel, stride=1 if cnf.dilation > 1 else cnf.stride,
                                      dilation=cnf.dilation, groups=exp, norm_layer=norm_layer,
                                      activation_layer=act, key=keys[1]))
        if cnf.use_se:
            seq.append(se_layer(exp, _make_divisible(exp // 4, 8), key=keys[2]))
        seq.append(ConvNormActivation(exp, cnf.out_channels, kernel_size=1, norm_layer=norm_layer,
                                      activation_layer=None, key=keys[3]))
        self.block = nn.Sequential(seq)
        self.out_channels = cnf.out_channels

    def __call__(self, x, *, key=None):
        y = self.block(x, key=key)
        return y + x if self.use_res_connect else y


class MobileNetV3(nn.Module):
    """`torchvision.models.mobilenetv3` layout as ported by the reference (mobilenetv3.py:135-247)."""

    features: nn.Sequential
    avgpool: nn.AdaptiveAvgPool2d
    classifier: nn.Sequential

    def __init__(
        self,
        inverted_residual_setting: List["_InvertedResidualConfig"],
        last_channel: int,
        num_classes: int = 1000,
        block=None,
        norm_layer=None,
        dropout: float = 0.2,
        *,
        key=None,
    ) -> None:
        keys = jrandom.split(jrandom.PRNGKey(0) if key is None else key, 5)
        if not inverted_residual_setting:
            raise ValueError("The inverted_residual_setting should not be empty")
        if not (isinstance(inverted_residual_setting, Sequence)
                and all(isinstance(s, _InvertedResidualConfig) for s in inverted_residual_setting)):
            raise TypeError("The inverted_residual_setting should be List[InvertedResidualConfig]")
        block = _InvertedResidual if block is None else block
        norm_layer = partial(nn.BatchNorm, eps=0.001, momentum=0.01) if norm_layer is None else norm_layer

        stem_out = inverted_residual_setting[0].input_channels
        seq: List[nn.Module] = [ConvNormActivation(3, stem_out, kernel_size=3, stride=2, norm_layer=norm_layer,
                                                   activation_layer=F.hard_swish, key=keys[0])]
        seq += [block(cnf, norm_layer, key=keys[1]) for cnf in inverted_residual_setting]  # one key (mobilenetv3.py:209)
        tail_in = inverted_residual_setting[-1].out_channels
        tail_out = 6 * tail_in
        seq.append(ConvNormActivation(tail_in, tail_out, kernel_size=1, norm_layer=norm_layer,
                                      activation_layer=F.hard_swish, key=keys[2]))
        self.features = nn.Sequential(seq)
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.classifier = nn.Sequential([
            nn.Linear(tail_out, last_channel, key=keys[3]),
            nn.Lambda(F.hard_swish),
            nn.Dropout(p=dropout),
            nn.Linear(last_channel, num_classes, key=keys[4]),
        ])

    def __call__(self, x, *, key=None):
        k1, k2, k3 = jrandom.split(key, 3)
        x = self.avgpool(self.features(x, key=k1), key=k2)
        return self.classifier(F.ravel(x), key=k3)


# (in, kernel, expanded, out, use_se, activation, stride, dilated?, reduced?) — the last two flags mark the
# C4/C5 rows whose width is divided by `reduce_divider` and whose depthwise conv takes `dilation`
_LARGE = [(16, 3, 16, 16, False, "RE", 1), (16, 3, 64, 24, False, "RE", 2), (24, 3, 72, 24, False, "RE", 1),
          (24, 5, 72, 40, True, "RE", 2), (40, 5, 120, 40, True, "RE", 1), (40, 5, 120, 40, True, "RE", 1),
          (40, 3, 240, 80, False, "HS", 2), (80, 3, 200, 80, False, "HS", 1), (80, 3, 184, 80, False, "HS", 1),
          (80, 3, 184, 80, False, "HS", 1), (80, 3, 480, 112, True, "HS", 1), (112, 3, 672, 112, True, "HS", 1)]
_SMALL = [(16, 3, 16, 16, True, "RE", 2), (16, 3, 72, 24, False, "RE", 2), (24, 3, 88, 24, False, "RE", 1),
          (24, 5, 96, 40, True, "HS", 2), (40, 5, 240, 40, True, "HS", 1), (40, 5, 240, 40, True, "HS", 1),
          (40, 5, 120, 48, True, "HS", 1), (48, 5, 144, 48, True, "HS", 1)]


def _mobilenet_v3_conf(arch: str, width_mult: float = 1.0, reduced_tail: bool = False, dilated: bool = False,
                       **kwargs: Any):
    rd = 2 if reduced_tail else 1
    dil = 2 if dilated else 1
    row = partial(_InvertedResidualConfig, width_mult=width_mult)
    adjust = partial(_InvertedResidualConfig.adjust_channels, width_mult=width_mult)
    if arch == "mobilenet_v3_large":
        setting = [row(*r, 1) for r in _LARGE]
        setting += [row(112, 5, 672, 160 // rd, True, "HS", 2, dil),
                    row(160 // rd, 5, 960 // rd, 160 // rd, True, "HS", 1, dil),
                    row(160 // rd, 5, 960 // rd, 160 // rd, True, "HS", 1, dil)]
        last_channel = adjust(1280 // rd)
    elif arch == "mobilenet_v3_small":
        setting = [row(*r, 1) for r in _SMALL]
        setting += [row(48, 5, 288, 96 // rd, True, "HS", 2, dil),
                    row(96 // rd, 5, 576 // rd, 96 // rd, True, "HS", 1, dil),
                    row(96 // rd, 5, 576 // rd, 96 // rd, True, "HS", 1, dil)]
        last_channel = adjust(1024 // rd)
    else:
        raise ValueError(f"Unsupported model type {arch}")
    return setting, last_channel


def _mobilenet_v3(arch, inverted_residual_setting, last_channel, **kwargs: Any):
    return MobileNetV3(inverted_residual_setting, last_channel, **kwargs)


def mobilenet_v3_large(torch_weights: str = None, **kwargs: Any) -> MobileNetV3:
    """MobileNetV3-Large (mobilenetv3.py:353-371). `dilated=True` gives the segmentation backbone."""
    arch = "mobilenet_v3_large"
    dilated = kwargs.pop("dilated", False)
    setting, last_channel = _mobilenet_v3_conf(arch, dilated=dilated, **kwargs)
    model = _mobilenet_v3(arch, setting, last_channel, **kwargs)
    if torch_weights:
        model = load_torch_weights(model, torch_weights=torch_weights)
    return model


def mobilenet_v3_small(torch_weights: str = None, **kwargs: Any) -> MobileNetV3:
    """MobileNetV3-Small (mobilenetv3.py:374-389)."""
    arch = "mobilenet_v3_small"
    setting, last_channel = _mobilenet_v3_conf(arch, **kwargs)
    model = _mobilenet_v3(arch, setting, last_channel, **kwargs)
    if torch_weights:
        model = load_torch_weights(model, torch_weights=torch_weights)
    return model
