"""EfficientNet v1 (B0-B7) and v2 (S/M/L) (reference: models/classification/efficientnet.py).

Block structure and field order follow the reference (positional loader contract):
  _MBConv.block      = [expand 1x1 CNA]? , depthwise kxk CNA, SqueezeExcitation, project 1x1 CNA(no act)
  _FusedMBConv.block = expand kxk CNA + project 1x1 CNA(no act)   |   single kxk CNA
Device lowering: 1x1 convs -> tcgen05 GEMM (BN + SiLU folded), depthwise -> shared-memory-free
channels-last stencil (csrc/depthwise.cu, no tensor cores), SE -> global pool + two tiny GEMMs +
one channel-scale pass, `result += x` -> residual epilogue of the project GEMM.
Reference quirks kept: BatchNorm eps 1e-5 for B0-B4 and 1e-3 for B5-B7 / v2 (efficientnet.py:320-321,
606-713); `_MBConv` builds DropPath(mode="per_channel") (efficientnet.py:177) - inert in inference.
"""
import copy
import math
from functools import partial
from typing import Any, Callable, List, Optional, Sequence, Tuple, Union

from ... import functional as F
from ... import nn
from ... import random as jrandom
from ...layers import ConvNormActivation, DropPath, SqueezeExcitation
from ...utils import _make_divisible, load_torch_weights


class _MBConvConfigData:
    def __init__(self, expand_ratio: float, kernel: int, stride: int, input_channels: int, out_channels: int,
                 num_layers: int, block: Callable[..., nn.Module]):
        self.expand_ratio, self.kernel, self.stride = expand_ratio, kernel, stride
        self.input_channels, self.out_channels, self.num_layers, self.block = \
            input_channels, out_channels, num_layers, block

    @staticmethod
    def adjust_channels(channels: int, width_mult: float, min_value: Optional[int] = None) -> int:
        return _make_divisible(channels * width_mult, 8, min_value)


class _MBConvConfig(_MBConvConfigData):
    """one row of Table 1 (EfficientNet) / Table 4 (EfficientNetV2), scaled by width/depth multipliers"""

    def __init__(self, expand_ratio, kernel, stride, input_channels, out_channels, num_layers,
                 width_mult: float = 1.0, depth_mult: float = 1.0, block=None) -> None:
        super().__init__(expand_ratio, kernel, stride,
                         self.adjust_channels(input_channels, width_mult),
                         self.adjust_channels(out_channels, width_mult),
                         self.adjust_depth(num_layers, depth_mult),
                         _MBConv if block is None else block)

    @staticmethod
    def adjust_depth(num_layers: int, depth_mult: float):
        return int(math.ceil(num_layers * depth_mult))


class _FusedMBConvConfig(_MBConvConfigData):
    def __init__(self, expand_ratio, kernel, stride, input_channels, out_channels, num_layers, block=None) -> None:
        super().__init__(expand_ratio, kernel, stride, input_channels, out_channels, num_layers,
                         _FusedMBConv if block is None else block)


def _residual_call(self, x, key):
    # shared by both block types: block(x) (+ x when stride 1 and in == out), efficientnet.py:180-186
    k1, k2 = jrandom.split(key, 2)
    y = self.block(x, key=k1)
    if self.use_res_connect:
        y = self.stochastic_depth(y, key=k2) + x
    return y


class _MBConv(nn.Module):
    use_res_connect: bool
    block: nn.Sequential
    stochastic_depth: DropPath
    out_channels: int

    def __init__(self, cnf: _MBConvConfig, stochastic_depth_prob: float, norm_layer, se_layer=SqueezeExcitation,
                 *, key=None) -> None:
        if not (1 <= cnf.stride <= 2):
            raise ValueError("illegal stride value")
        keys = jrandom.split(jrandom.PRNGKey(0) if key is None else key, 4)
        self.use_res_connect = cnf.stride == 1 and cnf.input_channels == cnf.out_channels
        act = F.silu
        wide = cnf.adjust_channels(cnf.input_channels, cnf.expand_ratio)
        seq: List[nn.Module] = []
        if wide != cnf.input_channels:  # expand (skipped at ratio 1, efficientnet.py:127)
            seq.append(ConvNormActivation(cnf.input_channels, wide, kernel_size=1, norm_layer=norm_layer,
                                          activation_layer=act, key=keys[0]))
        seq.append(ConvNormActivation(wide, wide, kernel_size=cnf.kernel, stride=cnf.stride, groups=wide,
                                      norm_layer=norm_layer, activation_layer=act, key=keys[1]))
        seq.append(se_layer(wide, max(1, cnf.input_channels // 4), activation=act, key=keys[2]))
        seq.append(ConvNormActivation(wide, cnf.out_channels, kernel_size=1, norm_layer=norm_layer,
                                      activation_layer=None, key=keys[3]))
        self.block = nn.Sequential(seq)
        self.stochastic_depth = DropPath(stochastic_depth_prob, mode="per_channel")
        self.out_channels = cnf.out_channels

    def __call__(self, x, *, key=None):
        return _residual_call(self, x, key)


class _FusedMBConv(nn.Module):
    use_res_connect: bool
    block: nn.Sequential
    stochastic_depth: DropPath
    out_channels: int

    def __init__(self, cnf: _FusedMBConvConfig, stochastic_depth_prob: float, norm_layer, *, key=None) -> None:
        if not (1 <= cnf.stride <= 2):
            raise ValueError("illegal stride value")
        keys = jrandom.split(jrandom.PRNGKey(0) if key is None else key, 3)
        self.use_res_connect = cnf.stride == 1 and cnf.input_channels == cnf.out_channels
        act = F.silu
        wide = cnf.adjust_channels(cnf.input_channels, cnf.expand_ratio)
        if wide != cnf.input_channels:
            seq = [ConvNormActivation(cnf.input_channels, wide, kernel_size=cnf.kernel, stride=cnf.stride,
                                      norm_layer=norm_layer, activation_layer=act, key=keys[0]),
                   ConvNormActivation(wide, cnf.out_channels, kernel_size=1, norm_layer=norm_layer,
                                      activation_layer=None, key=keys[1])]
        else:
            seq = [ConvNormActivation(cnf.input_channels, cnf.out_channels, kernel_size=cnf.kernel,
                                      stride=cnf.stride, norm_layer=norm_layer, activation_layer=act, key=keys[2])]
        self.block = nn.Sequential(seq)
        self.stochastic_depth = DropPath(stochastic_depth_prob, mode="local")
        self.out_channels = cnf.out_channels

    def __call__(self, x, *, key=None):
        return _residual_call(self, x, key)


class EfficientNet(nn.Module):
    """`torchvision.models.efficientnet` layout as ported by the reference (efficientnet.py:269-403)."""

    features: nn.Sequential
    avgpool: nn.AdaptiveAvgPool2d
    classifier: nn.Sequential

    def __init__(
        self,
        inverted_residual_setting: Sequence[Union["_MBConvConfig", "_FusedMBConvConfig"]],
        dropout: float,
        stochastic_depth_prob: float = 0.2,
        num_classes: int = 1000,
        norm_layer=None,
        last_channel: Optional[int] = None,
        *,
        key=None,
    ) -> None:
        if not inverted_residual_setting:
            raise ValueError("The inverted_residual_setting should not be empty")
        if not (isinstance(inverted_residual_setting, Sequence)
                and all(isinstance(s, _MBConvConfigData) for s in inverted_residual_setting)):
            raise TypeError("The inverted_residual_setting should be List[MBConvConfig]")
        keys = jrandom.split(jrandom.PRNGKey(0) if key is None else key, 3)
        norm_layer = nn.BatchNorm if norm_layer is None else norm_layer

        stem_out = inverted_residual_setting[0].input_channels
        seq: List[nn.Module] = [ConvNormActivation(3, stem_out, kernel_size=3, stride=2, norm_layer=norm_layer,
                                                   activation_layer=F.silu, key=keys[0])]
        n_blocks = sum(c.num_layers for c in inverted_residual_setting)
        done = 0
        for cnf in inverted_residual_setting:
            stage: List[nn.Module] = []
            for i in range(cnf.num_layers):
                keys = jrandom.split(keys[1], 2)
                bc = copy.copy(cnf)
                if i > 0:  # only the first block of a stage changes width / resolution
                    bc.input_channels, bc.stride = bc.out_channels, 1
                stage.append(bc.block(bc, stochastic_depth_prob * float(done) / n_blocks, norm_layer, key=keys[0]))
                done += 1
            seq.append(nn.Sequential(stage))
        keys = jrandom.split(keys[1], 2)
        head_in = inverted_residual_setting[-1].out_channels
        head_out = last_channel if last_channel is not None else 4 * head_in
        seq.append(ConvNormActivation(head_in, head_out, kernel_size=1, norm_layer=norm_layer,
                                      activation_layer=F.silu, key=keys[0]))
        self.features = nn.Sequential(seq)
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        self.classifier = nn.Sequential([nn.Dropout(p=dropout), nn.Linear(head_out, num_classes, key=keys[1])])

    def __call__(self, x, *, key=None):
        k1, k2 = jrandom.split(key, 2)
        x = self.avgpool(self.features(x, key=k1))
        return self.classifier(F.ravel(x), key=k2)


def _efficientnet(arch, inverted_residual_setting, dropout, last_channel, torch_weights, **kwargs) -> EfficientNet:
    model = EfficientNet(inverted_residual_setting, dropout, last_channel=last_channel, **kwargs)
    if torch_weights:
        model = load_torch_weights(model, torch_weights=torch_weights)
    return model


# (expand, kernel, stride, in, out, layers) rows; "F" rows are fused blocks
_V1_ROWS = [(1, 3, 1, 32, 16, 1), (6, 3, 2, 16, 24, 2), (6, 5, 2, 24, 40, 2), (6, 3, 2, 40, 80, 3),
            (6, 5, 1, 80, 112, 3), (6, 5, 2, 112, 192, 4), (6, 3, 1, 192, 320, 1)]
_V2_ROWS = {
    "efficientnet_v2_s": [("F", 1, 3, 1, 24, 24, 2), ("F", 4, 3, 2, 24, 48, 4), ("F", 4, 3, 2, 48, 64, 4),
                          ("M", 4, 3, 2, 64, 128, 6), ("M", 6, 3, 1, 128, 160, 9), ("M", 6, 3, 2, 160, 256, 15)],
    "efficientnet_v2_m": [("F", 1, 3, 1, 24, 24, 3), ("F", 4, 3, 2, 24, 48, 5), ("F", 4, 3, 2, 48, 80, 5),
                          ("M", 4, 3, 2, 80, 160, 7), ("M", 6, 3, 1, 160, 176, 14), ("M", 6, 3, 2, 176, 304, 18),
                          ("M", 6, 3, 1, 304, 512, 5)],
    "efficientnet_v2_l": [("F", 1, 3, 1, 32, 32, 4), ("F", 4, 3, 2, 32, 64, 7), ("F", 4, 3, 2, 64, 96, 7),
                          ("M", 4, 3, 2, 96, 192, 10), ("M", 6, 3, 1, 192, 224, 19), ("M", 6, 3, 2, 224, 384, 25),
                          ("M", 6, 3, 1, 384, 640, 7)],
}


def _efficientnet_conf(arch: str, **kwargs: Any) -> Tuple[Sequence[_MBConvConfigData], Optional[int]]:
    if arch.startswith("efficientnet_b"):
        wm, dm = kwargs.pop("width_mult"), kwargs.pop("depth_mult")
        return [_MBConvConfig(*row, width_mult=wm, depth_mult=dm) for row in _V1_ROWS], None
    for name, rows in _V2_ROWS.items():
        if arch.startswith(name):
            return [(_FusedMBConvConfig if kind == "F" else _MBConvConfig)(*row) for kind, *row in rows], 1280
    raise ValueError(f"Unsupported model type {arch}")


# arch -> (width_mult, depth_mult, dropout, BatchNorm kwargs)
_V1 = {"efficientnet_b0": (1.0, 1.0, 0.2, None), "efficientnet_b1": (1.0, 1.1, 0.2, None),
       "efficientnet_b2": (1.1, 1.2, 0.3, None), "efficientnet_b3": (1.2, 1.4, 0.3, None),
       "efficientnet_b4": (1.4, 1.8, 0.4, None),
       "efficientnet_b5": (1.6, 2.2, 0.4, dict(eps=0.001, momentum=0.01)),
       "efficientnet_b6": (1.8, 2.6, 0.5, dict(eps=0.001, momentum=0.01)),
       "efficientnet_b7": (2.0, 3.1, 0.5, dict(eps=0.001, momentum=0.01))}
_V2 = {"efficientnet_v2_s": 0.2, "efficientnet_v2_m": 0.3, "efficientnet_v2_l": 0.4}


def _make_v1(arch):
    wm, dm, dropout, bn_kw = _V1[arch]

    def ctor(torch_weights: str = None, **kwargs: Any) -> EfficientNet:
        setting, last = _efficientnet_conf(arch, width_mult=wm, depth_mult=dm)
        if bn_kw is not None:
            kwargs.setdefault("norm_layer", partial(nn.BatchNorm, **bn_kw))
        return _efficientnet(arch, setting, dropout, last, torch_weights, **kwargs)

    ctor.__name__ = ctor.__qualname__ = arch
    ctor.__doc__ = f"{arch} (reference: efficientnet.py:482-651). `torch_weights`: path/URL of a torchvision checkpoint."
    return ctor


def _make_v2(arch):
    dropout = _V2[arch]

    def ctor(torch_weights: str = None, **kwargs: Any) -> EfficientNet:
        setting, last = _efficientnet_conf(arch)
        kwargs.setdefault("norm_layer", partial(nn.BatchNorm, eps=1e-03))
        return _efficientnet(arch, setting, dropout, last, torch_weights, **kwargs)

    ctor.__name__ = ctor.__qualname__ = arch
    ctor.__doc__ = f"{arch} (reference: efficientnet.py:654-715)."
    return ctor


efficientnet_b0, efficientnet_b1, efficientnet_b2, efficientnet_b3 = (_make_v1(f"efficientnet_b{i}") for i in range(4))
efficientnet_b4, efficientnet_b5, efficientnet_b6, efficientnet_b7 = (_make_v1(f"efficientnet_b{i}") for i in range(4, 8))
efficientnet_v2_s, efficientnet_v2_m, efficientnet_v2_l = (_make_v2(f"efficientnet_v2_{s}") for s in "sml")
