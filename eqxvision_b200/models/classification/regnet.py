"""RegNet X / Y (reference: models/classification/regnet.py).

stem = 3x3/2 CNA(32); each stage = `depth` ResBottleneckBlocks: proj(x) + f(x) -> ReLU with
f = 1x1 CNA -> 3x3 grouped CNA (stride on the 3x3, groups = w_b // group_width) -> [SqueezeExcitation, squeeze width
= round(se_ratio * width_in)] -> 1x1 CNA (no activation); proj = 1x1/stride CNA when the shape changes, else Identity.
Device lowering: the 1x1 convs are tcgen05 GEMMs with BN / ReLU / the residual add in the epilogue; the grouped 3x3 runs
on the block-diagonal 64-channel layout when the geometry allows it, otherwise as a dense implicit GEMM over the
host-expanded block-diagonal filter (`_engine._emit_grouped`); SE = global pool -> two 1x1 GEMMs -> channel gate.
"""
import math
from functools import partial
from typing import Any, Callable, List, Optional, Tuple

from ... import functional as F
from ... import nn
from ... import random as jr
from ...layers import ConvNormActivation, SqueezeExcitation
from ...utils import _make_divisible, load_torch_weights


class SimpleStemIN(ConvNormActivation):
    """Simple stem for ImageNet: 3x3, BN, ReLU (regnet.py:16-37)."""

    def __init__(self, width_in: int, width_out: int, norm_layer: Optional[Callable],
                 activation_layer: Optional[Callable], *, key=None) -> None:
        super().__init__(width_in, width_out, kernel_size=3, stride=2, norm_layer=norm_layer,
                         activation_layer=activation_layer, key=key)


class BottleneckTransform(nn.Sequential):
    """Bottleneck transformation: 1x1, 3x3 [+SE], 1x1 (regnet.py:39-110)."""

    def __init__(self, width_in: int, width_out: int, stride: int, norm_layer: Optional[Callable],
                 activation_layer: Optional[Callable], group_width: int, bottleneck_multiplier: float,
                 se_ratio: Optional[float], *, key=None) -> None:
        keys = jr.split(key, 4)
        w_b = int(round(width_out * bottleneck_multiplier))
        g = w_b // group_width
        seq: List[nn.Module] = [
            ConvNormActivation(width_in, w_b, kernel_size=1, stride=1, norm_layer=norm_layer,
                               activation_layer=activation_layer, key=keys[0]),
            ConvNormActivation(w_b, w_b, kernel_size=3, stride=stride, groups=g, norm_layer=norm_layer,
                               activation_layer=activation_layer, key=keys[1]),
        ]
        if se_ratio:
            # the SE reduction ratio is defined with respect to the beginning of the block (regnet.py:84-86)
            seq.append(SqueezeExcitation(input_channels=w_b, squeeze_channels=int(round(se_ratio * width_in)),
                                         activation=activation_layer, key=keys[2]))
        seq.append(ConvNormActivation(w_b, width_out, kernel_size=1, stride=1, norm_layer=norm_layer,
                                      activation_layer=None, key=keys[3]))
        super().__init__(seq)


class ResBottleneckBlock(nn.Module):
    """Residual bottleneck block: x + F(x), F = bottleneck transform (regnet.py:113-167)."""

    proj: nn.Module
    f: nn.Module
    activation: Callable

    def __init__(self, width_in: int, width_out: int, stride: int, norm_layer: Optional[Callable],
                 activation_layer: Optional[Callable], group_width: int = 1, bottleneck_multiplier: float = 1.0,
                 se_ratio: Optional[float] = None, *, key=None) -> None:
        keys = jr.split(key, 2)
        self.proj = nn.Identity()
        if (width_in != width_out) or (stride != 1):
            self.proj = ConvNormActivation(width_in, width_out, kernel_size=1, stride=stride, norm_layer=norm_layer,
                                           activation_layer=None, key=keys[0])
        self.f = BottleneckTransform(width_in, width_out, stride, norm_layer, activation_layer, group_width,
                                     bottleneck_multiplier, se_ratio, key=keys[1])
        self.activation = activation_layer

    def __call__(self, x, *, key=None):
        keys = [None, None] if key is None else jr.split(key, 2)
        x = self.proj(x, key=keys[0]) + self.f(x, key=keys[1])
        return self.activation(x)


class AnyStage(nn.Sequential):
    """AnyNet stage: a sequence of blocks with the same output shape (regnet.py:170-203)."""

    def __init__(self, width_in: int, width_out: int, stride: int, depth: int, block_constructor, norm_layer: Callable,
                 activation_layer: Callable, group_width: int, bottleneck_multiplier: float,
                 se_ratio: Optional[float] = None, *, key=None) -> None:
        keys = jr.split(key, depth)
        super().__init__([
            block_constructor(width_in if i == 0 else width_out, width_out, stride if i == 0 else 1, norm_layer,
                              activation_layer, group_width, bottleneck_multiplier, se_ratio, key=keys[i])
            for i in range(depth)])


class BlockParams:
    """Per-stage settings (regnet.py:206-325)."""

    def __init__(self, depths: List[int], widths: List[int], group_widths: List[int],
                 bottleneck_multipliers: List[float], strides: List[int], se_ratio: Optional[float] = None) -> None:
        self.depths = depths
        self.widths = widths
        self.group_widths = group_widths
        self.bottleneck_multipliers = bottleneck_multipliers
        self.strides = strides
        self.se_ratio = se_ratio

    @classmethod
    def from_init_params(cls, depth: int, w_0: int, w_a: float, w_m: float, group_width: int,
                         bottleneck_multiplier: float = 1.0, se_ratio: Optional[float] = None) -> "BlockParams":
        """Quantised linear width progression in log space, then one stage per distinct width (regnet.py:222-297).
        The reference evaluates it in float32 (`jnp`); so does this (`numpy.float32`)."""
        import numpy as np

        QUANT, STRIDE = 8, 2
        if w_a < 0 or w_0 <= 0 or w_m <= 1 or w_0 % 8 != 0:
            raise ValueError("Invalid RegNet settings")
        f32 = np.float32
        widths_cont = np.arange(depth, dtype=f32) * f32(w_a) + f32(w_0)
        block_capacity = np.round(np.log(widths_cont / f32(w_0)) / f32(math.log(w_m)))
        block_widths = (np.round(f32(w_0) * np.power(f32(w_m), block_capacity) / f32(QUANT)) * QUANT) \
            .astype(np.int32).tolist()
        num_stages = len(set(block_widths))
        split_helper = zip(block_widths + [0], [0] + block_widths, block_widths + [0], [0] + block_widths)
        splits = [w != wp or r != rp for w, wp, r, rp in split_helper]
        stage_widths = [w for w, t in zip(block_widths, splits[:-1]) if t]
        stage_depths = np.diff(np.asarray([d for d, t in enumerate(splits) if t])).astype(np.int32).tolist()
        strides = [STRIDE] * num_stages
        bottleneck_multipliers = [bottleneck_multiplier] * num_stages
        group_widths = [group_width] * num_stages
        stage_widths, group_widths = cls._adjust_widths_groups_compatibilty(stage_widths, bottleneck_multipliers,
                                                                            group_widths)
        return cls(depths=stage_depths, widths=stage_widths, group_widths=group_widths,
                   bottleneck_multipliers=bottleneck_multipliers, strides=strides, se_ratio=se_ratio)

    def _get_expanded_params(self):
        return zip(self.widths, self.strides, self.depths, self.group_widths, self.bottleneck_multipliers)

    @staticmethod
    def _adjust_widths_groups_compatibilty(stage_widths: List[int], bottleneck_ratios: List[float],
                                           group_widths: List[int]) -> Tuple[List[int], List[int]]:
        widths = [int(w * b) for w, b in zip(stage_widths, bottleneck_ratios)]
        group_widths_min = [min(g, w_bot) for g, w_bot in zip(group_widths, widths)]
        ws_bot = [_make_divisible(w_bot, g) for w_bot, g in zip(widths, group_widths_min)]
        stage_widths = [int(w_bot / b) for w_bot, b in zip(ws_bot, bottleneck_ratios)]
        return stage_widths, group_widths_min


class RegNet(nn.Module):
    """`torchvision.models.regnet` layout as ported by the reference (regnet.py:328-430)."""

    stem: nn.Module
    trunk_output: nn.Sequential
    avgpool: nn.AdaptiveAvgPool2d
    fc: nn.Module

    def __init__(self, block_params: BlockParams, num_classes: int = 1000, stem_width: int = 32, stem_type=None,
                 block_type=None, norm_layer=None, activation: Optional[Callable] = None, *, key=None) -> None:
        stem_type = SimpleStemIN if stem_type is None else stem_type
        norm_layer = nn.BatchNorm if norm_layer is None else norm_layer
        block_type = ResBottleneckBlock if block_type is None else block_type
        activation = F.relu if activation is None else activation
        keys = jr.split(jr.PRNGKey(0) if key is None else key, 2)
        self.stem = stem_type(3, stem_width, norm_layer, activation, key=keys[0])
        current_width = stem_width
        blocks = []
        for width_out, stride, depth, group_width, bottleneck_multiplier in block_params._get_expanded_params():
            keys = jr.split(keys[1], 2)
            blocks.append(AnyStage(current_width, width_out, stride, depth, block_type, norm_layer, activation,
                                   group_width, bottleneck_multiplier, block_params.se_ratio, key=keys[0]))
            current_width = width_out
        self.trunk_output = nn.Sequential(blocks)
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(in_features=current_width, out_features=num_classes, key=keys[1])

    def __call__(self, x, *, key=None):
        keys = [None, None] if key is None else jr.split(key, 2)
        x = self.stem(x, key=keys[0])
        x = self.trunk_output(x, key=keys[1])
        x = self.avgpool(x)
        return self.fc(F.ravel(x))


def _regnet(arch: str, block_params: BlockParams, torch_weights: str, **kwargs: Any) -> RegNet:
    norm_layer = kwargs.pop("norm_layer", partial(nn.BatchNorm, eps=1e-05, momentum=0.1))
    model = RegNet(block_params, norm_layer=norm_layer, **kwargs)
    if torch_weights:
        model = load_torch_weights(model, torch_weights=torch_weights)
    return model


# (depth, w_0, w_a, w_m, group_width, se_ratio): regnet.py:452-676
_REGNETS = {
    "regnet_y_400mf": (16, 48, 27.89, 2.09, 8, 0.25), "regnet_y_800mf": (14, 56, 38.84, 2.4, 16, 0.25),
    "regnet_y_1_6gf": (27, 48, 20.71, 2.65, 24, 0.25), "regnet_y_3_2gf": (21, 80, 42.63, 2.66, 24, 0.25),
    "regnet_y_8gf": (17, 192, 76.82, 2.19, 56, 0.25), "regnet_y_16gf": (18, 200, 106.23, 2.48, 112, 0.25),
    "regnet_y_32gf": (20, 232, 115.89, 2.53, 232, 0.25), "regnet_y_128gf": (27, 456, 160.83, 2.52, 264, 0.25),
    "regnet_x_400mf": (22, 24, 24.48, 2.54, 16, None), "regnet_x_800mf": (16, 56, 35.73, 2.28, 16, None),
    "regnet_x_1_6gf": (18, 80, 34.01, 2.25, 24, None), "regnet_x_3_2gf": (25, 88, 26.31, 2.25, 48, None),
    "regnet_x_8gf": (23, 80, 49.56, 2.88, 120, None), "regnet_x_16gf": (22, 216, 55.59, 2.1, 128, None),
    "regnet_x_32gf": (23, 320, 69.86, 2.0, 168, None),
}


def _ctor(arch: str):
    depth, w_0, w_a, w_m, gw, se = _REGNETS[arch]

    def build(torch_weights: str = None, **kwargs: Any) -> RegNet:
        params = BlockParams.from_init_params(depth=depth, w_0=w_0, w_a=w_a, w_m=w_m, group_width=gw, se_ratio=se)
        return _regnet(arch, params, torch_weights, **kwargs)

    build.__name__ = build.__qualname__ = arch
    build.__doc__ = (f"RegNet{arch[7].upper()}_{arch[9:].upper()} from 'Designing Network Design Spaces' "
                     "(arXiv 2003.13678); `torch_weights`: path or URL of the PyTorch checkpoint.")
    return build


regnet_y_400mf = _ctor("regnet_y_400mf")
regnet_y_800mf = _ctor("regnet_y_800mf")
regnet_y_1_6gf = _ctor("regnet_y_1_6gf")
regnet_y_3_2gf = _ctor("regnet_y_3_2gf")
regnet_y_8gf = _ctor("regnet_y_8gf")
regnet_y_16gf = _ctor("regnet_y_16gf")
regnet_y_32gf = _ctor("regnet_y_32gf")
regnet_y_128gf = _ctor("regnet_y_128gf")
regnet_x_400mf = _ctor("regnet_x_400mf")
regnet_x_800mf = _ctor("regnet_x_800mf")
regnet_x_1_6gf = _ctor("regnet_x_1_6gf")
regnet_x_3_2gf = _ctor("regnet_x_3_2gf")
regnet_x_8gf = _ctor("regnet_x_8gf")
regnet_x_16gf = _ctor("regnet_x_16gf")
regnet_x_32gf = _ctor("regnet_x_32gf")
