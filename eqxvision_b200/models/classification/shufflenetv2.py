"""ShuffleNetV2 x0.5 / x1.0 / x1.5 / x2.0 (reference: models/classification/shufflenetv2.py).

_InvertedResidual, stride 1: split the channels in two, out = concat(x1, branch2(x2)); stride 2:
out = concat(branch1(x), branch2(x)); then `_channel_shuffle(out, 2)`.
branch2 = 1x1 -> BN -> ReLU -> depthwise 3x3 (stride) -> BN -> 1x1 -> BN -> ReLU; branch1 = depthwise 3x3/2 -> BN -> 1x1
-> BN -> ReLU.
Device lowering: split and shuffle are lazy channel views (`_trace.ChannelView`); every 1x1 convolution that reads
one absorbs the gather into its filter's input channels, so the shuffle itself never runs. Only the pass-through half
x1 and the inputs of the stride-2 depthwise convs are materialised, as exact 0/1-matrix GEMMs that store straight
into their slice of the unit's output buffer. Branch widths that are not multiples of 8 channels (x1.0: 58, x2.0: 122)
get 16-byte aligned slots in that buffer.
"""
from typing import Any, List, Optional

from ... import functional as F
from ... import nn
from ... import random as jrandom
from ...utils import load_torch_weights


def _channel_shuffle(x, groups: int):
    return F.channel_shuffle(x, groups)


class _InvertedResidual(nn.Module):
    stride: int
    branch1: nn.Sequential
    branch2: nn.Sequential

    def __init__(self, inp: int, oup: int, stride: int, *, key=None) -> None:
        keys = jrandom.split(key, 5)
        if not (1 <= stride <= 3):
            raise ValueError("illegal stride value")
        branch_features = oup // 2
        assert (stride != 1) or (inp == branch_features << 1)
        self.stride = stride
        if stride > 1:
            self.branch1 = nn.Sequential([
                self.depthwise_conv(inp, inp, kernel_size=3, stride=self.stride, padding=1, key=keys[0]),
                nn.BatchNorm(inp, axis_name="batch"),
                nn.Conv2d(inp, branch_features, kernel_size=1, stride=1, padding=0, use_bias=False, key=keys[1]),
                nn.BatchNorm(branch_features, axis_name="batch"),
                nn.Lambda(F.relu),
            ])
        else:
            self.branch1 = nn.Sequential([nn.Identity()])   # never called (shufflenetv2.py:70 stores the class itself)
        self.branch2 = nn.Sequential([
            nn.Conv2d(inp if (self.stride > 1) else branch_features, branch_features, kernel_size=1, stride=1,
                      padding=0, use_bias=False, key=keys[2]),
            nn.BatchNorm(branch_features, axis_name="batch"),
            nn.Lambda(F.relu),
            self.depthwise_conv(branch_features, branch_features, kernel_size=3, stride=self.stride, padding=1,
                                key=keys[3]),
            nn.BatchNorm(branch_features, axis_name="batch"),
            nn.Conv2d(branch_features, branch_features, kernel_size=1, stride=1, padding=0, use_bias=False,
                      key=keys[4]),
            nn.BatchNorm(branch_features, axis_name="batch"),
            nn.Lambda(F.relu),
        ])

    @staticmethod
    def depthwise_conv(i: int, o: int, kernel_size: int, stride: int = 1, padding: int = 0, bias: bool = False,
                       key=None) -> nn.Conv2d:
        return nn.Conv2d(i, o, kernel_size, stride, padding, use_bias=bias, groups=i, key=key)

    def __call__(self, x, *, key=None):
        if self.stride == 1:
            x1, x2 = F.split_channels(x, 2)
            out = F.concat_channels((x1, self.branch2(x2)))
        else:
            out = F.concat_channels((self.branch1(x), self.branch2(x)))
        return _channel_shuffle(out, 2)


class ShuffleNetV2(nn.Module):
    """`torchvision.models.shufflenetv2` as ported by the reference (shufflenetv2.py:143-262)."""

    conv1: nn.Sequential
    maxpool: nn.MaxPool2d
    stage2: nn.Sequential
    stage3: nn.Sequential
    stage4: nn.Sequential
    conv5: nn.Sequential
    pool: nn.AdaptiveAvgPool2d
    fc: nn.Linear

    def __init__(self, stages_repeats: List[int], stages_out_channels: List[int], num_classes: int = 1000,
                 inverted_residual=_InvertedResidual, *, key: Optional[Any] = None) -> None:
        keys = jrandom.split(jrandom.PRNGKey(0) if key is None else key, 2)
        if len(stages_repeats) != 3:
            raise ValueError("expected stages_repeats as list of 3 positive ints")
        if len(stages_out_channels) != 5:
            raise ValueError("expected stages_out_channels as list of 5 positive ints")
        input_channels = 3
        output_channels = stages_out_channels[0]
        self.conv1 = nn.Sequential([nn.Conv2d(input_channels, output_channels, 3, 2, 1, use_bias=False, key=keys[0]),
                                    nn.BatchNorm(output_channels, axis_name="batch"), nn.Lambda(F.relu)])
        input_channels = output_channels
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        for name, repeats, output_channels in zip(["stage2", "stage3", "stage4"], stages_repeats,
                                                  stages_out_channels[1:]):
            keys = jrandom.split(keys[1], 2)
            seq = [inverted_residual(input_channels, output_channels, 2, key=keys[0])]
            for _ in range(repeats - 1):
                keys = jrandom.split(keys[1], 2)
                seq.append(inverted_residual(output_channels, output_channels, 1, key=keys[0]))
            setattr(self, name, nn.Sequential(seq))
            input_channels = output_channels
        keys = jrandom.split(keys[1], 2)
        output_channels = stages_out_channels[-1]
        self.conv5 = nn.Sequential([nn.Conv2d(input_channels, output_channels, 1, 1, 0, use_bias=False, key=keys[0]),
                                    nn.BatchNorm(output_channels, axis_name="batch"), nn.Lambda(F.relu)])
        self.pool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(output_channels, num_classes, key=keys[1])

    def __call__(self, x, *, key=None):
        keys = [None] * 5 if key is None else jrandom.split(key, 5)
        x = self.conv1(x, key=keys[0])
        x = self.maxpool(x)
        x = self.stage2(x, key=keys[1])
        x = self.stage3(x, key=keys[2])
        x = self.stage4(x, key=keys[3])
        x = self.conv5(x, key=keys[4])
        return self.fc(F.ravel(self.pool(x)))


def _shufflenetv2(*args: Any, **kwargs: Any) -> ShuffleNetV2:
    return ShuffleNetV2(*args, **kwargs)


def _make(name: str, channels: List[int], mult: str):
    def build(torch_weights: str = None, **kwargs: Any) -> ShuffleNetV2:
        model = _shufflenetv2([4, 8, 4], channels, **kwargs)
        if torch_weights:
            model = load_torch_weights(model, torch_weights=torch_weights)
        return model

    build.__name__ = build.__qualname__ = name
    build.__doc__ = (f"ShuffleNetV2 with {mult} output channels (arXiv 1807.11164); `torch_weights`: path or URL of "
                     "the PyTorch checkpoint (shufflenetv2.py:270-321).")
    return build


shufflenet_v2_x0_5 = _make("shufflenet_v2_x0_5", [24, 48, 96, 192, 1024], "0.5x")
shufflenet_v2_x1_0 = _make("shufflenet_v2_x1_0", [24, 116, 232, 464, 1024], "1.0x")
shufflenet_v2_x1_5 = _make("shufflenet_v2_x1_5", [24, 176, 352, 704, 1024], "1.5x")
shufflenet_v2_x2_0 = _make("shufflenet_v2_x2_0", [24, 244, 488, 976, 2048], "2.0x")
