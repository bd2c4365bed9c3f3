"""ResNet family (reference: models/classification/resnet.py).

Module tree and field order mirror the reference (they define the positional weight-loading
contract); the forward bodies are written against the tracer so that each
bottleneck lowers to three fused implicit-GEMM launches:

    conv1x1+BN+ReLU -> conv3x3(stride, dilation)+BN+ReLU -> conv1x1+BN (+identity) +ReLU

(`out += identity; relu(out)` of resnet.py:159-160 becomes the residual epilogue of the last GEMM).
"""
from typing import Any, Callable, List, Optional, Sequence, Type, Union

from ... import functional as F
from ... import nn
from ... import random as jrandom
from ...utils import load_torch_weights


def _conv3x3(cin, cout, stride=1, groups=1, dilation=1, key=None):
    # padding == dilation keeps the spatial size at stride 1 (resnet.py:15-27)
    return nn.Conv2d(cin, cout, kernel_size=3, stride=stride, padding=dilation, groups=groups,
                     use_bias=False, dilation=dilation, key=key)


def _conv1x1(cin, cout, stride=1, key=None):
    return nn.Conv2d(cin, cout, kernel_size=1, stride=stride, use_bias=False, key=key)


class _ResNetBasicBlock(nn.Module):
    expansion: int
    conv1: nn.Module
    bn1: nn.Module
    relu: Callable
    conv2: nn.Module
    bn2: nn.Module
    downsample: nn.Module
    stride: int

    def __init__(self, inplanes, planes, stride=1, downsample=None, groups=1, base_width=64, dilation=1,
                 norm_layer=None, key=None):
        norm_layer = nn.BatchNorm if norm_layer is None else norm_layer
        if groups != 1 or base_width != 64:
            raise ValueError("BasicBlock only supports groups=1 and base_width=64")
        if dilation > 1:
            raise NotImplementedError("Dilation > 1 not supported in BasicBlock")
        k1, k2 = jrandom.split(key, 2)
        self.expansion = 1
        self.conv1 = _conv3x3(inplanes, planes, stride, key=k1)
        self.bn1 = norm_layer(planes, axis_name="batch")
        self.relu = F.relu
        self.conv2 = _conv3x3(planes, planes, key=k2)
        self.bn2 = norm_layer(planes, axis_name="batch")
        self.downsample = downsample if downsample else nn.Identity()
        self.stride = stride

    def __call__(self, x, *, key=None):
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        return self.relu(y + self.downsample(x))


class _ResNetBottleneck(nn.Module):
    # stride sits on the 3x3 (torchvision "v1.5"), resnet.py:96-100,133
    expansion: int
    conv1: nn.Module
    bn1: nn.Module
    conv2: nn.Module
    bn2: nn.Module
    conv3: nn.Module
    bn3: nn.Module
    relu: Callable
    downsample: nn.Module
    stride: int

    def __init__(self, inplanes, planes, stride=1, downsample=None, groups=1, base_width=64, dilation=1,
                 norm_layer=None, key=None):
        norm_layer = nn.BatchNorm if norm_layer is None else norm_layer
        self.expansion = 4
        k1, k2, k3 = jrandom.split(key, 3)
        width = int(planes * (base_width / 64.0)) * groups
        self.conv1 = _conv1x1(inplanes, width, key=k1)
        self.bn1 = norm_layer(width, axis_name="batch")
        self.conv2 = _conv3x3(width, width, stride, groups, dilation, key=k2)
        self.bn2 = norm_layer(width, axis_name="batch")
        self.conv3 = _conv1x1(width, planes * self.expansion, key=k3)
        self.bn3 = norm_layer(planes * self.expansion, axis_name="batch")
        self.relu = F.relu
        self.downsample = downsample if downsample else nn.Identity()
        self.stride = stride

    def __call__(self, x, *, key=None):
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.relu(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        return self.relu(y + self.downsample(x))


EXPANSIONS = {_ResNetBasicBlock: 1, _ResNetBottleneck: 4}


class ResNet(nn.Module):
    """`torchvision.models.resnet` layout, as ported by the reference (resnet.py:168-358)."""

    inplanes: int
    dilation: int
    groups: Sequence[int]
    base_width: int
    conv1: nn.Module
    bn1: nn.Module
    relu: Callable
    maxpool: nn.Module
    layer1: nn.Module
    layer2: nn.Module
    layer3: nn.Module
    layer4: nn.Module
    avgpool: nn.Module
    fc: nn.Module

    def __init__(
        self,
        block: Type[Union["_ResNetBasicBlock", "_ResNetBottleneck"]],
        layers: List[int],
        num_classes: int = 1000,
        groups: int = 1,
        width_per_group: int = 64,
        replace_stride_with_dilation: List[bool] = None,
        norm_layer: Any = None,
        *,
        key=None,
    ):
        if not norm_layer:
            norm_layer = nn.BatchNorm
        if norm_layer != nn.BatchNorm:
            raise NotImplementedError(
                f"{type(norm_layer)} is not currently supported. Use `eqxvision_b200.nn.BatchNorm` instead."
            )
        key = jrandom.PRNGKey(0) if key is None else key
        keys = jrandom.split(key, 6)
        self.inplanes = 64
        self.dilation = 1
        if replace_stride_with_dilation is None:
            replace_stride_with_dilation = [False, False, False]
        if len(replace_stride_with_dilation) != 3:
            raise ValueError(
                "replace_stride_with_dilation should be None "
                "or a 3-element tuple, got {}".format(replace_stride_with_dilation)
            )
        self.groups = groups
        self.base_width = width_per_group
        self.conv1 = nn.Conv2d(3, self.inplanes, kernel_size=7, stride=2, padding=3, use_bias=False, key=keys[0])
        self.bn1 = norm_layer(input_size=self.inplanes, axis_name="batch")
        self.relu = F.relu
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1)
        self.layer1 = self._make_layer(block, 64, layers[0], norm_layer, key=keys[1])
        stage_args = zip((128, 256, 512), layers[1:], replace_stride_with_dilation, keys[2:5])
        stages = [self._make_layer(block, planes, depth, norm_layer, stride=2, dilate=dilate, key=k)
                  for planes, depth, dilate, k in stage_args]
        self.layer2, self.layer3, self.layer4 = stages
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(512 * EXPANSIONS[block], num_classes, key=keys[5])

    def _make_layer(self, block, planes, blocks, norm_layer, stride=1, dilate=False, key=None):
        keys = jrandom.split(key, blocks + 1)
        out_planes = planes * EXPANSIONS[block]
        first_dilation = self.dilation  # the first block of a dilated stage keeps the old dilation
        if dilate:
            self.dilation *= stride
            stride = 1
        shortcut = None
        if stride != 1 or self.inplanes != out_planes:
            shortcut = nn.Sequential([
                _conv1x1(self.inplanes, out_planes, stride, key=keys[0]),
                norm_layer(out_planes, axis_name="batch"),
            ])
        stack = [block(self.inplanes, planes, stride, shortcut, self.groups, self.base_width,
                       first_dilation, norm_layer, key=keys[1])]
        self.inplanes = out_planes
        for i in range(1, blocks):
            stack.append(block(self.inplanes, planes, groups=self.groups, base_width=self.base_width,
                               dilation=self.dilation, norm_layer=norm_layer, key=keys[i + 1]))
        return nn.Sequential(stack)

    def __call__(self, x, *, key=None):
        if key is None:
            raise RuntimeError("The model requires a PRNGKey.")
        keys = jrandom.split(key, 6)
        x = self.maxpool(self.relu(self.bn1(self.conv1(x, key=keys[0]))))
        for stage, k in zip((self.layer1, self.layer2, self.layer3, self.layer4), keys[1:5]):
            x = stage(x, key=k)
        x = F.ravel(self.avgpool(x))
        return self.fc(x, key=keys[5])


def _resnet(block, layers, torch_weights, **kwargs):
    model = ResNet(block, layers, **kwargs)
    if torch_weights:
        model = load_torch_weights(model, torch_weights=torch_weights)
    return model


def resnet18(torch_weights: str = None, **kwargs: Any) -> ResNet:
    """ResNet-18. `torch_weights`: path or URL of a torchvision checkpoint."""
    return _resnet(_ResNetBasicBlock, [2, 2, 2, 2], torch_weights, **kwargs)


def resnet34(torch_weights: str = None, **kwargs: Any) -> ResNet:
    return _resnet(_ResNetBasicBlock, [3, 4, 6, 3], torch_weights, **kwargs)


def resnet50(torch_weights: str = None, **kwargs: Any) -> ResNet:
    return _resnet(_ResNetBottleneck, [3, 4, 6, 3], torch_weights, **kwargs)


def resnet101(torch_weights: str = None, **kwargs: Any) -> ResNet:
    return _resnet(_ResNetBottleneck, [3, 4, 23, 3], torch_weights, **kwargs)


def resnet152(torch_weights: str = None, **kwargs: Any) -> ResNet:
    return _resnet(_ResNetBottleneck, [3, 8, 36, 3], torch_weights, **kwargs)


def resnext50_32x4d(torch_weights: str = None, **kwargs: Any) -> ResNet:
    kwargs["groups"] = 32
    kwargs["width_per_group"] = 4
    return _resnet(_ResNetBottleneck, [3, 4, 6, 3], torch_weights, **kwargs)


def resnext101_32x8d(torch_weights: str = None, **kwargs: Any) -> ResNet:
    kwargs["groups"] = 32
    kwargs["width_per_group"] = 8
    return _resnet(_ResNetBottleneck, [3, 4, 23, 3], torch_weights, **kwargs)


def wide_resnet50_2(torch_weights: str = None, **kwargs: Any) -> ResNet:
    kwargs["width_per_group"] = 64 * 2
    return _resnet(_ResNetBottleneck, [3, 4, 6, 3], torch_weights, **kwargs)


def wide_resnet101_2(torch_weights: str = None, **kwargs: Any) -> ResNet:
    kwargs["width_per_group"] = 64 * 2
    return _resnet(_ResNetBottleneck, [3, 4, 23, 3], torch_weights, **kwargs)
