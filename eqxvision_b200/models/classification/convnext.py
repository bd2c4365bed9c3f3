"""ConvNeXt tiny / small / base / large (reference: models/classification/convnext.py).

CNBlock: x + layer_scale * [depthwise 7x7 p3 (bias) -> LayerNorm2d -> Linear2d(C, 4C) -> GELU -> Linear2d(4C, C)](x),
stochastic depth is the identity in inference. Stem = conv4x4/4 (bias) -> LayerNorm2d; between stages
LayerNorm2d -> conv2x2/2; head = global pool -> LayerNorm2d -> ravel -> Linear.
Reference quirks kept (parity is with the reference): GELU is `jnn.gelu`, i.e. the tanh approximation
(convnext.py:52); the block's LayerNorm2d is built with the equinox default eps 1e-5 because `CNBlock` defaults
`norm_layer=LayerNorm2d` (convnext.py:24,33-34), while stem / downsample / head norms use 1e-6 (convnext.py:120).
Device lowering: channels-last, so LayerNorm2d / Linear2d need no transposes; the depthwise 7x7 is the channels-last
stencil, the two Linear2d are tcgen05 GEMMs (GELU in the first epilogue); `layer_scale` is folded into the rows of the
second GEMM's weight on the host and the residual `+ x` is that GEMM's epilogue.
"""
import os
from functools import partial
from typing import Any, Callable, List, Optional, Sequence

import torch

from ... import functional as F
from ... import nn
from ... import random as jrandom
from ...layers import ConvNormActivation, DropPath, LayerNorm2d, Linear2d
from ...utils import CLASSIFICATION_URLS, load_torch_weights


class CNBlock(nn.Module):
    layer_scale: torch.Tensor
    block: nn.Sequential
    stochastic_depth: DropPath

    def __init__(self, dim, layer_scale: float, stochastic_depth_prob: float,
                 norm_layer: Optional[Callable[..., nn.Module]] = LayerNorm2d, *, key=None) -> None:
        if norm_layer is None:
            norm_layer = partial(nn.LayerNorm, eps=1e-6)
        keys = jrandom.split(key, 4)
        self.block = nn.Sequential([
            nn.Conv2d(dim, dim, kernel_size=7, padding=3, groups=dim, use_bias=True, key=keys[0]),
            norm_layer(dim),
            Linear2d(in_features=dim, out_features=4 * dim, use_bias=True, key=keys[1]),
            nn.Lambda(F.gelu),
            Linear2d(in_features=4 * dim, out_features=dim, use_bias=True, key=keys[2]),
        ])
        self.layer_scale = torch.ones((dim, 1, 1), dtype=torch.float32) * layer_scale
        self.stochastic_depth = DropPath(p=stochastic_depth_prob, mode="local")

    def __call__(self, x, *, key=None):
        keys = [None, None] if key is None else jrandom.split(key, 2)
        result = self.layer_scale * self.block(x, key=keys[0])
        result = self.stochastic_depth(result, key=keys[1])
        result += x
        return result


class _CNBlockConfig:
    """one stage of Section 3 of the ConvNeXt paper (convnext.py:69-88)"""

    def __init__(self, input_channels: int, out_channels: Optional[int], num_layers: int) -> None:
        self.input_channels = input_channels
        self.out_channels = out_channels
        self.num_layers = num_layers

    def __repr__(self) -> str:
        return (f"{self.__class__.__name__}(input_channels={self.input_channels}, out_channels={self.out_channels}, "
                f"num_layers={self.num_layers})")


class ConvNeXt(nn.Module):
    """`torchvision.models.convnext` as ported by the reference (convnext.py:91-204)."""

    features: nn.Sequential
    avgpool: nn.AdaptiveAvgPool2d
    classifier: nn.Sequential

    def __init__(self, block_setting: Sequence["_CNBlockConfig"], stochastic_depth_prob: float = 0.0,
                 layer_scale: float = 1e-6, num_classes: int = 1000, block=None, norm_layer=None, *, key=None) -> None:
        if not block_setting:
            raise ValueError("The block_setting should not be empty")
        elif not (isinstance(block_setting, Sequence) and all(isinstance(s, _CNBlockConfig) for s in block_setting)):
            raise TypeError("The block_setting should be List[CNBlockConfig]")
        keys = jrandom.split(jrandom.PRNGKey(0) if key is None else key, 2)
        block = CNBlock if block is None else block
        norm_layer = partial(LayerNorm2d, eps=1e-6) if norm_layer is None else norm_layer
        layers: List[nn.Module] = [
            ConvNormActivation(in_channels=3, out_channels=block_setting[0].input_channels, kernel_size=4, stride=4,
                               padding=0, norm_layer=norm_layer, activation_layer=None, use_bias=True, key=keys[0])]
        total_stage_blocks = sum(cnf.num_layers for cnf in block_setting)
        stage_block_id = 0
        for cnf in block_setting:
            stage: List[nn.Module] = []
            for _ in range(cnf.num_layers):
                keys = jrandom.split(keys[1], 2)
                sd_prob = stochastic_depth_prob * stage_block_id / (total_stage_blocks - 1.0)
                stage.append(block(cnf.input_channels, layer_scale, sd_prob, key=keys[0]))
                stage_block_id += 1
            layers.append(nn.Sequential(stage))
            if cnf.out_channels is not None:
                keys = jrandom.split(keys[1], 2)
                layers.append(nn.Sequential([
                    norm_layer(cnf.input_channels),
                    nn.Conv2d(cnf.input_channels, cnf.out_channels, kernel_size=2, stride=2, key=keys[0])]))
        self.features = nn.Sequential(layers)
        self.avgpool = nn.AdaptiveAvgPool2d(1)
        lastblock = block_setting[-1]
        last_out = lastblock.out_channels if lastblock.out_channels is not None else lastblock.input_channels
        self.classifier = nn.Sequential([norm_layer(last_out), nn.Lambda(F.ravel),
                                         nn.Linear(last_out, num_classes, key=keys[1])])

    def __call__(self, x, *, key=None):
        x = self.features(x, key=key)
        x = self.avgpool(x)
        return self.classifier(x)


def _convnext(arch: str, block_setting: List[_CNBlockConfig], stochastic_depth_prob: float, torch_weights: str,
              **kwargs: Any) -> ConvNeXt:
    model = ConvNeXt(block_setting, stochastic_depth_prob=stochastic_depth_prob, **kwargs)
    if torch_weights:
        if arch not in CLASSIFICATION_URLS:
            raise ValueError(f"No checkpoint is available for model type {arch}")
        # the reference ignores the value and always fetches CLASSIFICATION_URLS[arch] (convnext.py:218-221); a path
        # that exists on disk is honoured here so that the model can be loaded without network access
        local = isinstance(torch_weights, (str, os.PathLike)) and os.path.exists(str(torch_weights))
        model = load_torch_weights(model, torch_weights=str(torch_weights) if local else CLASSIFICATION_URLS[arch])
    return model


_SETTINGS = {  # (stage (in, out, depth) list, default stochastic depth): convnext.py:233-324
    "convnext_tiny": ([(96, 192, 3), (192, 384, 3), (384, 768, 9), (768, None, 3)], 0.1),
    "convnext_small": ([(96, 192, 3), (192, 384, 3), (384, 768, 27), (768, None, 3)], 0.4),
    "convnext_base": ([(128, 256, 3), (256, 512, 3), (512, 1024, 27), (1024, None, 3)], 0.5),
    "convnext_large": ([(192, 384, 3), (384, 768, 3), (768, 1536, 27), (1536, None, 3)], 0.5),
}


def _build(arch: str, torch_weights, kwargs) -> ConvNeXt:
    stages, sd = _SETTINGS[arch]
    sd = kwargs.pop("stochastic_depth_prob", sd)
    return _convnext(arch, [_CNBlockConfig(*s) for s in stages], sd, torch_weights, **kwargs)


def convnext_tiny(*, torch_weights: str = None, **kwargs: Any) -> ConvNeXt:
    """ConvNeXt Tiny ("A ConvNet for the 2020s", arXiv 2201.03545)."""
    return _build("convnext_tiny", torch_weights, kwargs)


def convnext_small(*, torch_weights: str = None, **kwargs: Any) -> ConvNeXt:
    """ConvNeXt Small."""
    return _build("convnext_small", torch_weights, kwargs)


def convnext_base(*, torch_weights: str = None, **kwargs: Any) -> ConvNeXt:
    """ConvNeXt Base."""
    return _build("convnext_base", torch_weights, kwargs)


def convnext_large(*, torch_weights: str = None, **kwargs: Any) -> ConvNeXt:
    """ConvNeXt Large."""
    return _build("convnext_large", torch_weights, kwargs)
