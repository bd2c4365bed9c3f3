"""SqueezeNet 1.0 / 1.1 (reference: models/classification/squeezenet.py).

_Fire = squeeze 1x1 -> ReLU -> concat(expand 1x1 -> ReLU, expand 3x3 p1 -> ReLU); all convolutions carry a bias and
there is no normalisation. Max-pools are 3x3/2 with `use_ceil=True`. The classifier is convolutional:
Dropout -> conv1x1(512, classes) -> ReLU -> global average pool, then `jnp.ravel`.
Device lowering: both expand convolutions store straight into their channel slice of one Fire output buffer (no concat
pass); the ceil-mode pools are `eqxv_maxpool2d_ceil_nhwc_bf16` (partial windows clipped at the map edge).
"""
from typing import Any, Optional

from ... import functional as F
from ... import nn
from ... import random as jrandom
from ...utils import load_torch_weights


class _Fire(nn.Module):
    inplanes: int
    squeeze: nn.Conv2d
    squeeze_activation: nn.Lambda
    expand1x1: nn.Conv2d
    expand1x1_activation: nn.Lambda
    expand3x3: nn.Conv2d
    expand3x3_activation: nn.Lambda

    def __init__(self, inplanes: int, squeeze_planes: int, expand1x1_planes: int, expand3x3_planes: int,
                 key=None) -> None:
        keys = jrandom.split(key, 3)
        self.inplanes = inplanes
        self.squeeze = nn.Conv2d(inplanes, squeeze_planes, kernel_size=1, key=keys[0])
        self.squeeze_activation = nn.Lambda(F.relu)
        self.expand1x1 = nn.Conv2d(squeeze_planes, expand1x1_planes, kernel_size=1, key=keys[1])
        self.expand1x1_activation = nn.Lambda(F.relu)
        self.expand3x3 = nn.Conv2d(squeeze_planes, expand3x3_planes, kernel_size=3, padding=1, key=keys[2])
        self.expand3x3_activation = nn.Lambda(F.relu)

    def __call__(self, x, *, key=None):
        x = self.squeeze_activation(self.squeeze(x))
        return F.concat_channels([self.expand1x1_activation(self.expand1x1(x)),
                                  self.expand3x3_activation(self.expand3x3(x))])      # squeezenet.py:49-55


def _pool():
    return nn.MaxPool2d(kernel_size=3, stride=2, use_ceil=True)


class SqueezeNet(nn.Module):
    """`torchvision.models.squeezenet` as ported by the reference (squeezenet.py:58-135)."""

    features: nn.Sequential
    classifier: nn.Sequential

    def __init__(self, version: str = "1_0", num_classes: int = 1000, dropout: float = 0.5, *,
                 key: Optional[Any] = None) -> None:
        keys = jrandom.split(jrandom.PRNGKey(0) if key is None else key, 10)
        if version == "1_0":
            self.features = nn.Sequential([
                nn.Conv2d(3, 96, kernel_size=7, stride=2, key=keys[0]), nn.Lambda(F.relu), _pool(),
                _Fire(96, 16, 64, 64, key=keys[1]), _Fire(128, 16, 64, 64, key=keys[2]),
                _Fire(128, 32, 128, 128, key=keys[3]), _pool(),
                _Fire(256, 32, 128, 128, key=keys[4]), _Fire(256, 48, 192, 192, key=keys[5]),
                _Fire(384, 48, 192, 192, key=keys[6]), _Fire(384, 64, 256, 256, key=keys[7]), _pool(),
                _Fire(512, 64, 256, 256, key=keys[8]),
            ])
        elif version == "1_1":
            self.features = nn.Sequential([
                nn.Conv2d(3, 64, kernel_size=3, stride=2, key=keys[0]), nn.Lambda(F.relu), _pool(),
                _Fire(64, 16, 64, 64, key=keys[1]), _Fire(128, 16, 64, 64, key=keys[2]), _pool(),
                _Fire(128, 32, 128, 128, key=keys[3]), _Fire(256, 32, 128, 128, key=keys[4]), _pool(),
                _Fire(256, 48, 192, 192, key=keys[5]), _Fire(384, 48, 192, 192, key=keys[6]),
                _Fire(384, 64, 256, 256, key=keys[7]), _Fire(512, 64, 256, 256, key=keys[8]),
            ])
        else:
            # the reference leaves `features` unset for any other string and fails later with an AttributeError
            # (squeezenet.py:83-118); fail at construction instead
            raise ValueError(f"Unsupported SqueezeNet version {version}: 1_0 or 1_1 expected")
        final_conv = nn.Conv2d(512, num_classes, kernel_size=1, key=keys[9])
        self.classifier = nn.Sequential([nn.Dropout(p=dropout), final_conv, nn.Lambda(F.relu),
                                         nn.AdaptiveAvgPool2d((1, 1))])

    def __call__(self, x, *, key=None):
        x = self.features(x)
        x = self.classifier(x, key=key)
        return F.ravel(x)


def _squeezenet(version: str, torch_weights: str, **kwargs: Any) -> SqueezeNet:
    model = SqueezeNet(version, **kwargs)
    if torch_weights:
        model = load_torch_weights(model, torch_weights=torch_weights)
    return model


def squeezenet1_0(torch_weights: str = None, **kwargs: Any) -> SqueezeNet:
    """SqueezeNet 1.0 (arXiv 1602.07360); minimum input 21x21 (squeezenet.py:145-156)."""
    return _squeezenet("1_0", torch_weights, **kwargs)


def squeezenet1_1(torch_weights: str = None, **kwargs: Any) -> SqueezeNet:
    """SqueezeNet 1.1: 2.4x less computation than 1.0; minimum input 17x17 (squeezenet.py:159-172)."""
    return _squeezenet("1_1", torch_weights, **kwargs)
