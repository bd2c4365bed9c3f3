"""AlexNet (reference: models/classification/alexnet.py) - BASELINE config 0, the README example.

features   = conv11x11/4 p2 -> ReLU -> maxpool3/2 -> conv5x5 p2 -> ReLU -> maxpool3/2 -> 3 x (conv3x3 p1 -> ReLU)
             -> maxpool3/2; every convolution carries a bias, there is no normalisation layer
classifier = Dropout -> Linear(9216, 4096) -> ReLU -> Dropout -> Linear(4096, 4096) -> ReLU -> Linear
Lowering: the 11x11 stride-4 first layer is wider than the 8-tap rows of the first-layer kernel, so it runs
on the generic implicit GEMM over the 8-channel NHWC image (121 taps, TMA element strides = the conv stride);
`jnp.ravel` of the (256,6,6) map is C,H,W order, handled by the column-permuted first classifier weight
(`_engine._emit_Linear`), as for VGG.
"""
from typing import Any, Optional

from ... import functional as F
from ... import nn
from ... import random as jrandom
from ...utils import load_torch_weights


class AlexNet(nn.Module):
    """`torchvision.models.alexnet` as ported by the reference (alexnet.py:14-89)."""

    features: nn.Sequential
    avgpool: nn.AdaptiveAvgPool2d
    classifier: nn.Sequential

    def __init__(self, num_classes: int = 1000, dropout: float = 0.5, *, key: Optional[Any] = None) -> None:
        keys = jrandom.split(jrandom.PRNGKey(0) if key is None else key, 8)
        relu = lambda: nn.Lambda(F.relu)  # noqa: E731
        self.features = nn.Sequential([
            nn.Conv2d(3, 64, kernel_size=11, stride=4, padding=2, key=keys[0]), relu(),
            nn.MaxPool2d(kernel_size=3, stride=2),
            nn.Conv2d(64, 192, kernel_size=5, padding=2, key=keys[1]), relu(),
            nn.MaxPool2d(kernel_size=3, stride=2),
            nn.Conv2d(192, 384, kernel_size=3, padding=1, key=keys[2]), relu(),
            nn.Conv2d(384, 256, kernel_size=3, padding=1, key=keys[3]), relu(),
            nn.Conv2d(256, 256, kernel_size=3, padding=1, key=keys[4]), relu(),
            nn.MaxPool2d(kernel_size=3, stride=2),
        ])
        self.avgpool = nn.AdaptiveAvgPool2d((6, 6))
        self.classifier = nn.Sequential([
            nn.Dropout(p=dropout),
            nn.Linear(256 * 6 * 6, 4096, key=keys[5]), relu(),
            nn.Dropout(p=dropout),
            nn.Linear(4096, 4096, key=keys[6]), relu(),
            nn.Linear(4096, num_classes, key=keys[7]),
        ])

    def __call__(self, x, *, key=None):
        if key is None:  # alexnet.py:78-79
            raise RuntimeError("The model requires a PRNGKey.")
        k1, k2 = jrandom.split(key, 2)
        x = self.avgpool(self.features(x, key=k1))
        return self.classifier(F.ravel(x), key=k2)

    def __iter__(self):  # alexnet.py:87-89
        for attr, value in self.__dict__.items():
            yield attr, value


def alexnet(torch_weights: str = None, **kwargs: Any) -> AlexNet:
    """AlexNet ("One weird trick ...", arXiv 1404.5997); minimum input 63x63 (alexnet.py:92-103)."""
    model = AlexNet(**kwargs)
    if torch_weights:
        model = load_torch_weights(model, torch_weights=torch_weights)
    return model
