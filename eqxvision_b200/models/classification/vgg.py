"""VGG 11/13/16/19 with and without BatchNorm (reference: models/classification/vgg.py).

features   = [conv3x3 pad1 WITH bias, (BatchNorm), ReLU]* interleaved with 2x2/2 max-pools
classifier = Linear -> Dropout -> Linear -> ReLU -> Dropout -> Linear
Reference quirk kept on purpose: there is NO ReLU after the first Linear
(vgg.py:97-106), unlike torchvision - the reference's own test only compares `.features`.
`jnp.ravel` flattens the (512,7,7) map in C,H,W order; the device buffer is H,W,C, so the first
classifier GEMM runs on a column-permuted copy of the weight (`_engine._emit_Linear`).
"""
from typing import Any, Dict, List, Optional, Union, cast

from ... import functional as F
from ... import nn
from ... import random as jrandom
from ...utils import load_torch_weights

_cfgs: Dict[str, List[Union[str, int]]] = {
    "A": [64, "M", 128, "M", 256, 256, "M", 512, 512, "M", 512, 512, "M"],
    "B": [64, 64, "M", 128, 128, "M", 256, 256, "M", 512, 512, "M", 512, 512, "M"],
    "D": [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M"],
    "E": [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512, "M"],
}


class VGG(nn.Module):
    """`torchvision.models.vgg` layout as ported by the reference (vgg.py:64-119)."""

    features: nn.Sequential
    avgpool: nn.AdaptiveAvgPool2d
    classifier: nn.Sequential

    def __init__(self, cfg: List[Union[str, int]] = None, num_classes: int = 1000, batch_norm: bool = True,
                 dropout: float = 0.5, *, key=None) -> None:
        keys = jrandom.split(jrandom.PRNGKey(0) if key is None else key, 4)
        self.features = _make_layers(cfg, batch_norm, key=keys[0])
        self.avgpool = nn.AdaptiveAvgPool2d((7, 7))
        self.classifier = nn.Sequential([
            nn.Linear(512 * 7 * 7, 4096, key=keys[1]),
            nn.Dropout(p=dropout),
            nn.Linear(4096, 4096, key=keys[2]),
            nn.Lambda(F.relu),
            nn.Dropout(p=dropout),
            nn.Linear(4096, num_classes, key=keys[3]),
        ])

    def __call__(self, x, *, key=None):
        k1, k2 = jrandom.split(key, 2)
        x = self.avgpool(self.features(x, key=k1))
        return self.classifier(F.ravel(x), key=k2)


def _make_layers(cfg: List[Union[str, int]], batch_norm: bool = False, key=None) -> nn.Sequential:
    widths = [cast(int, v) for v in cfg if v != "M"]
    keys = iter(jrandom.split(key, len(widths)))
    seq: List[nn.Module] = []
    cin = 3
    for v in cfg:
        if v == "M":
            seq.append(nn.MaxPool2d(kernel_size=2, stride=2))
            continue
        seq.append(nn.Conv2d(cin, v, kernel_size=3, padding=1, key=next(keys)))
        if batch_norm:
            seq.append(nn.BatchNorm(v, axis_name="batch"))
        seq.append(nn.Lambda(F.relu))
        cin = v
    return nn.Sequential(seq)


def _vgg(cfg: str, batch_norm: bool, torch_weights: Optional[str], **kwargs: Any) -> VGG:
    model = VGG(cfg=_cfgs[cfg], batch_norm=batch_norm, **kwargs)
    if torch_weights:
        model = load_torch_weights(model, torch_weights=torch_weights)
    return model


def vgg11(torch_weights: str = None, **kwargs: Any) -> VGG:
    """VGG-11 (configuration "A")."""
    return _vgg("A", False, torch_weights, **kwargs)


def vgg11_bn(torch_weights: str = None, **kwargs: Any) -> VGG:
    """VGG-11 with batch normalisation."""
    return _vgg("A", True, torch_weights, **kwargs)


def vgg13(torch_weights: str = None, **kwargs: Any) -> VGG:
    """VGG-13 (configuration "B")."""
    return _vgg("B", False, torch_weights, **kwargs)


def vgg13_bn(torch_weights: str = None, **kwargs: Any) -> VGG:
    return _vgg("B", True, torch_weights, **kwargs)


def vgg16(torch_weights: str = None, **kwargs: Any) -> VGG:
    """VGG-16 (configuration "D")."""
    return _vgg("D", False, torch_weights, **kwargs)


def vgg16_bn(torch_weights: str = None, **kwargs: Any) -> VGG:
    return _vgg("D", True, torch_weights, **kwargs)


def vgg19(torch_weights: str = None, **kwargs: Any) -> VGG:
    """VGG-19 (configuration "E")."""
    return _vgg("E", False, torch_weights, **kwargs)


def vgg19_bn(torch_weights: str = None, **kwargs: Any) -> VGG:
    return _vgg("E", True, torch_weights, **kwargs)
