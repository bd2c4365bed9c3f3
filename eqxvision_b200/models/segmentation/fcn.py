"""FCN head and model (reference: models/segmentation/fcn.py). `FCNHead` is DeepLabV3's auxiliary head:
conv3x3 (no bias) -> BN -> ReLU -> Dropout(0.1) -> conv1x1 (with bias)."""
from typing import Callable, Optional

from ... import functional as F
from ... import nn
from ... import random as jrandom
from ...experimental import intermediate_layer_getter
from ...utils import load_torch_weights
from ..classification import resnet
from ._utils import _SimpleSegmentationModel


class FCN(_SimpleSegmentationModel):
    """Ported from `torchvision.models.segmentation.fcn`"""


class FCNHead(nn.Sequential):
    def __init__(self, in_channels: int, out_channels: int, *, key) -> None:
        k1, k2 = jrandom.split(key, 2)
        mid = in_channels // 4
        super().__init__([
            nn.Conv2d(in_channels, mid, 3, padding=1, use_bias=False, key=k1),
            nn.BatchNorm(mid, axis_name="batch"),
            nn.Lambda(F.relu),
            nn.Dropout(0.1),
            nn.Conv2d(mid, out_channels, 1, key=k2),
        ])


def _check_taps(aux_in_channels, num_layers):
    if aux_in_channels is not None and num_layers != 2:
        raise ValueError("aux_in_channels requires the intermediate_layers to return exactly 2 layers "
                         "corresponding to aux and final.")
    if aux_in_channels is None and num_layers != 1:
        raise ValueError(f"With no aux_in_channels, the aux layer is disabled. Received {num_layers} "
                         f"from intermediate_layers, expected number of layers is 1.")


def _assemble(model_cls, backbone, intermediate_layers, silence_layers, classifier_module, classifier_in_channels,
              aux_module, aux_in_channels, num_classes, torch_weights, key):
    k1, k2 = jrandom.split(jrandom.PRNGKey(0) if key is None else key, 2)
    if backbone is None:
        backbone = resnet.resnet50(replace_stride_with_dilation=[False, True, True])
    n_taps = len(intermediate_layers(backbone))
    if silence_layers is None:
        silence_layers = lambda m: m.fc  # noqa: E731
    _check_taps(aux_in_channels, n_taps)
    # the classification head is dropped (Identity) so that its weights are not part of the checkpoint walk
    backbone = nn.tree_at(silence_layers, backbone, replace_fn=lambda _: nn.Identity())
    backbone = intermediate_layer_getter(backbone, intermediate_layers)
    classifier = classifier_module(in_channels=classifier_in_channels, out_channels=num_classes, key=k1)
    aux = None
    if aux_in_channels is not None:
        aux = aux_module(in_channels=aux_in_channels, out_channels=num_classes, key=k2)
    model = model_cls(backbone, classifier, aux)
    if torch_weights:
        return load_torch_weights(model, torch_weights=torch_weights)
    return model


def fcn(
    num_classes: Optional[int] = 21,
    backbone: nn.Module = None,
    intermediate_layers: Callable = None,
    classifier_module: nn.Module = None,
    classifier_in_channels: int = 2048,
    aux_in_channels: int = None,
    silence_layers: Callable = None,
    torch_weights: str = None,
    *,
    key=None,
) -> FCN:
    """FCN-ResNet50 by default (fcn.py:37-120); the aux head reuses `classifier_module`."""
    head = FCNHead if classifier_module is None else classifier_module
    return _assemble(FCN, backbone, intermediate_layers, silence_layers, head, classifier_in_channels, head,
                     aux_in_channels, num_classes, torch_weights, key)
