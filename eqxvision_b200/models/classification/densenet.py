"""DenseNet-121/161/169/201 (reference: models/classification/densenet.py).

Pre-activation order (densenet.py:64-65): concat(previous features) -> BN -> ReLU -> conv1x1 -> BN ->
ReLU -> conv3x3 (growth_rate channels); transitions BN -> ReLU -> conv1x1 -> avgpool 2x2.
Device lowering: the BN+ReLU in front of conv1x1 is a standalone per-channel affine pass
(`eqxv_eltwise_bf16`), the BN+ReLU after conv1x1 is that GEMM's epilogue, and the channel concat is
never materialised: every layer's 3x3 conv stores straight into its channel slice of one block-wide
NHWC buffer (the `capacity` hint below tells the engine how wide the block will get).
"""
from typing import Any, Optional, Sequence, Tuple, Union

from ... import _trace as T
from ... import functional as F
from ... import nn
from ... import random as jrandom
from ...utils import load_torch_weights


class _DenseLayer(nn.Module):
    norm1: nn.BatchNorm
    relu: nn.Lambda
    conv1: nn.Conv2d
    norm2: nn.BatchNorm
    conv2: nn.Conv2d
    dropout: nn.Dropout

    def __init__(self, num_input_features: int, growth_rate: int, bn_size: int, drop_rate: float, key) -> None:
        k1, k2 = jrandom.split(key, 2)
        mid = bn_size * growth_rate
        self.norm1 = nn.BatchNorm(num_input_features, axis_name="batch")
        self.relu = nn.Lambda(F.relu)
        self.conv1 = nn.Conv2d(num_input_features, mid, kernel_size=1, stride=1, use_bias=False, key=k1)
        self.norm2 = nn.BatchNorm(mid, axis_name="batch")
        self.conv2 = nn.Conv2d(mid, growth_rate, kernel_size=3, stride=1, padding=1, use_bias=False, key=k2)
        self.dropout = nn.Dropout(p=float(drop_rate))

    def __call__(self, x, *, key=None, capacity: Optional[int] = None):
        feats = [x] if T.is_sym(x) else list(x)
        cat = T.concat_channels(feats, capacity=capacity)
        y = self.conv1(self.relu(self.norm1(cat)))
        y = self.conv2(self.relu(self.norm2(y)))
        return self.dropout(y, key=key)


class _DenseBlock(nn.Module):
    layers: Sequence[nn.Module]
    num_layers: int

    def __init__(self, num_layers: int, num_input_features: int, bn_size: int, growth_rate: int, drop_rate: float,
                 key=None) -> None:
        keys = jrandom.split(key, num_layers)
        self.layers = [_DenseLayer(num_input_features + i * growth_rate, growth_rate=growth_rate, bn_size=bn_size,
                                   drop_rate=drop_rate, key=keys[i]) for i in range(num_layers)]
        self.num_layers = num_layers

    def __call__(self, x, *, key=None):
        keys = jrandom.split(key, self.num_layers)
        feats = [x]
        final = x.shape[0] + sum(l.conv2.out_channels for l in self.layers)
        for layer, k in zip(self.layers, keys):
            feats.append(layer(feats, key=k, capacity=final))
        return T.concat_channels(feats, capacity=final)


class _Transition(nn.Module):
    layers: nn.Sequential

    def __init__(self, num_input_features: int, num_output_features: int, key=None) -> None:
        self.layers = nn.Sequential([
            nn.BatchNorm(num_input_features, axis_name="batch"),
            nn.Lambda(F.relu),
            nn.Conv2d(num_input_features, num_output_features, kernel_size=1, stride=1, use_bias=False, key=key),
            nn.AvgPool2d(kernel_size=2, stride=2),
        ])

    def __call__(self, x, *, key=None):
        return self.layers(x, key=key)


class DenseNet(nn.Module):
    """`torchvision.models.densenet` layout as ported by the reference (densenet.py:136-229)."""

    features: nn.Sequential
    classifier: nn.Linear

    def __init__(self, growth_rate: int = 32, block_config: Tuple[int, int, int, int] = (6, 12, 24, 16),
                 num_init_features: int = 64, bn_size: int = 4, drop_rate: float = 0, num_classes: int = 1000,
                 *, key=None) -> None:
        key = jrandom.PRNGKey(0) if key is None else key
        keys = jrandom.split(key, 2 * len(block_config) + 2)
        pick = lambda ks, i: ks[min(i, len(ks) - 1)]  # noqa: E731  jax clamps out-of-range key indices
        seq = [
            nn.Conv2d(3, num_init_features, kernel_size=7, stride=2, padding=3, use_bias=False, key=keys[0]),
            nn.BatchNorm(num_init_features, axis_name="batch"),
            nn.Lambda(F.relu),
            nn.MaxPool2d(kernel_size=3, stride=2, padding=1),
        ]
        width = num_init_features
        for i, depth in enumerate(block_config):
            keys = jrandom.split(pick(keys, i * 2 + 1), 3)
            seq.append(_DenseBlock(num_layers=depth, num_input_features=width, bn_size=bn_size,
                                   growth_rate=growth_rate, drop_rate=drop_rate, key=keys[0]))
            width += depth * growth_rate
            if i != len(block_config) - 1:
                seq.append(_Transition(num_input_features=width, num_output_features=width // 2,
                                       key=pick(keys, i * 2 + 2)))
                width //= 2
        seq += [nn.BatchNorm(width, axis_name="batch"), nn.Lambda(F.relu), nn.AdaptiveAvgPool2d((1, 1))]
        self.features = nn.Sequential(seq)
        self.classifier = nn.Linear(width, num_classes, key=keys[-1])

    def __call__(self, x, *, key=None):
        return self.classifier(F.ravel(self.features(x, key=key)))


def _densenet(growth_rate, block_config, num_init_features, torch_weights=None, **kwargs: Any) -> DenseNet:
    model = DenseNet(growth_rate, block_config, num_init_features, **kwargs)
    if torch_weights:
        model = load_torch_weights(model, torch_weights=torch_weights)
    return model


def densenet121(torch_weights: str = None, **kwargs: Any) -> DenseNet:
    """DenseNet-121 (densenet.py:242-256)."""
    return _densenet(32, (6, 12, 24, 16), 64, torch_weights, **kwargs)


def densenet161(torch_weights: str = None, **kwargs: Any) -> DenseNet:
    return _densenet(48, (6, 12, 36, 24), 96, torch_weights, **kwargs)


def densenet169(torch_weights: str = None, **kwargs: Any) -> DenseNet:
    return _densenet(32, (6, 12, 32, 32), 64, torch_weights, **kwargs)


def densenet201(torch_weights: str = None, **kwargs: Any) -> DenseNet:
    return _densenet(32, (6, 12, 48, 32), 64, torch_weights, **kwargs)
