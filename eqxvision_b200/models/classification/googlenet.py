"""GoogLeNet / Inception v1 (reference: models/classification/googlenet.py).

BasicConv2d = conv (no bias) -> BatchNorm(eps 1e-3) -> ReLU. _Inception = channel concat of four branches: 1x1 | 1x1 ->
3x3 | 1x1 -> 3x3 (the "5x5" branch is 3x3, the known torchvision quirk kept by the reference, googlenet.py:214-216) |
max-pool 3x3/1 p1 -> 1x1. Stage pools are 3x3/2 (2x2/2 before inception5) with `use_ceil=True`.
Device lowering: every branch's last convolution stores straight into its channel slice of the block's output buffer
(no concat pass); BN + ReLU are GEMM epilogues.
The auxiliary classifiers (`aux_logits=True`) pool 14x14 maps adaptively to 4x4: an uneven split, which Equinox cuts
into consecutive blocks of 4, 4, 3, 3 (torch would use overlapping windows) - the device kernel follows
Equinox (`eqxv_adaptive_avgpool_nhwc_bf16`), and the model returns `(logits, aux2, aux1)` like the reference
(googlenet.py:174-175).
"""
import copy
import warnings
from typing import Any, Callable, List, Optional

from ... import functional as F
from ... import nn
from ... import random as jrandom
from ...utils import load_torch_weights


class BasicConv2d(nn.Module):
    conv: nn.Conv2d
    bn: nn.BatchNorm

    def __init__(self, in_channels: int, out_channels: int, *, key=None, **kwargs: Any) -> None:
        self.conv = nn.Conv2d(in_channels, out_channels, use_bias=False, key=key, **kwargs)
        self.bn = nn.BatchNorm(out_channels, axis_name="batch", eps=0.001)

    def __call__(self, x, *, key=None):
        return F.relu(self.bn(self.conv(x), key=key))


class _Inception(nn.Module):
    branch1: nn.Module
    branch2: nn.Sequential
    branch3: nn.Sequential
    branch4: nn.Sequential

    def __init__(self, in_channels: int, ch1x1: int, ch3x3red: int, ch3x3: int, ch5x5red: int, ch5x5: int,
                 pool_proj: int, conv_block: Optional[Callable[..., nn.Module]] = None, *, key=None) -> None:
        conv_block = BasicConv2d if conv_block is None else conv_block
        # the reference splits 5 keys and indexes keys[5] (googlenet.py:201,224), which jax clamps to the last one
        keys = list(jrandom.split(key, 5))
        keys.append(keys[-1])
        self.branch1 = conv_block(in_channels, ch1x1, kernel_size=1, key=keys[0])
        self.branch2 = nn.Sequential([conv_block(in_channels, ch3x3red, kernel_size=1, key=keys[1]),
                                      conv_block(ch3x3red, ch3x3, kernel_size=3, padding=1, key=keys[2])])
        self.branch3 = nn.Sequential([conv_block(in_channels, ch5x5red, kernel_size=1, key=keys[3]),
                                      conv_block(ch5x5red, ch5x5, kernel_size=3, padding=1, key=keys[4])])
        self.branch4 = nn.Sequential([nn.MaxPool2d(kernel_size=3, stride=1, padding=1, use_ceil=True),
                                      conv_block(in_channels, pool_proj, kernel_size=1, key=keys[5])])

    def __call__(self, x, *, key=None):
        keys = [None] * 4 if key is None else jrandom.split(key, 4)
        return F.concat_channels([self.branch1(x, key=keys[0]), self.branch2(x, key=keys[1]),
                                  self.branch3(x, key=keys[2]), self.branch4(x, key=keys[3])])


class InceptionAux(nn.Module):
    conv: nn.Module
    fc1: nn.Linear
    fc2: nn.Linear
    dropout: nn.Dropout
    avgpool: nn.AdaptiveAvgPool2d

    def __init__(self, in_channels: int, num_classes: int, conv_block: Optional[Callable[..., nn.Module]] = None,
                 dropout: float = 0.7, *, key=None) -> None:
        conv_block = BasicConv2d if conv_block is None else conv_block
        keys = jrandom.split(key, 3)
        self.conv = conv_block(in_channels, 128, kernel_size=1, key=keys[0])
        self.fc1 = nn.Linear(2048, 1024, key=keys[1])
        self.fc2 = nn.Linear(1024, num_classes, key=keys[2])
        self.dropout = nn.Dropout(p=dropout)
        self.avgpool = nn.AdaptiveAvgPool2d((4, 4))

    def __call__(self, x, *, key=None):
        keys = [None] * 2 if key is None else jrandom.split(key, 2)
        x = self.conv(self.avgpool(x), key=keys[0])
        x = F.relu(self.fc1(F.ravel(x)))
        return self.fc2(self.dropout(x, key=keys[1]))


class GoogLeNet(nn.Module):
    """`torchvision.models.googlenet` as ported by the reference (googlenet.py:15-177)."""

    aux_logits: bool
    conv1: nn.Module
    maxpool1: nn.MaxPool2d
    conv2: nn.Module
    conv3: nn.Module
    maxpool2: nn.MaxPool2d
    inception3a: nn.Module
    inception3b: nn.Module
    maxpool3: nn.MaxPool2d
    inception4a: nn.Module
    inception4b: nn.Module
    inception4c: nn.Module
    inception4d: nn.Module
    inception4e: nn.Module
    maxpool4: nn.MaxPool2d
    inception5a: nn.Module
    inception5b: nn.Module
    aux1: nn.Module
    aux2: nn.Module
    avgpool: nn.AdaptiveAvgPool2d
    dropout: nn.Dropout
    fc: nn.Linear

    def __init__(self, num_classes: int = 1000, aux_logits: bool = False, blocks: Optional[List[Any]] = None,
                 dropout: float = 0.2, dropout_aux: float = 0.7, *, key=None) -> None:
        blocks = [BasicConv2d, _Inception, InceptionAux] if blocks is None else blocks
        assert len(blocks) == 3
        conv_block, inception_block, inception_aux_block = blocks
        keys = jrandom.split(jrandom.PRNGKey(0) if key is None else key, 20)
        self.aux_logits = aux_logits
        self.conv1 = conv_block(3, 64, kernel_size=7, stride=2, padding=3, key=keys[0])
        self.maxpool1 = nn.MaxPool2d(3, stride=2, use_ceil=True)
        self.conv2 = conv_block(64, 64, kernel_size=1, key=keys[1])
        self.conv3 = conv_block(64, 192, kernel_size=3, padding=1, key=keys[2])
        self.maxpool2 = nn.MaxPool2d(3, stride=2, use_ceil=True)
        self.inception3a = inception_block(192, 64, 96, 128, 16, 32, 32, key=keys[3])
        self.inception3b = inception_block(256, 128, 128, 192, 32, 96, 64, key=keys[4])
        self.maxpool3 = nn.MaxPool2d(3, stride=2, use_ceil=True)
        self.inception4a = inception_block(480, 192, 96, 208, 16, 48, 64, key=keys[5])
        self.inception4b = inception_block(512, 160, 112, 224, 24, 64, 64, key=keys[6])
        self.inception4c = inception_block(512, 128, 128, 256, 24, 64, 64, key=keys[7])
        self.inception4d = inception_block(512, 112, 144, 288, 32, 64, 64, key=keys[8])
        self.inception4e = inception_block(528, 256, 160, 320, 32, 128, 128, key=keys[9])
        self.maxpool4 = nn.MaxPool2d(2, stride=2, use_ceil=True)
        self.inception5a = inception_block(832, 256, 160, 320, 32, 128, 128, key=keys[10])
        self.inception5b = inception_block(832, 384, 192, 384, 48, 128, 128, key=keys[11])
        self.aux1 = None
        self.aux2 = None
        if aux_logits:
            self.aux1 = inception_aux_block(512, num_classes, dropout=dropout_aux, key=keys[12])
            self.aux2 = inception_aux_block(528, num_classes, dropout=dropout_aux, key=keys[13])
        self.avgpool = nn.AdaptiveAvgPool2d((1, 1))
        self.dropout = nn.Dropout(p=dropout)
        self.fc = nn.Linear(1024, num_classes, key=keys[14])

    def __call__(self, x, *, key=None):
        if key is None:  # googlenet.py:112-113
            raise RuntimeError("The model requires a PRNGKey.")
        # 14 keys are split and keys[14], keys[15] are used (googlenet.py:114,160,167): jax clamps the index
        keys = list(jrandom.split(key, 14))
        keys += [keys[-1]] * 2
        x = self.conv1(x, key=keys[0])
        x = self.maxpool1(x, key=keys[1])
        x = self.conv2(x, key=keys[2])
        x = self.conv3(x, key=keys[3])
        x = self.maxpool2(x)
        x = self.inception3a(x, key=keys[4])
        x = self.inception3b(x, key=keys[5])
        x = self.maxpool3(x)
        x = self.inception4a(x, key=keys[6])
        if self.aux_logits:
            aux1 = self.aux1(x, key=keys[7])
        x = self.inception4b(x, key=keys[8])
        x = self.inception4c(x, key=keys[9])
        x = self.inception4d(x, key=keys[10])
        if self.aux_logits:
            aux2 = self.aux2(x, key=keys[11])
        x = self.inception4e(x, key=keys[12])
        x = self.maxpool4(x)
        x = self.inception5a(x, key=keys[13])
        x = self.inception5b(x, key=keys[14])
        x = F.ravel(self.avgpool(x))
        x = self.fc(self.dropout(x, key=keys[15]))
        if self.aux_logits:
            return x, aux2, aux1
        return x


def googlenet(torch_weights: str = None, **kwargs: Any) -> GoogLeNet:
    """GoogLeNet ("Going Deeper with Convolutions", arXiv 1409.4842); minimum input 15x15. With `torch_weights` the
    model is built WITH the auxiliary heads so that a torchvision checkpoint loads positionally, then `aux_logits` is
    switched off unless requested (googlenet.py:320-335). (In the reference `googlenet(torch_weights=..., aux_logits=True)`
    raises a TypeError because `aux_logits` reaches the constructor twice, googlenet.py:323-325; here it is accepted.)"""
    if torch_weights:
        use_aux = kwargs.get("aux_logits", False)
        kwargs.pop("aux_logits", None)
        model = GoogLeNet(aux_logits=True, **kwargs)
        model = load_torch_weights(model, torch_weights=torch_weights)
        if not use_aux:
            # eqx.tree_at(lambda m: m.aux_logits, model, replace=False) (googlenet.py:327): a functional update
            model = copy.copy(model)
            model.__dict__.pop("_eqxv_plans", None)
            object.__setattr__(model, "aux_logits", False)
        else:
            warnings.warn("Loaded torch_weights weights for GoogLeNet. But, aux-branch weights are un-trained.")
    else:
        model = GoogLeNet(**kwargs)
    return model
