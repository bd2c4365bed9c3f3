"""Deterministic synthetic checkpoints and inputs (workload generator for tests, smoke() and bench.py).

No pretrained weights can be downloaded here, so parity runs on seeded random `state_dict`s laid out
exactly like the files the reference loads: torchvision key order for the CNNs (the architectures
come from torchvision itself, which is the lineage of the reference's model files), DINO/timm key
order for ViT.  BatchNorm running statistics are calibrated with one
train-mode pass over seeded noise so that activations stay O(1) through the depth of the net, and
the affine parameters are perturbed so that a swapped gamma/beta or mean/var cannot go unnoticed.
Everything is generated with the CPU generator: the same seed gives the same tensors on every box.
"""
from __future__ import annotations

import collections
import math
from typing import Dict

import torch


IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


def synthetic_images(n: int, c: int = 3, h: int = 224, w: int = 224, seed: int = 0,
                     normalize: bool = True) -> torch.Tensor:
    """Seeded fp32 NCHW batch: U[0,1) pixels (the README's `jr.uniform(key, (B,3,224,224))`,
    README.md:45), by default passed through the ImageNet normalisation the reference's own test
    fixture applies (tests/conftest.py:27) so that the stem sees zero-centred inputs as it would in use."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand((n, c, h, w), generator=g)
    if normalize and c == 3:
        mean = torch.tensor(IMAGENET_MEAN).reshape(1, 3, 1, 1)
        std = torch.tensor(IMAGENET_STD).reshape(1, 3, 1, 1)
        x = (x - mean) / std
    return x


def _perturb_and_calibrate(model: torch.nn.Module, seed: int, calib_shape=(4, 3, 96, 96)) -> None:
    g = torch.Generator().manual_seed(seed + 1000)
    bns = [m for m in model.modules() if isinstance(m, torch.nn.modules.batchnorm._BatchNorm)]
    with torch.no_grad():
        for m in bns:
            m.weight.copy_(torch.rand(m.weight.shape, generator=g) * 0.8 + 0.6)
            m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
            m.momentum = 1.0  # running stats := statistics of the calibration batch
        # the last BatchNorm of every residual branch gets a small gain, as in trained ResNets
        # (torchvision's zero_init_residual targets): an untrained residual net with unit gains is
        # chaotic and amplifies ANY rounding (bf16 or fp32 summation order) far more than a real model
        for m in model.modules():
            last = {"Bottleneck": "bn3", "BasicBlock": "bn2"}.get(type(m).__name__)
            if last is not None:
                bn = getattr(m, last)
                bn.weight.copy_(torch.rand(bn.weight.shape, generator=g) * 0.3 + 0.2)
            if type(m).__name__ == "CNBlock":  # ConvNeXt: the default layer scale 1e-6 would hide the whole branch
                m.layer_scale.copy_(torch.rand(m.layer_scale.shape, generator=g) * 0.4 + 0.1)
            if type(m).__name__ == "LayerNorm2d":  # torchvision.models.convnext.LayerNorm2d only
                m.weight.copy_(torch.rand(m.weight.shape, generator=g) * 0.8 + 0.6)
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.1)
            if type(m).__name__ == "ResBottleneckBlock":  # RegNet: x + f(x), the branch ends in f.c = (conv, bn)
                bn = m.f.c[1]
                bn.weight.copy_(torch.rand(bn.weight.shape, generator=g) * 0.3 + 0.2)
        for m in model.modules():
            if isinstance(m, (torch.nn.Linear, torch.nn.Conv2d)) and m.bias is not None:
                m.bias.copy_(torch.randn(m.bias.shape, generator=g) * 0.05)
    if bns:
        model.train()
        with torch.no_grad():
            x = torch.rand(calib_shape, generator=g)
            if calib_shape[1] == 3:  # same input distribution as synthetic_images(normalize=True)
                x = (x - torch.tensor(IMAGENET_MEAN).reshape(1, 3, 1, 1)) / torch.tensor(IMAGENET_STD).reshape(1, 3, 1, 1)
            out = model(x)
            del out
        # Statistics estimated from a handful of samples can be degenerate (the ASPP pooling branch sees
        # one 1x1 map per calibration image): floor the variance so that no BatchNorm turns into a
        # x30 amplifier of rounding noise, which no trained network has.
        with torch.no_grad():
            for m in bns:
                m.running_var.clamp_(min=0.05)
    model.eval()


def torchvision_model(arch: str, seed: int = 0, calib_hw: int = 96, **kwargs) -> torch.nn.Module:
    """seeded torchvision architecture with calibrated BN (eval mode)"""
    import torchvision

    torch.manual_seed(seed)
    if arch.startswith("deeplabv3") or arch.startswith("fcn") or arch.startswith("lraspp"):
        ctor = getattr(torchvision.models.segmentation, arch)
        model = ctor(weights=None, weights_backbone=None, **kwargs)
    else:
        model = getattr(torchvision.models, arch)(weights=None, **kwargs)
    _perturb_and_calibrate(model, seed, (4, 3, calib_hw, calib_hw))
    return model


def torchvision_state_dict(arch: str, seed: int = 0, **kwargs) -> "collections.OrderedDict[str, torch.Tensor]":
    return torchvision_model(arch, seed, **kwargs).state_dict()


def vit_state_dict(embed_dim=768, depth=12, heads=12, mlp_ratio=4, patch=16, img=224, num_classes=1000,
                   seed: int = 0) -> "collections.OrderedDict[str, torch.Tensor]":
    """DINO/timm-ordered ViT checkpoint (the only order the reference can load, vit.py:163-171)."""
    g = torch.Generator().manual_seed(seed)

    def rn(*shape, std=0.02):
        return torch.randn(shape, generator=g) * std

    d = embed_dim
    np_ = (img // patch) ** 2
    sd = collections.OrderedDict()
    sd["cls_token"] = rn(1, 1, d, std=0.5)
    sd["pos_embed"] = rn(1, np_ + 1, d, std=0.5)
    sd["patch_embed.proj.weight"] = rn(d, 3, patch, patch, std=(3 * patch * patch) ** -0.5)
    sd["patch_embed.proj.bias"] = rn(d, std=0.1)
    hidden = int(d * mlp_ratio)
    for i in range(depth):
        p = f"blocks.{i}."
        sd[p + "norm1.weight"] = 1.0 + rn(d, std=0.1)
        sd[p + "norm1.bias"] = rn(d, std=0.1)
        sd[p + "attn.qkv.weight"] = rn(3 * d, d, std=d ** -0.5)
        sd[p + "attn.qkv.bias"] = rn(3 * d, std=0.1)
        sd[p + "attn.proj.weight"] = rn(d, d, std=d ** -0.5)
        sd[p + "attn.proj.bias"] = rn(d, std=0.1)
        sd[p + "norm2.weight"] = 1.0 + rn(d, std=0.1)
        sd[p + "norm2.bias"] = rn(d, std=0.1)
        sd[p + "mlp.fc1.weight"] = rn(hidden, d, std=d ** -0.5)
        sd[p + "mlp.fc1.bias"] = rn(hidden, std=0.1)
        sd[p + "mlp.fc2.weight"] = rn(d, hidden, std=hidden ** -0.5)
        sd[p + "mlp.fc2.bias"] = rn(d, std=0.1)
    sd["norm.weight"] = 1.0 + rn(d, std=0.1)
    sd["norm.bias"] = rn(d, std=0.1)
    if num_classes > 0:
        sd["head.weight"] = rn(num_classes, d, std=d ** -0.5)
        sd["head.bias"] = rn(num_classes, std=0.1)
    return sd


def swin_model(arch: str = "swin_t", seed: int = 0, tanh_gelu: bool = True) -> torch.nn.Module:
    """Seeded torchvision Swin (the lineage of the reference's swin.py) with perturbed LayerNorm affine
    parameters, biases and relative-position tables. `tanh_gelu` swaps torchvision's erf-GELU for the
    tanh approximation the reference computes (jnn.gelu default, swin.py:567) so that the torchvision
    forward is a known-answer generator for the reference's arithmetic."""
    import torchvision

    torch.manual_seed(seed)
    if isinstance(arch, dict):   # a custom (small) configuration: kwargs of torchvision's SwinTransformer
        m = torchvision.models.swin_transformer.SwinTransformer(**arch).eval()
    else:
        m = getattr(torchvision.models, arch)(weights=None).eval()
    g = torch.Generator().manual_seed(seed + 1000)
    with torch.no_grad():
        for n, p in m.named_parameters():
            if "norm" in n and n.endswith("weight"):
                p.copy_(1 + 0.1 * torch.randn(p.shape, generator=g))
            elif n.endswith("bias"):
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
            elif "relative_position_bias_table" in n:
                p.copy_(0.5 * torch.randn(p.shape, generator=g))
            elif n.endswith("logit_scale"):     # Swin-V2: some heads above the log(100) clamp, some below
                p.copy_(p + 2.5 * torch.rand(p.shape, generator=g))
    if tanh_gelu:
        for mod in m.modules():
            if isinstance(mod, torchvision.ops.misc.MLP):
                mod[1] = torch.nn.GELU(approximate="tanh")
    return m
