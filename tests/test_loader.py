"""Positional checkpoint import (reference contract: eqxvision/utils.py:120-219)."""
import collections

import pytest
import torch

import eqxvision_b200 as eb
from eqxvision_b200 import nn
from eqxvision_b200.utils import _make_divisible, load_torch_weights
from oracle import checkpoints as ck


def test_resnet50_positional_load(save_checkpoint):
    sd = ck.torchvision_state_dict("resnet50", seed=1)
    net = eb.models.resnet50(torch_weights=save_checkpoint(sd))
    assert torch.equal(net.conv1.weight, sd["conv1.weight"])
    assert torch.equal(net.layer1.layers[0].downsample.layers[0].weight, sd["layer1.0.downsample.0.weight"])
    assert torch.equal(net.layer3.layers[5].bn2.weight, sd["layer3.5.bn2.weight"])
    assert torch.equal(net.layer3.layers[5].bn2.running_mean, sd["layer3.5.bn2.running_mean"])
    assert torch.equal(net.layer4.layers[2].bn3.running_var, sd["layer4.2.bn3.running_var"])
    assert torch.equal(net.fc.weight, sd["fc.weight"]) and torch.equal(net.fc.bias, sd["fc.bias"])
    arrays = [l for l in nn.tree_leaves(net) if isinstance(l, torch.Tensor)]
    assert len(arrays) == 161
    states = [l for l in nn.tree_leaves(net) if isinstance(l, nn.StateIndex)]
    assert len(states) == 2 * 53 and states[0].value is False


def test_load_returns_new_model_and_keeps_input_untouched(save_checkpoint):
    sd = ck.torchvision_state_dict("resnet18", seed=1)
    fresh = eb.models.resnet18()
    w0 = fresh.conv1.weight.clone()
    m0 = fresh.bn1.running_mean.clone()
    loaded = load_torch_weights(fresh, save_checkpoint(sd))
    assert loaded is not fresh
    assert torch.equal(fresh.conv1.weight, w0) and torch.equal(fresh.bn1.running_mean, m0)
    assert torch.equal(loaded.bn1.running_mean, sd["bn1.running_mean"])


def test_vit_dino_order_and_reshapes(save_checkpoint):
    sd = ck.vit_state_dict(embed_dim=192, depth=2, heads=3, num_classes=10, seed=3)
    net = eb.models.vit_tiny(depth=2, num_classes=10, torch_weights=save_checkpoint(sd))
    assert net.cls_token.shape == (1, 192) and net.pos_embed.shape == (197, 192)  # (1,1,D)->(1,D), utils.py:197
    assert torch.equal(net.cls_token, sd["cls_token"].reshape(1, 192))
    assert torch.equal(net.patch_embed.proj.bias, sd["patch_embed.proj.bias"].reshape(192, 1, 1))
    assert torch.equal(net.blocks[1].attn.qkv.weight, sd["blocks.1.attn.qkv.weight"])
    assert torch.equal(net.blocks[1].mlp.fc2.bias, sd["blocks.1.mlp.fc2.bias"])
    assert torch.equal(net.fc.weight, sd["head.weight"])


def test_conv_bias_reshaped_to_o11(save_checkpoint):
    conv = nn.Conv2d(8, 4, 1, key=eb.random.PRNGKey(0))
    sd = collections.OrderedDict(weight=torch.randn(4, 8, 1, 1), bias=torch.randn(4))
    loaded = load_torch_weights(conv, save_checkpoint(sd))
    assert loaded.bias.shape == (4, 1, 1) and torch.equal(loaded.bias.reshape(-1), sd["bias"])


def test_short_checkpoint_raises_stopiteration(save_checkpoint):
    sd = ck.torchvision_state_dict("resnet18", seed=1)
    short = collections.OrderedDict(list(sd.items())[:30])
    with pytest.raises((StopIteration, RuntimeError)):
        eb.models.resnet18(torch_weights=save_checkpoint(short))


def test_extra_trailing_tensors_are_ignored_and_shapes_are_reshaped_silently(save_checkpoint):
    lin = nn.Linear(6, 4, key=eb.random.PRNGKey(0))
    sd = collections.OrderedDict(w=torch.arange(24.).reshape(6, 4), b=torch.zeros(4), extra=torch.ones(3))
    loaded = load_torch_weights(lin, save_checkpoint(sd))
    assert loaded.weight.shape == (4, 6)  # same size, different shape: reshaped (utils.py:196-197)
    assert torch.equal(loaded.weight.reshape(-1), torch.arange(24.))


def test_empty_path_raises_valueerror():
    with pytest.raises(ValueError):
        load_torch_weights(eb.models.resnet18(), None)


def test_url_tables_and_make_divisible():
    from eqxvision_b200.utils import CLASSIFICATION_URLS, SEGMENTATION_URLS

    assert CLASSIFICATION_URLS["resnet50"].endswith(".pth") and "deeplabv3_resnet50" in SEGMENTATION_URLS
    assert len(CLASSIFICATION_URLS) == 70
    assert _make_divisible(32 * 1.4, 8) == 48 and _make_divisible(24 * 1.4, 8) == 32
    assert _make_divisible(10, 8) == 16 and _make_divisible(7, 8) == 8


def test_bn_folding_matches_definition():
    bn = nn.BatchNorm(5, axis_name="batch")
    bn.weight = torch.rand(5) + 0.5
    bn.bias = torch.randn(5)
    bn.state_index.value = (torch.randn(5), torch.rand(5) + 0.1)
    scale, shift = bn.folded()
    x = torch.randn(7, 5)
    ref = (x - bn.running_mean) / torch.sqrt(bn.running_var + 1e-5) * bn.weight + bn.bias
    assert torch.allclose(x * scale + shift, ref, atol=1e-5)
