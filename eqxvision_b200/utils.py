"""`eqxvision.utils` surface: checkpoint import, URL tables, `_make_divisible`.

`load_torch_weights` keeps the reference's POSITIONAL contract (utils.py:120-219, SURVEY.md
Appendix B): tensors of the torch `state_dict` whose key contains neither "running" nor
"num_batches" are consumed in file order by the model's array leaves in pytree (field) order and
reshaped to the leaf's shape; `running_mean`/`running_var` pairs are consumed in file order by the
BatchNorm state slots in pytree order. No name matching, no transposes.
"""
from __future__ import annotations

import json
import logging
import os
from typing import Optional

import torch

from . import nn

_TEMP_DIR = "/tmp/.eqx"

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "_urls.json")) as _f:
    _tables = json.load(_f)
SEGMENTATION_URLS = _tables["SEGMENTATION_URLS"]
CLASSIFICATION_URLS = _tables["CLASSIFICATION_URLS"]


def _make_divisible(v: float, divisor: int, min_value: Optional[int] = None) -> int:
    """Channel rounding used by MobileNet/EfficientNet configs (utils.py:104-117)."""
    floor = divisor if min_value is None else min_value
    rounded = max(floor, int(v + divisor / 2) // divisor * divisor)
    return rounded + divisor if rounded < 0.9 * v else rounded


def _resolve(torch_weights: str) -> str:
    if os.path.exists(torch_weights):
        return torch_weights
    cached = os.path.join(_TEMP_DIR, os.path.basename(torch_weights))
    if os.path.exists(cached):
        logging.info(f"Downloaded file exists at {cached}. Using the cached file!")
        return cached
    os.makedirs(_TEMP_DIR, exist_ok=True)
    torch.hub.download_url_to_file(torch_weights, cached)  # needs network, as in the reference
    return cached


def _is_weight_leaf(leaf) -> bool:
    # utils.py:193-195: an array leaf that is not a boolean scalar
    return isinstance(leaf, torch.Tensor) and not (leaf.numel() == 1 and leaf.dtype == torch.bool)


def load_torch_weights(model: nn.Module, torch_weights: str = None) -> nn.Module:
    """Returns a copy of `model` whose array leaves / BatchNorm statistics come from the checkpoint."""
    if torch_weights is None:
        raise ValueError("torch_weights parameter cannot be empty!")
    saved = torch.load(_resolve(torch_weights), map_location="cpu")
    weights = iter([(k, v.detach()) for k, v in saved.items()
                    if "running" not in k and "num_batches" not in k])
    means = [v.detach().float() for k, v in saved.items() if "running_mean" in k]
    variances = [v.detach().float() for k, v in saved.items() if "running_var" in k]
    stats = iter(list(zip(means, variances)))

    def replace(leaf):
        if _is_weight_leaf(leaf):
            _, new = next(weights)  # StopIteration if the checkpoint is too short (as the reference)
            new = new.reshape(leaf.shape)
            return new.float() if new.is_floating_point() else new
        if isinstance(leaf, nn.StateIndex):
            return nn.StateIndex(leaf.value)  # fresh slot: the input model stays untouched
        return leaf

    model = nn.tree_map_leaves(model, replace)

    # BatchNorm statistics: two StateIndex leaves per BatchNorm in tree order (utils.py:203-218);
    # the first (first_time_index) is set to False, the second receives (running_mean, running_var).
    expect_flag = True
    for leaf in nn.tree_leaves(model):
        if isinstance(leaf, nn.StateIndex):
            if expect_flag:
                leaf.value = False
            else:
                leaf.value = next(stats)
            expect_flag = not expect_flag
    return model
