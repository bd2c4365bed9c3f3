"""`eqxvision.experimental.intermediate_layer_getter` (reference: experimental.py:35-88).

The reference wraps chosen sub-modules so that their outputs are stashed in a Python closure at
trace time. Tracing here is symbolic as well, so the same approach works: the wrapper records the
symbolic output of the most recent call and the getter returns `(out, [taps...])`.
"""
from typing import Callable

from . import nn


class _Tap:
    def __init__(self):
        self.data = None


class _TapWrapper(nn.Module):
    layer: nn.Module

    def __init__(self, layer, tap):
        self.layer = layer
        self._tap = tap

    def __call__(self, x, *, key=None):
        out = self.layer(x, key=key)
        self._tap.data = out
        return out


def intermediate_layer_getter(model: nn.Module, get_target_layers: Callable) -> nn.Module:
    targets = list(get_target_layers(model))
    taps = [_Tap() for _ in targets]
    if isinstance(model, nn.Sequential) and all(isinstance(t, int) for t in targets):
        layers = [(_TapWrapper(m, taps[targets.index(i)]) if i in targets else m)
                  for i, m in enumerate(model.layers)]
        wrapped = nn.Sequential(layers)
    else:
        wrappers = [_TapWrapper(t, tap) for t, tap in zip(targets, taps)]
        wrapped = nn.tree_at(lambda m: targets, model, replace=wrappers)

    class IntermediateLayerGetter(nn.Module):
        model: nn.Module

        def __init__(self, model):
            self.model = model

        def __call__(self, x, *, key=None):
            out = self.model(x, key=key)
            return out, [t.data for t in taps]

    return IntermediateLayerGetter(wrapped)
