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


def _replace_modules(obj, targets, wrappers):
    """structure-preserving copy of `obj` with each module in `targets` (by identity) wrapped"""
    for t, w in zip(targets, wrappers):
        if obj is t:
            return w
    if isinstance(obj, nn.Module):
        import copy

        new = copy.copy(obj)
        new.__dict__.pop("_eqxv_plans", None)
        for f in obj._fields:
            if f in obj.__dict__:
                object.__setattr__(new, f, _replace_modules(obj.__dict__[f], targets, wrappers))
        return new
    if isinstance(obj, list):
        return [_replace_modules(v, targets, wrappers) for v in obj]
    if isinstance(obj, tuple):
        return tuple(_replace_modules(v, targets, wrappers) for v in obj)
    return obj


def intermediate_layer_getter(model: nn.Module, get_target_layers: Callable) -> nn.Module:
    targets = list(get_target_layers(model))
    taps = [_Tap() for _ in targets]
    if isinstance(model, nn.Sequential) and all(isinstance(t, int) for t in targets):
        layers = [(_TapWrapper(m, taps[targets.index(i)]) if i in targets else m)
                  for i, m in enumerate(model.layers)]
        wrapped = nn.Sequential(layers)
    else:
        wrappers = [_TapWrapper(t, tap) for t, tap in zip(targets, taps)]
        wrapped = _replace_modules(model, targets, wrappers)

    class IntermediateLayerGetter(nn.Module):
        model: nn.Module

        def __init__(self, model):
            self.model = model

        def __call__(self, x, *, key=None):
            out = self.model(x, key=key)
            return out, [t.data for t in taps]

    return IntermediateLayerGetter(wrapped)
