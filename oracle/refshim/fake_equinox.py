"""ORACLE (test infrastructure) - stand-ins for `equinox`, `equinox.nn`, `equinox.experimental` as the reference uses
them (Equinox 0.7-0.10: signatures, field order, inference semantics). See oracle/refshim/__init__.py.
Arrays are numpy (`fake_jax.Array`); convolution / pooling run in torch-CPU float32.
"""
from __future__ import annotations

import inspect
import types
from typing import Any, Callable, Optional, Sequence, Tuple

import numpy as np

from . import pytree
from .fake_jax import split as _split
from .fake_jax import uniform as _uniform
from .fake_jax import wrap


def _t(x):
    import torch

    return torch.from_numpy(np.ascontiguousarray(np.asarray(x, dtype=np.float32)))


def _pair(v) -> Tuple[int, int]:
    return (v, v) if isinstance(v, int) else tuple(v)


class Module:
    """dataclass-style module: fields are the class annotations; classes without their own __init__ get one that
    takes the fields in order (experimental.py:24-31 relies on it)"""

    def __init__(self, *args, **kwargs):
        names = []
        for base in reversed(type(self).__mro__):
            for n in base.__dict__.get("__annotations__", {}):
                if n not in names:
                    names.append(n)
        for n, v in zip(names, args):
            self.__dict__[n] = v
        self.__dict__.update(kwargs)

    def __repr__(self):
        return f"{type(self).__name__}(...)"


# ---- equinox.nn -----------------------------------------------------------------------------------------------
class Identity(Module):
    def __init__(self, *args, **kwargs):
        pass

    def __call__(self, x, *, key=None):
        return x


class Lambda(Module):
    fn: Callable

    def __init__(self, fn):
        self.fn = fn

    def __call__(self, x, *, key=None):
        return self.fn(x)


class Sequential(Module):
    layers: Sequence[Module]

    def __init__(self, layers):
        self.layers = list(layers)

    def __call__(self, x, *, key=None):
        keys = [None] * len(self.layers) if key is None else _split(key, len(self.layers))
        for layer, k in zip(self.layers, keys):
            x = layer(x, key=k)
        return x

    def __getitem__(self, i):
        if isinstance(i, slice):
            return Sequential(self.layers[i])
        return self.layers[i]

    def __iter__(self):
        return iter(self.layers)

    def __len__(self):
        return len(self.layers)


class Conv2d(Module):
    num_spatial_dims: int
    weight: np.ndarray
    bias: Optional[np.ndarray]
    in_channels: int
    out_channels: int
    kernel_size: Tuple[int, int]
    stride: Tuple[int, int]
    padding: Tuple[int, int]
    dilation: Tuple[int, int]
    groups: int
    use_bias: bool

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 use_bias=True, *, key=None, **kwargs):
        self.num_spatial_dims = 2
        kh, kw = _pair(kernel_size)
        lim = 1.0 / np.sqrt(max(1, in_channels // groups * kh * kw))
        k1, k2 = _split(key, 2) if key is not None else (np.array([0, 1], np.uint32), np.array([0, 2], np.uint32))
        self.weight = _uniform(k1, (out_channels, in_channels // groups, kh, kw), minval=-lim, maxval=lim)
        self.bias = _uniform(k2, (out_channels, 1, 1), minval=-lim, maxval=lim) if use_bias else None
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride = (kh, kw), _pair(stride)
        self.padding, self.dilation = _pair(padding), _pair(dilation)
        self.groups, self.use_bias = groups, use_bias

    def __call__(self, x, *, key=None):
        import torch.nn.functional as F

        y = F.conv2d(_t(x)[None], _t(self.weight), None, self.stride, self.padding, self.dilation, self.groups)[0]
        y = y.numpy()
        if self.use_bias:
            y = y + np.asarray(self.bias, dtype=np.float32)
        return wrap(y)


class Linear(Module):
    weight: np.ndarray
    bias: Optional[np.ndarray]
    in_features: int
    out_features: int
    use_bias: bool

    def __init__(self, in_features, out_features, use_bias=True, *, key=None, **kwargs):
        lim = 1.0 / np.sqrt(max(1, in_features))
        k1, k2 = _split(key, 2) if key is not None else (np.array([0, 1], np.uint32), np.array([0, 2], np.uint32))
        self.weight = _uniform(k1, (out_features, in_features), minval=-lim, maxval=lim)
        self.bias = _uniform(k2, (out_features,), minval=-lim, maxval=lim) if use_bias else None
        self.in_features, self.out_features, self.use_bias = in_features, out_features, use_bias

    def __call__(self, x, *, key=None):
        y = (_t(x) @ _t(self.weight).t()).numpy()
        if self.use_bias:
            y = y + np.asarray(self.bias, dtype=np.float32)
        return wrap(y)


class LayerNorm(Module):
    shape: Tuple[int, ...]
    eps: float
    elementwise_affine: bool
    weight: Optional[np.ndarray]
    bias: Optional[np.ndarray]

    def __init__(self, shape, eps=1e-5, elementwise_affine=True, **kwargs):
        self.shape = (shape,) if isinstance(shape, int) else tuple(shape)
        self.eps, self.elementwise_affine = eps, elementwise_affine
        self.weight = wrap(np.ones(self.shape, np.float32)) if elementwise_affine else None
        self.bias = wrap(np.zeros(self.shape, np.float32)) if elementwise_affine else None

    def __call__(self, x, *, key=None):
        x = np.asarray(x, dtype=np.float32)
        axes = tuple(range(x.ndim - len(self.shape), x.ndim))
        mean = x.mean(axes, keepdims=True)
        var = ((x - mean) ** 2).mean(axes, keepdims=True)       # biased variance (jnp.var)
        y = (x - mean) / np.sqrt(var + np.float32(self.eps))
        if self.elementwise_affine:
            y = y * np.asarray(self.weight) + np.asarray(self.bias)
        return wrap(y.astype(np.float32))


def batchable(fn) -> bool:
    """Linear / LayerNorm act on the trailing axes only, so `jax.vmap(layer)(rows)` == `layer(rows)`"""
    plain = (Linear.__call__, LayerNorm.__call__)
    if inspect.ismethod(fn):
        return fn.__func__ in plain
    return isinstance(fn, (Linear, LayerNorm)) and type(fn).__call__ in plain


class Dropout(Module):
    p: float
    inference: bool

    def __init__(self, p=0.5, inference=False, *, deterministic=None):
        self.p, self.inference = p, inference if deterministic is None else deterministic

    def __call__(self, x, *, key=None, inference=None, deterministic=None):
        inf = self.inference if inference is None else inference
        if inf or self.p == 0:
            return x
        raise NotImplementedError("training-mode Dropout: the shim only runs the inference path")


class _Pool(Module):
    kernel_size: Tuple[int, int]
    stride: Tuple[int, int]
    padding: Tuple[int, int]
    use_ceil: bool

    def __init__(self, kernel_size, stride=1, padding=0, use_ceil=False, **kwargs):
        self.kernel_size, self.stride, self.padding = _pair(kernel_size), _pair(stride), _pair(padding)
        self.use_ceil = use_ceil

    def _padded(self, x, value):
        """equinox Pool: symmetric padding, plus one extra stride on the right / bottom when use_ceil and the window
        sweep does not divide evenly (`_update_padding_for_ceil`)"""
        import torch.nn.functional as F

        pads = []
        for size, p, k, s in zip(x.shape[1:], self.padding, self.kernel_size, self.stride):
            right = p + s if (self.use_ceil and (size + 2 * p - k) % s != 0) else p
            pads.append((p, right))
        (pt, pb), (pl, pr) = pads
        return F.pad(_t(x)[None], (pl, pr, pt, pb), value=value)


class MaxPool2d(_Pool):
    def __call__(self, x, *, key=None):
        import torch.nn.functional as F

        return wrap(F.max_pool2d(self._padded(x, float("-inf")), self.kernel_size, self.stride)[0].numpy())


class AvgPool2d(_Pool):
    def __call__(self, x, *, key=None):
        import torch.nn.functional as F

        return wrap(F.avg_pool2d(self._padded(x, 0.0), self.kernel_size, self.stride)[0].numpy())


class AdaptivePool(Module):
    target_shape: Tuple[int, ...]

    def __init__(self, target_shape, num_spatial_dims=2, operation=None, **kwargs):
        self.target_shape = _pair(target_shape)


def _adaptive_mean_1d(x: np.ndarray, target: int, axis: int) -> np.ndarray:
    """equinox `_adaptive_pool1d`: `size % target` leading blocks of size // target + 1, then blocks of size // target"""
    size = x.shape[axis]
    if size < target:
        raise ValueError("adaptive pool: target larger than the input")
    head, block = size % target, size // target
    x = np.moveaxis(x, axis, -1)
    parts = []
    if head:
        parts.append(x[..., : head * (block + 1)].reshape(x.shape[:-1] + (head, block + 1)).mean(-1))
    parts.append(x[..., head * (block + 1):].reshape(x.shape[:-1] + (target - head, block)).mean(-1))
    return np.moveaxis(np.concatenate(parts, -1), -1, axis)


class AdaptiveAvgPool2d(AdaptivePool):
    def __call__(self, x, *, key=None):
        x = np.asarray(x, dtype=np.float32)
        for axis, t in zip((1, 2), self.target_shape):
            x = _adaptive_mean_1d(x, t, axis)
        return wrap(x.astype(np.float32))


# ---- equinox.experimental ------------------------------------------------------------------------------------
class StateIndex:
    """a leaf that addresses one slot of state (equinox 0.7-0.10: the state lives outside the pytree)"""

    def __init__(self, inference: bool = False):
        self._state = None


def set_state(index: StateIndex, state) -> None:
    index._state = state


def get_state(index: StateIndex, like=None):
    if index._state is None:
        raise RuntimeError("Cannot get state before it has been set")
    return index._state


class BatchNorm(Module):
    weight: Optional[np.ndarray]
    bias: Optional[np.ndarray]
    first_time_index: StateIndex
    state_index: StateIndex
    axis_name: Any
    inference: bool
    input_size: int
    eps: float
    channelwise_affine: bool
    momentum: float

    def __init__(self, input_size, axis_name, eps=1e-5, channelwise_affine=True, momentum=0.99, inference=False,
                 **kwargs):
        self.weight = wrap(np.ones((input_size,), np.float32)) if channelwise_affine else None
        self.bias = wrap(np.zeros((input_size,), np.float32)) if channelwise_affine else None
        self.first_time_index, self.state_index = StateIndex(), StateIndex()
        self.axis_name, self.inference, self.input_size = axis_name, inference, input_size
        self.eps, self.channelwise_affine, self.momentum = eps, channelwise_affine, momentum

    def __call__(self, x, *, key=None, inference=None):
        inf = self.inference if inference is None else inference
        if not inf:
            raise NotImplementedError("training-mode BatchNorm: the shim only runs the inference path")
        mean, var = get_state(self.state_index)
        x = np.asarray(x, dtype=np.float32)
        shape = (-1,) + (1,) * (x.ndim - 1)
        y = (x - np.asarray(mean).reshape(shape)) / np.sqrt(np.asarray(var).reshape(shape) + np.float32(self.eps))
        if self.channelwise_affine:
            y = np.asarray(self.weight).reshape(shape) * y + np.asarray(self.bias).reshape(shape)
        return wrap(y.astype(np.float32))


# ---- equinox top level ---------------------------------------------------------------------------------------
def tree_inference(tree, value: bool):
    def rec(x):
        node = pytree._children(x) if x is not None else None
        if node is None:
            return x
        new = pytree._rebuild(x, node[1], [rec(v) for v in node[2]])
        if isinstance(new, Module) and "inference" in new.__dict__:
            new.__dict__["inference"] = value
        return new

    return rec(tree)


def filter_jit(fun=None, **kwargs):
    return fun if fun is not None else (lambda f: f)


def modules():
    nn = types.ModuleType("equinox.nn")
    for c in (Identity, Lambda, Sequential, Conv2d, Linear, LayerNorm, Dropout, MaxPool2d, AvgPool2d, AdaptivePool,
              AdaptiveAvgPool2d):
        setattr(nn, c.__name__, c)
    nn.Conv = Conv2d
    experimental = types.ModuleType("equinox.experimental")
    experimental.BatchNorm, experimental.StateIndex = BatchNorm, StateIndex
    experimental.set_state, experimental.get_state = set_state, get_state
    eqx = types.ModuleType("equinox")
    eqx.Module, eqx.nn, eqx.experimental = Module, nn, experimental
    eqx.tree_at, eqx.tree_inference, eqx.filter_jit = pytree.tree_at, tree_inference, filter_jit
    eqx.is_array = lambda x: isinstance(x, np.ndarray)
    eqx.static_field = lambda **kw: None
    return {"equinox": eqx, "equinox.nn": nn, "equinox.experimental": experimental}
