"""Module system and the building-block layers the reference takes from Equinox.

The reference models are `equinox.Module` pytrees built from `equinox.nn.*` and
`equinox.experimental.BatchNorm` (third-party, not in the reference tree; semantics restated in
SURVEY.md §8(c)-S). What matters for a drop-in is kept here:

* field ORDER: `load_torch_weights` (utils.py:189-201) matches checkpoint tensors to pytree leaves
  purely by position, so every class below declares its array fields in Equinox's order
  (Conv: weight, bias; Linear: weight, bias; LayerNorm: weight, bias; BatchNorm: weight, bias,
  first_time_index, state_index);
* constructor signatures and defaults of Equinox 0.7-0.10;
* `__call__(x, *, key=None)` on ONE sample. Called on a symbolic value the layers record graph
  nodes (`_trace.py`); called on a real array they hand the whole module to the engine, which
  traces it once per shape and replays CUDA kernels.

Parameters are host fp32 `torch.Tensor`s (the master copy); the engine packs them into device
layouts (bf16, K-major, BatchNorm folded) when a plan is built.
"""
from __future__ import annotations

import copy
import math
from typing import Any, Callable, Iterator, List, Optional, Sequence, Tuple, Union

import torch

from . import _trace as T
from . import random as jrandom


# ------------------------------------------------------------------------------------------------
# Module base / pytree utilities
# ------------------------------------------------------------------------------------------------


def _is_concrete_array(x) -> bool:
    if isinstance(x, torch.Tensor):
        return True
    return type(x).__module__.split(".")[0] in ("numpy", "jaxlib", "jax") and hasattr(x, "shape")


def entrypoint(fn: Callable) -> Callable:
    """Marks a per-sample method as callable on real arrays (dispatches to the engine)."""

    def wrapped(self, x, *args, **kwargs):
        if T.is_sym(x):
            return fn(self, x, *args, **kwargs)
        if _is_concrete_array(x):
            from . import _engine

            return _engine.run_single(self, fn.__name__, x, args, kwargs)
        return fn(self, x, *args, **kwargs)

    wrapped.__name__ = fn.__name__
    wrapped.__qualname__ = getattr(fn, "__qualname__", fn.__name__)
    wrapped.__doc__ = fn.__doc__
    wrapped.__wrapped__ = fn
    return wrapped


class Module:
    """Ordered-field container (the role `equinox.Module` plays in the reference)."""

    _fields: Tuple[str, ...] = ()

    def __init_subclass__(cls, **kwargs):
        super().__init_subclass__(**kwargs)
        fields: List[str] = []
        for base in reversed(cls.__mro__):
            for name in base.__dict__.get("__annotations__", {}):
                if name not in fields:
                    fields.append(name)
        cls._fields = tuple(fields)
        call = cls.__dict__.get("__call__")
        if call is not None and not hasattr(call, "__wrapped__"):
            cls.__call__ = entrypoint(call)

    def __repr__(self):
        inner = ", ".join(f"{f}={_short(getattr(self, f, None))}" for f in self._fields)
        return f"{type(self).__name__}({inner})"


def _short(v):
    if isinstance(v, torch.Tensor):
        return f"f32{list(v.shape)}"
    if isinstance(v, (list, tuple)) and len(v) > 3:
        return f"[{len(v)} items]"
    return repr(v)


def tree_leaves(obj) -> Iterator[Any]:
    """Leaves in the order jax.tree_util.tree_flatten visits an Equinox module: dataclass field
    order, sequences in order, dict keys sorted."""
    if isinstance(obj, Module):
        for f in obj._fields:
            yield from tree_leaves(getattr(obj, f, None))
    elif isinstance(obj, (list, tuple)):
        for v in obj:
            yield from tree_leaves(v)
    elif isinstance(obj, dict):
        for k in sorted(obj):
            yield from tree_leaves(obj[k])
    elif obj is None:
        return
    else:
        yield obj


def tree_map_leaves(obj, fn: Callable[[Any], Any]):
    """Structure-preserving copy with every leaf replaced by fn(leaf), visited in tree_leaves order."""
    if isinstance(obj, Module):
        new = copy.copy(obj)
        new.__dict__.pop("_eqxv_plans", None)
        for f in obj._fields:
            if f in obj.__dict__:
                object.__setattr__(new, f, tree_map_leaves(obj.__dict__[f], fn))
        return new
    if isinstance(obj, list):
        return [tree_map_leaves(v, fn) for v in obj]
    if isinstance(obj, tuple):
        return tuple(tree_map_leaves(v, fn) for v in obj)
    if isinstance(obj, dict):
        out = {}
        for k in sorted(obj):
            out[k] = tree_map_leaves(obj[k], fn)
        return out
    if obj is None:
        return None
    return fn(obj)


def is_array(x) -> bool:
    return isinstance(x, torch.Tensor)


def tree_inference(pytree, value: bool):
    """`equinox.tree_inference`: copy of the model with every `inference` flag set to `value`
    (Dropout, BatchNorm, DropPath drop_path.py:14, VisionTransformer.inference vit.py:171)."""

    def rec(obj):
        if isinstance(obj, Module):
            new = copy.copy(obj)
            new.__dict__.pop("_eqxv_plans", None)
            for f in obj._fields:
                if f in obj.__dict__:
                    v = obj.__dict__[f]
                    object.__setattr__(new, f, bool(value) if f == "inference" else rec(v))
            return new
        if isinstance(obj, list):
            return [rec(v) for v in obj]
        if isinstance(obj, tuple):
            return tuple(rec(v) for v in obj)
        if isinstance(obj, dict):
            return {k: rec(v) for k, v in obj.items()}
        return obj

    return rec(pytree)


def tree_at(where: Callable, pytree, replace=None, replace_fn: Callable = None):
    """`equinox.tree_at` for module nodes: copy of `pytree` in which the node(s) returned by
    `where(pytree)` are replaced (matched by identity), as used by deeplabv3.py:209."""
    targets = where(pytree)
    single = not isinstance(targets, (list, tuple))
    targets = [targets] if single else list(targets)
    if replace is not None:
        repl = [replace] if single else list(replace)
    else:
        repl = [replace_fn(t) for t in targets]

    def rec(obj):
        for t, r in zip(targets, repl):
            if obj is t:
                return r
        if isinstance(obj, Module):
            new = copy.copy(obj)
            new.__dict__.pop("_eqxv_plans", None)
            for f in obj._fields:
                if f in obj.__dict__:
                    object.__setattr__(new, f, rec(obj.__dict__[f]))
            return new
        if isinstance(obj, list):
            return [rec(v) for v in obj]
        if isinstance(obj, tuple):
            return tuple(rec(v) for v in obj)
        if isinstance(obj, dict):
            return {k: rec(v) for k, v in obj.items()}
        return obj

    return rec(pytree)


def _uniform_init(key, shape, fan_in) -> torch.Tensor:
    lim = 1.0 / math.sqrt(max(fan_in, 1))
    return jrandom.uniform(key, shape, -lim, lim)


def _tup2(v) -> Tuple[int, int]:
    return (v, v) if isinstance(v, int) else (int(v[0]), int(v[1]))


# ------------------------------------------------------------------------------------------------
# layers
# ------------------------------------------------------------------------------------------------


class Identity(Module):
    def __init__(self, *args, **kwargs):
        pass

    def __call__(self, x, *, key=None):
        return x


class Lambda(Module):
    fn: Callable

    def __init__(self, fn: Callable):
        self.fn = fn

    def __call__(self, x, *, key=None):
        return self.fn(x)


class Conv2d(Module):
    """equinox.nn.Conv2d: weight (O, I/groups, kh, kw), bias (O,1,1), cross-correlation with symmetric
    zero padding == lax.conv_general_dilated(..., rhs_dilation, feature_group_count)."""
    weight: torch.Tensor
    bias: Optional[torch.Tensor]
    in_channels: int
    out_channels: int
    kernel_size: Tuple[int, int]
    stride: Tuple[int, int]
    padding: Tuple[int, int]
    dilation: Tuple[int, int]
    groups: int
    use_bias: bool

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 use_bias=True, *, key=None):
        wkey, bkey = jrandom.split(key, 2)
        self.kernel_size = _tup2(kernel_size)
        self.stride, self.padding, self.dilation = _tup2(stride), _tup2(padding), _tup2(dilation)
        if in_channels % groups != 0 or out_channels % groups != 0:
            raise ValueError("in_channels and out_channels must be divisible by groups")
        fan_in = in_channels // groups * self.kernel_size[0] * self.kernel_size[1]
        self.weight = _uniform_init(wkey, (out_channels, in_channels // groups) + self.kernel_size, fan_in)
        self.bias = _uniform_init(bkey, (out_channels, 1, 1), fan_in) if use_bias else None
        self.in_channels, self.out_channels = in_channels, out_channels
        self.groups, self.use_bias = groups, use_bias

    def __call__(self, x, *, key=None):
        return T.conv2d(x, self.weight, self.bias, self.stride, self.padding, self.dilation, self.groups)


class Linear(Module):
    """equinox.nn.Linear on a 1-D vector: weight (out, in), bias (out,). On a token matrix it is
    applied row-wise (what `jax.vmap(self.qkv)(x)` does in vit.py:64)."""
    weight: torch.Tensor
    bias: Optional[torch.Tensor]
    in_features: int
    out_features: int
    use_bias: bool

    def __init__(self, in_features, out_features, use_bias=True, *, key=None):
        wkey, bkey = jrandom.split(key, 2)
        self.weight = _uniform_init(wkey, (out_features, in_features), in_features)
        self.bias = _uniform_init(bkey, (out_features,), in_features) if use_bias else None
        self.in_features, self.out_features, self.use_bias = in_features, out_features, use_bias

    def __call__(self, x, *, key=None):
        return T.linear(x, self.weight, self.bias)


class LayerNorm(Module):
    """equinox.nn.LayerNorm(shape, eps=1e-5): biased variance over the last dim."""
    shape: Tuple[int, ...]
    eps: float
    elementwise_affine: bool
    weight: Optional[torch.Tensor]
    bias: Optional[torch.Tensor]

    def __init__(self, shape, eps=1e-5, elementwise_affine=True, **kwargs):
        self.shape = (shape,) if isinstance(shape, int) else tuple(shape)
        if len(self.shape) != 1:
            raise NotImplementedError("LayerNorm over more than one axis is not on the hot path")
        self.eps, self.elementwise_affine = eps, elementwise_affine
        self.weight = torch.ones(self.shape) if elementwise_affine else None
        self.bias = torch.zeros(self.shape) if elementwise_affine else None

    def __call__(self, x, *, key=None):
        w = self.weight if self.weight is not None else torch.ones(self.shape)
        b = self.bias if self.bias is not None else torch.zeros(self.shape)
        return T.layer_norm(x, w, b, self.eps)


class StateIndex:
    """Stand-in for equinox.experimental.StateIndex: holds BatchNorm's running statistics, which the
    reference keeps outside the pytree (utils.py:203-218 writes them with set_state)."""

    def __init__(self, value=None):
        self.value = value

    def __repr__(self):
        return "StateIndex()"


class BatchNorm(Module):
    """equinox.experimental.BatchNorm, inference mode only:
    y = (x - mean_c) / sqrt(var_c + eps) * weight_c + bias_c with the running statistics."""
    weight: Optional[torch.Tensor]
    bias: Optional[torch.Tensor]
    first_time_index: StateIndex
    state_index: StateIndex
    axis_name: Any
    inference: bool
    input_size: int
    eps: float
    channelwise_affine: bool
    momentum: float

    def __init__(self, input_size, axis_name=None, eps=1e-5, channelwise_affine=True, momentum=0.99,
                 inference=False, **kwargs):
        self.weight = torch.ones(input_size) if channelwise_affine else None
        self.bias = torch.zeros(input_size) if channelwise_affine else None
        self.first_time_index = StateIndex(True)
        self.state_index = StateIndex((torch.zeros(input_size), torch.ones(input_size)))
        self.axis_name, self.inference, self.input_size = axis_name, inference, input_size
        self.eps, self.channelwise_affine, self.momentum = eps, channelwise_affine, momentum

    @property
    def running_mean(self):
        return self.state_index.value[0]

    @property
    def running_var(self):
        return self.state_index.value[1]

    def folded(self) -> Tuple[torch.Tensor, torch.Tensor]:
        """per-channel (scale, shift) of the inference affine, fp32"""
        mean, var = self.state_index.value
        scale = torch.rsqrt(var.double() + self.eps)
        if self.weight is not None:
            scale = scale * self.weight.double()
        shift = -mean.double() * scale
        if self.bias is not None:
            shift = shift + self.bias.double()
        return scale.float(), shift.float()

    def __call__(self, x, *, key=None, inference=None):
        inf = self.inference if inference is None else inference
        if not inf:
            raise NotImplementedError(
                "training-mode BatchNorm (batch statistics + pmean) is outside the inference hot "
                "path: switch the model with eqxvision_b200.tree_inference(model, True)")
        return T.batch_norm(x, self)


class Dropout(Module):
    p: float
    inference: bool

    def __init__(self, p=0.5, inference=False, **kwargs):
        self.p, self.inference = p, inference

    def __call__(self, x, *, key=None, inference=None):
        inf = self.inference if inference is None else inference
        if inf or self.p == 0:
            return x
        raise NotImplementedError(
            "training-mode Dropout is outside the inference hot path: switch the model with "
            "eqxvision_b200.tree_inference(model, True)")


class MaxPool2d(Module):
    kernel_size: Tuple[int, int]
    stride: Tuple[int, int]
    padding: Tuple[int, int]
    use_ceil: bool

    def __init__(self, kernel_size, stride=1, padding=0, use_ceil=False, **kwargs):
        self.kernel_size, self.stride, self.padding = _tup2(kernel_size), _tup2(stride), _tup2(padding)
        self.use_ceil = bool(use_ceil)

    def __call__(self, x, *, key=None):
        return T.pool2d(x, "max", self.kernel_size, self.stride, self.padding, ceil=self.use_ceil)


class AvgPool2d(Module):
    kernel_size: Tuple[int, int]
    stride: Tuple[int, int]
    padding: Tuple[int, int]
    use_ceil: bool

    def __init__(self, kernel_size, stride=1, padding=0, use_ceil=False, **kwargs):
        self.kernel_size, self.stride, self.padding = _tup2(kernel_size), _tup2(stride), _tup2(padding)
        if use_ceil or self.padding != (0, 0):
            raise NotImplementedError("padded / ceil-mode average pooling is not on the hot path")
        self.use_ceil = use_ceil

    def __call__(self, x, *, key=None):
        return T.pool2d(x, "avg", self.kernel_size, self.stride, self.padding)


class AdaptiveAvgPool2d(Module):
    target_shape: Tuple[int, int]

    def __init__(self, target_shape, **kwargs):
        self.target_shape = _tup2(target_shape)

    def __call__(self, x, *, key=None):
        return T.adaptive_avg_pool2d(x, self.target_shape)


class Sequential(Module):
    layers: tuple

    def __init__(self, layers: Sequence[Module]):
        self.layers = tuple(layers)

    def __call__(self, x, *, key=None):
        keys = [None] * len(self.layers) if key is None else jrandom.split(key, len(self.layers))
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


def vmap_rows(module: Module) -> Callable:
    """`jax.vmap(layer)(tokens)`: Linear / LayerNorm / MLP already act row-wise on token matrices."""

    def apply(x, *args, **kwargs):
        kwargs.pop("key", None)
        return module(x, *args, **kwargs) if not kwargs and not args else module(x, *args, **kwargs)

    return apply
