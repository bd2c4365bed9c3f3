"""Stand-ins for the three third-party calls every reference call site makes (README.md:37-46):

    @eqx.filter_jit
    def forward(net, images, keys):
        return jax.vmap(net, axis_name="batch")(images, key=keys)
    net = eqx.tree_inference(net, True)

`vmap(net)` returns a callable that runs the whole batch through one captured CUDA graph;
`filter_jit` is the identity (plans are cached per (module, shape) by the engine, which is what
jit's compilation cache provides); `tree_inference` flips the `inference` flags.
"""
from __future__ import annotations

from typing import Any, Callable

from . import _engine
from .nn import Module, tree_inference  # noqa: F401


def vmap(fun: Callable, in_axes: Any = 0, out_axes: Any = 0, axis_name: Any = None, **unused) -> Callable:
    """Batch a per-sample module (or bound per-sample method) over axis 0 of its first argument.
    Extra positional arguments must be unbatched (`in_axes=(0, None, ...)`), `key=` is accepted and
    ignored (dead in inference)."""
    if isinstance(in_axes, (tuple, list)):
        if in_axes[0] != 0 or any(a is not None for a in in_axes[1:]):
            raise NotImplementedError("vmap: only in_axes=(0, None, ...) is supported")
    elif in_axes != 0:
        raise NotImplementedError("vmap: only batching over axis 0 is supported")
    if isinstance(fun, Module):
        module, method = fun, "__call__"
    elif hasattr(fun, "__self__") and isinstance(fun.__self__, Module):
        module, method = fun.__self__, fun.__name__
    else:
        raise TypeError("vmap expects an eqxvision_b200 Module or one of its per-sample methods")

    def batched(x, *args, **kwargs):
        return _engine.run_batched(module, method, x, args, kwargs)

    return batched


def filter_jit(fun: Callable = None, **unused):
    if fun is None:
        return lambda f: f
    return fun
