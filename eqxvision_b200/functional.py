"""Activation functions and array helpers the models are written with (stand-ins for `jax.nn.*`
and the handful of `jax.numpy` calls in the reference model files).

Each activation carries an `act_name` so that the tracer can fold it into the producing kernel's
epilogue (K15 in SURVEY.md: activations are never standalone kernels when a conv/linear precedes).
Semantics follow jax.nn: `gelu` is the tanh approximation (jax default `approximate=True`, used by
vit.py:96 and mlps.py:62), `hard_sigmoid = relu6(x+3)/6`, `hard_swish = x*hard_sigmoid(x)`.
"""
from __future__ import annotations

from . import _trace as T


def _make(name: str):
    def fn(x, *, key=None):
        if T.is_sym(x):
            return T.activation(x, name)
        raise TypeError(
            f"{name}: expected a symbolic activation; eqxvision_b200 activations only run inside a "
            "model forward (there is no CPU/eager fallback)")

    fn.act_name = name
    fn.__name__ = name
    fn.__qualname__ = name
    return fn


relu = _make("relu")
relu6 = _make("relu6")
silu = _make("silu")
swish = silu
gelu = _make("gelu")
hard_swish = _make("hard_swish")
hard_sigmoid = _make("hard_sigmoid")
sigmoid = _make("sigmoid")


def identity(x, *, key=None):
    return x


identity.act_name = None

# graph helpers re-exported for model code
ravel = T.ravel
concat_channels = T.concat_channels
split_channels = T.split_channels
channel_shuffle = T.channel_shuffle
resize_bilinear = T.resize_bilinear
prepend_cls_add_pos = T.prepend_cls_add_pos
attention = T.attention
to_tokens = T.to_tokens
to_map = T.to_map
window_attention = T.window_attention
patch_merge = T.patch_merge
