"""ORACLE (test infrastructure) - stand-ins for the parts of jax the reference uses. See oracle/refshim/__init__.py.

Arrays are numpy arrays of the subclass `Array`, which keeps jax's default 32-bit discipline (float64 / int64
results are cast down, as jax does with x64 disabled) so that e.g. RegNet's width quantisation (regnet.py:254-262) is
evaluated in float32 exactly as under jax. Heavy operators live in fake_equinox (torch CPU, float32).
"""
from __future__ import annotations

import inspect
import types

import numpy as np


class Array(np.ndarray):
    def __array_ufunc__(self, ufunc, method, *inputs, **kwargs):
        ins = [np.asarray(i) if isinstance(i, np.ndarray) else i for i in inputs]
        if "out" in kwargs:
            kwargs["out"] = tuple(np.asarray(o) if isinstance(o, np.ndarray) else o for o in kwargs["out"])
        return wrap(getattr(ufunc, method)(*ins, **kwargs))

    def __array_function__(self, func, types_, args, kwargs):
        return wrap(func(*_unwrap(args), **_unwrap(kwargs)))

    def __getitem__(self, idx):
        return wrap(np.asarray(self).__getitem__(idx))

    def astype(self, dtype, *a, **k):
        return wrap(np.asarray(self).astype(dtype, *a, **k), downcast=False)

    def reshape(self, *shape, **k):
        return wrap(np.asarray(self).reshape(*shape, **k))

    def tolist(self):
        return np.asarray(self).tolist()

    def item(self, *a):
        return np.asarray(self).item(*a)

    @property
    def T(self):
        return wrap(np.asarray(self).T)

    @property
    def at(self):
        """jax's functional index update `x.at[idx].set(v)` (swin.py:148, 424: zeroing the k third of the qkv bias)"""
        return _At(self)


class _At:
    def __init__(self, arr):
        self.arr = arr

    def __getitem__(self, idx):
        arr = self.arr

        class _Ref:
            def set(self, value):
                out = np.array(np.asarray(arr), copy=True)
                out[_unwrap(idx)] = _unwrap(value)
                return wrap(out, downcast=False)

        return _Ref()


def _unwrap(x):
    if isinstance(x, Array):
        return np.asarray(x)
    if isinstance(x, (list, tuple)):
        return type(x)(_unwrap(i) for i in x)
    if isinstance(x, dict):
        return {k: _unwrap(v) for k, v in x.items()}
    return x


def wrap(x, downcast=True):
    if isinstance(x, np.ndarray):
        x = np.asarray(x)
        if downcast and x.dtype == np.float64:
            x = x.astype(np.float32)
        elif downcast and x.dtype == np.int64:
            x = x.astype(np.int32)
        return x.view(Array)
    if isinstance(x, np.generic):
        return wrap(np.asarray(x), downcast)
    if isinstance(x, (list, tuple)):
        return type(x)(wrap(i, downcast) for i in x)
    return x


class _Numpy(types.ModuleType):
    """jax.numpy: numpy's functions over `Array`"""

    def __init__(self):
        super().__init__("jax.numpy")
        self.ndarray = np.ndarray
        self.inf = np.inf
        self.pi = np.pi
        self.float32, self.int32, self.bool_ = np.float32, np.int32, np.bool_
        self.linalg = types.SimpleNamespace(norm=lambda *a, **k: wrap(np.linalg.norm(*_unwrap(a), **_unwrap(k))))

    def __getattr__(self, name):
        fn = getattr(np, name)
        if not callable(fn) or isinstance(fn, type):
            return fn

        def call(*a, **k):
            return wrap(fn(*_unwrap(a), **_unwrap(k)))

        call.__name__ = name
        return call

    def asarray(self, x, dtype=None):
        if hasattr(x, "detach"):
            x = x.detach().numpy()
        return wrap(np.asarray(_unwrap(x), dtype=dtype), downcast=dtype is None)

    array = asarray

    def ones(self, shape, dtype=np.float32):
        return wrap(np.ones(shape, dtype))

    def zeros(self, shape, dtype=np.float32):
        return wrap(np.zeros(shape, dtype))

    def arange(self, *a, **k):
        return wrap(np.arange(*a, **k))


jnp = _Numpy()


# ---- jax.nn -----------------------------------------------------------------------------------------------
def _f(x):
    return np.asarray(x, dtype=np.float32)


def relu(x):
    return wrap(np.maximum(_f(x), 0))


def relu6(x):
    return wrap(np.clip(_f(x), 0, 6))


def sigmoid(x):
    return wrap(1.0 / (1.0 + np.exp(-_f(x))))


def silu(x):
    x = _f(x)
    return wrap(x / (1.0 + np.exp(-x)))


def gelu(x, approximate=True):
    x = _f(x)
    if approximate:  # jax default
        return wrap(0.5 * x * (1.0 + np.tanh(np.float32(np.sqrt(2.0 / np.pi)) * (x + np.float32(0.044715) * x ** 3))))
    import math

    return wrap(0.5 * x * (1.0 + np.vectorize(math.erf)(x / np.sqrt(2.0)).astype(np.float32)))


def hard_sigmoid(x):
    return wrap(np.clip(_f(x) + 3.0, 0, 6) / np.float32(6.0))


def hard_swish(x):
    x = _f(x)
    return wrap(x * (np.clip(x + 3.0, 0, 6) / np.float32(6.0)))


def softmax(x, axis=-1):
    x = _f(x)
    e = np.exp(x - x.max(axis=axis, keepdims=True))
    return wrap(e / e.sum(axis=axis, keepdims=True))


# ---- jax.random -------------------------------------------------------------------------------------------
class KeyArray(Array):
    """a stack of keys; integer indexing clamps like jax does for static out-of-range indices (the reference relies
    on it: googlenet.py:201,224 splits 5 keys and uses keys[5])"""

    def __getitem__(self, idx):
        if isinstance(idx, int) and self.ndim == 2:
            idx = min(idx, self.shape[0] - 1)
        out = np.asarray(self).__getitem__(idx)
        return out.view(KeyArray) if isinstance(out, np.ndarray) else out

    def __iter__(self):   # explicit: the sequence-protocol fallback would never see an IndexError from the clamp
        base = np.asarray(self)
        return iter([base[i].view(KeyArray) for i in range(base.shape[0])])

    def __len__(self):
        return np.asarray(self).shape[0]


def PRNGKey(seed):
    return np.array([0, int(seed) & 0xFFFFFFFF], dtype=np.uint32).view(KeyArray)


def _rng(key):
    k = np.asarray(key).astype(np.uint64).reshape(-1)
    return np.random.default_rng(int(k[0]) * 4294967296 + int(k[-1]))


def split(key, num=2):
    r = _rng(key)
    return r.integers(0, 2 ** 32, size=(num, 2), dtype=np.uint64).astype(np.uint32).view(KeyArray)


def uniform(key, shape=(), dtype=np.float32, minval=0.0, maxval=1.0):
    return wrap(_rng(key).uniform(minval, maxval, size=shape).astype(np.float32))


def normal(key, shape=(), dtype=np.float32):
    return wrap(_rng(key).standard_normal(size=shape).astype(np.float32))


def truncated_normal(key, lower, upper, shape=(), dtype=np.float32):
    lo, hi = min(lower, upper), max(lower, upper)
    return wrap(np.clip(_rng(key).standard_normal(size=shape), lo, hi).astype(np.float32))


def bernoulli(key, p=0.5, shape=()):
    return wrap(_rng(key).uniform(size=shape) < p)


# ---- jax.vmap / tree_util / image / lax -----------------------------------------------------------------------
def _is_none(x):
    return x is None


def vmap(fun, in_axes=0, out_axes=0, axis_name=None):
    """maps over axis 0 of every positional and keyword array argument (all the reference ever asks for)"""
    from . import fake_equinox as fe

    if fe.batchable(fun):            # Linear / LayerNorm broadcast over leading axes: same numbers, no Python loop
        return fun

    def mapped(*args, **kwargs):
        n = next(a.shape[0] for a in list(args) + list(kwargs.values()) if isinstance(a, np.ndarray))
        outs = [fun(*[a[i] if isinstance(a, np.ndarray) else a for a in args],
                    **{k: (v[i] if isinstance(v, np.ndarray) else v) for k, v in kwargs.items()}) for i in range(n)]
        return _stack(outs)

    return mapped


def _stack(outs):
    first = outs[0]
    if first is None:
        return None
    if isinstance(first, (tuple, list)):
        return type(first)(_stack([o[j] for o in outs]) for j in range(len(first)))
    return wrap(np.stack([np.asarray(o) for o in outs]))


def resize(image, shape, method="bilinear", antialias=True):
    """jax.image.resize for the upsampling the reference does (segmentation/_utils.py:52, deeplabv3.py:74): half-pixel
    centres, edge weights renormalised == F.interpolate(align_corners=False) (asserted at 1e-4 by the reference's own
    test_deeplabv3.py:27)"""
    import torch
    import torch.nn.functional as F

    x = torch.from_numpy(np.ascontiguousarray(np.asarray(image, dtype=np.float32)))
    if method not in ("bilinear", "linear"):
        raise NotImplementedError(method)
    assert x.dim() == 3 and tuple(shape)[0] == x.shape[0]
    y = F.interpolate(x[None], size=tuple(shape)[1:], mode="bilinear", align_corners=False)[0]
    return wrap(y.numpy())


def modules():
    import types as _t

    tree_util = _t.ModuleType("jax.tree_util")
    from . import pytree

    tree_util.tree_flatten = pytree.tree_flatten
    tree_util.tree_unflatten = pytree.tree_unflatten
    tree_util.tree_map = pytree.tree_map
    tree_util.tree_leaves = pytree.tree_leaves
    nn = _t.ModuleType("jax.nn")
    for f in (relu, relu6, sigmoid, silu, gelu, hard_sigmoid, hard_swish, softmax):
        setattr(nn, f.__name__, f)
    nn.swish = silu
    random = _t.ModuleType("jax.random")
    for f in (PRNGKey, split, uniform, normal, truncated_normal, bernoulli):
        setattr(random, f.__name__, f)
    image = _t.ModuleType("jax.image")
    image.resize = resize
    lax = _t.ModuleType("jax.lax")

    def scan(f, init, xs, length=None):
        """jax.lax.scan as a Python loop (swin.py:147-151, 423-427 use it to zero a bias slice element by element)"""
        carry, ys = init, []
        for x in (range(length) if xs is None else xs):
            carry, y = f(carry, x)
            ys.append(y)
        return carry, (wrap(np.stack([np.asarray(y) for y in ys])) if ys and ys[0] is not None else None)

    def clamp(min, x, max):   # noqa: A002  (jax.lax.clamp(min, x, max), called with keywords at swin.py:165)
        return wrap(np.clip(_unwrap(x), _unwrap(min), _unwrap(max)))

    lax.scan, lax.clamp = scan, clamp
    jax = _t.ModuleType("jax")
    jax.numpy, jax.nn, jax.random, jax.image, jax.lax, jax.tree_util = jnp, nn, random, image, lax, tree_util
    jax.vmap = vmap
    jax.jit = lambda f, *a, **k: f
    jax.Array = np.ndarray
    return {"jax": jax, "jax.numpy": jnp, "jax.nn": nn, "jax.random": random, "jax.image": image, "jax.lax": lax,
            "jax.tree_util": tree_util}
