"""Symbolic per-sample values and the lazy expression DAG the models are written against.

The reference models are per-sample functions batched by `jax.vmap` and traced once by
`eqx.filter_jit` (README.md:37-46). Here a model's `__call__` runs once on a `Sym` (a per-sample
shape plus an expression node); nothing is computed while tracing. Conv/Linear nodes absorb the
BatchNorm, activation and residual-add that follow them (immutably: folding returns a new node), so
that what reaches the plan builder (`_engine.py`) is already the fused-kernel granularity of
libeqxv_b200.so.
"""
from __future__ import annotations

from typing import Optional, Sequence, Tuple

import torch

# ------------------------------------------------------------------------------------------------
# expression nodes
# ------------------------------------------------------------------------------------------------


class Expr:
    __slots__ = ()


class Input(Expr):
    """the fp32 NCHW image batch handed to the model"""
    __slots__ = ()


class Const(Expr):
    """a parameter tensor used as a data operand (cls_token, pos_embed, ...)"""
    __slots__ = ("value",)

    def __init__(self, value):
        self.value = value


class Conv(Expr):
    __slots__ = ("x", "weight", "bias", "stride", "padding", "dilation", "groups", "bn", "act1", "res", "act2")

    def __init__(self, x, weight, bias, stride, padding, dilation, groups, bn=None, act1=None, res=None,
                 act2=None):
        self.x, self.weight, self.bias = x, weight, bias
        self.stride, self.padding, self.dilation, self.groups = stride, padding, dilation, groups
        self.bn, self.act1, self.res, self.act2 = bn, act1, res, act2

    def replace(self, **kw):
        d = {k: getattr(self, k) for k in self.__slots__}
        d.update(kw)
        return Conv(**d)


class Linear(Expr):
    __slots__ = ("x", "weight", "bias", "act1", "res", "act2")

    def __init__(self, x, weight, bias, act1=None, res=None, act2=None):
        self.x, self.weight, self.bias, self.act1, self.res, self.act2 = x, weight, bias, act1, res, act2

    def replace(self, **kw):
        d = {k: getattr(self, k) for k in self.__slots__}
        d.update(kw)
        return Linear(**d)


class BNAct(Expr):  # standalone per-channel affine (+activation): densenet.py:64-65
    __slots__ = ("x", "bn", "act")

    def __init__(self, x, bn, act=None):
        self.x, self.bn, self.act = x, bn, act


class Act(Expr):
    __slots__ = ("x", "act")

    def __init__(self, x, act):
        self.x, self.act = x, act


class Add(Expr):
    __slots__ = ("a", "b", "act")

    def __init__(self, a, b, act=None):
        self.a, self.b, self.act = a, b, act


class ChannelScale(Expr):  # x * s with s of shape (C,1,1): squeeze.py:61
    __slots__ = ("x", "s")

    def __init__(self, x, s):
        self.x, self.s = x, s


class Pool(Expr):
    __slots__ = ("x", "mode", "k", "stride", "pad", "ceil")

    def __init__(self, x, mode, k, stride, pad, ceil=False):
        self.x, self.mode, self.k, self.stride, self.pad, self.ceil = x, mode, k, stride, pad, ceil


class AdaptiveAvgPool(Expr):
    __slots__ = ("x", "oh", "ow")

    def __init__(self, x, oh, ow):
        self.x, self.oh, self.ow = x, oh, ow


class Ravel(Expr):  # (C,H,W) -> (C*H*W,) in C,H,W order (jnp.ravel)
    __slots__ = ("x",)

    def __init__(self, x):
        self.x = x


class ToTokens(Expr):  # (C,H,W) -> (H*W, C): vmap(ravel) + moveaxis, patch_embed.py:81-82
    __slots__ = ("x",)

    def __init__(self, x):
        self.x = x


class ToMap(Expr):  # (H*W, C) -> (C,H,W): extensions_2d.py:27-28
    __slots__ = ("x", "h", "w")

    def __init__(self, x, h, w):
        self.x, self.h, self.w = x, h, w


class LayerNormE(Expr):
    __slots__ = ("x", "weight", "bias", "eps")

    def __init__(self, x, weight, bias, eps):
        self.x, self.weight, self.bias, self.eps = x, weight, bias, eps


class Attention(Expr):  # vit.py:62-73
    __slots__ = ("qkv", "heads", "scale")

    def __init__(self, qkv, heads, scale):
        self.qkv, self.heads, self.scale = qkv, heads, scale


class AttentionProbs(Expr):  # the (1, heads, T, T) softmax matrix, vit.py:70 / 151-152
    __slots__ = ("qkv", "heads", "scale")

    def __init__(self, qkv, heads, scale):
        self.qkv, self.heads, self.scale = qkv, heads, scale


class ClsPos(Expr):  # concat([cls, x]) + pos_embed, vit.py:269
    __slots__ = ("x", "cls", "pos")

    def __init__(self, x, cls, pos):
        self.x, self.cls, self.pos = x, cls, pos


class SelectRow(Expr):
    __slots__ = ("x", "row")

    def __init__(self, x, row):
        self.x, self.row = x, row


class Concat(Expr):  # channel concatenation, densenet.py:63, deeplabv3.py:134
    __slots__ = ("xs", "capacity", "align")

    def __init__(self, xs, capacity=None, align=1):
        self.xs = tuple(xs)
        self.capacity = capacity  # hint: final width of the buffer this concat will grow into
        self.align = align        # every member starts at a multiple of `align` channels (pad channels stay zero)


class ChannelView(Expr):  # lazy channel gather: out[j] = x[idx[j]] (jnp.split / _channel_shuffle, shufflenetv2.py:14-20,134)
    __slots__ = ("x", "idx")

    def __init__(self, x, idx):
        self.x, self.idx = x, tuple(int(i) for i in idx)


class Resize(Expr):  # jax.image.resize(method="bilinear"), _utils.py:52
    __slots__ = ("x", "h", "w")

    def __init__(self, x, h, w):
        self.x, self.h, self.w = x, h, w


class WindowAttention(Expr):  # swin.py:90-255 between the qkv and proj matmuls
    __slots__ = ("qkv", "h", "w", "heads", "window", "shift", "bias", "scale", "cosine_scale")

    def __init__(self, qkv, h, w, heads, window, shift, bias, scale, cosine_scale=None):
        self.qkv, self.h, self.w, self.heads = qkv, h, w, heads
        self.window, self.shift, self.bias, self.scale = window, shift, bias, scale
        self.cosine_scale = cosine_scale   # Swin-V2: per-head exp(min(logit_scale, log 100)); q, k normalised first


class PatchMerge(Expr):  # swin.py:23-33: 2x2 strided gather, channels (x0,x1,x2,x3)
    __slots__ = ("x",)

    def __init__(self, x):
        self.x = x


# ------------------------------------------------------------------------------------------------
# symbolic value
# ------------------------------------------------------------------------------------------------


class Sym:
    """A per-sample array placeholder. kind: 'chw' (C,H,W), 'tokens' (T,D), 'vec' (D,), 'attn'."""
    __slots__ = ("kind", "shape", "expr", "__weakref__")

    def __init__(self, kind: str, shape: Tuple[int, ...], expr: Expr):
        self.kind, self.shape, self.expr = kind, tuple(int(s) for s in shape), expr

    # -- arithmetic the reference models use on activations --
    def __add__(self, other):
        return add(self, other)

    __radd__ = __add__

    def __iadd__(self, other):  # `out += identity` (resnet.py:159): Syms are immutable
        return add(self, other)

    def __mul__(self, other):
        return mul(self, other)

    __rmul__ = __mul__

    def __getitem__(self, idx):
        if self.kind == "tokens" and isinstance(idx, int):
            return select_row(self, idx)
        raise TypeError(f"unsupported index {idx!r} on a {self.kind} Sym")

    @property
    def ndim(self):
        return len(self.shape)

    def __repr__(self):
        return f"Sym({self.kind}{self.shape} <- {type(self.expr).__name__})"


def is_sym(x) -> bool:
    return isinstance(x, Sym)


def _act_name(fn) -> Optional[str]:
    if fn is None:
        return None
    if isinstance(fn, str):
        return fn
    name = getattr(fn, "act_name", None)
    if name is None:
        raise TypeError(f"{fn!r} is not an eqxvision_b200 activation (use eqxvision_b200.functional.*)")
    return name


# ------------------------------------------------------------------------------------------------
# graph-building primitives (used by eqxvision_b200.nn / functional)
# ------------------------------------------------------------------------------------------------


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def conv2d(x: Sym, weight, bias, stride, padding, dilation, groups) -> Sym:
    if x.kind != "chw":
        raise ValueError(f"Conv2d expects a (C,H,W) input, got {x}")
    cout, cin_g, kh, kw = weight.shape
    c, h, w = x.shape
    if c != cin_g * groups:
        raise ValueError(f"Conv2d: input has {c} channels, weight expects {cin_g * groups}")
    if isinstance(x.expr, ChannelView) and groups == 1:
        # a dense convolution over gathered channels == the same convolution over the underlying buffer with the
        # filter's input channels scattered to their physical positions (zeros elsewhere): the gather costs nothing
        base = x.expr.x
        wide = torch.zeros((cout, base.shape[0], kh, kw), dtype=torch.float32)
        wide.index_add_(1, torch.tensor(x.expr.idx, dtype=torch.long), weight.detach().float())
        x, weight, c = base, wide, base.shape[0]
    (sh, sw), (ph, pw), (dh, dw) = _pair(stride), _pair(padding), _pair(dilation)
    ho = (h + 2 * ph - dh * (kh - 1) - 1) // sh + 1
    wo = (w + 2 * pw - dw * (kw - 1) - 1) // sw + 1
    return Sym("chw", (cout, ho, wo), Conv(x, weight, bias, (sh, sw), (ph, pw), (dh, dw), groups))


def linear(x: Sym, weight, bias) -> Sym:
    out_f, in_f = weight.shape
    if x.kind == "vec":
        if x.shape[0] != in_f:
            raise ValueError(f"Linear: expected {in_f} features, got {x.shape[0]}")
        return Sym("vec", (out_f,), Linear(x, weight, bias))
    if x.kind == "tokens":
        if x.shape[1] != in_f:
            raise ValueError(f"Linear: expected {in_f} features, got {x.shape[1]}")
        return Sym("tokens", (x.shape[0], out_f), Linear(x, weight, bias))
    raise ValueError(f"Linear expects a vector or token matrix, got {x}")


def batch_norm(x: Sym, bn) -> Sym:
    e = x.expr
    if isinstance(e, Conv) and e.bn is None and e.act1 is None and e.res is None and e.act2 is None:
        return Sym(x.kind, x.shape, e.replace(bn=bn))
    return Sym(x.kind, x.shape, BNAct(x, bn))


def _under_view(x: Sym):
    """(C,H,W) <-> (H*W,C) re-interpretations share one channels-last buffer (extensions_2d.py): look
    through them so that an activation / residual add still folds into the producing GEMM."""
    e = x.expr
    if isinstance(e, ToMap):
        return e.x, lambda y: to_map(y, e.h, e.w)
    return None, None


def activation(x: Sym, fn) -> Sym:
    name = _act_name(fn)
    if name is None:
        return x
    inner, rewrap = _under_view(x)
    if inner is not None and isinstance(inner.expr, Linear):
        return rewrap(activation(inner, name))
    e = x.expr
    if isinstance(e, (Conv, Linear)):
        if e.res is None and e.act1 is None and e.act2 is None:
            return Sym(x.kind, x.shape, e.replace(act1=name))
        if e.res is not None and e.act1 is None and e.act2 is None:
            return Sym(x.kind, x.shape, e.replace(act2=name))
    if isinstance(e, BNAct) and e.act is None:
        return Sym(x.kind, x.shape, BNAct(e.x, e.bn, name))
    if isinstance(e, Add) and e.act is None:
        return Sym(x.kind, x.shape, Add(e.a, e.b, name))
    return Sym(x.kind, x.shape, Act(x, name))


def add(a, b) -> Sym:
    if not is_sym(a) or not is_sym(b):
        raise TypeError("add: both operands must be symbolic activations")
    if a.shape != b.shape:
        raise ValueError(f"add: shape mismatch {a.shape} vs {b.shape}")
    for p, q in ((a, b), (b, a)):
        inner, rewrap = _under_view(p)
        if inner is not None and isinstance(inner.expr, Linear) and inner.expr.res is None \
                and inner.expr.act2 is None and q.kind == "chw":
            return rewrap(add(inner, to_tokens(q)))
    # `x + block(x)` (mobilenetv2.py:86, regnet.py:166): x is an input of the other operand, so folding the add into
    # x's producer would compute that producer twice (once plain for block(x), once with the residual); fold into
    # the operand that is not an ancestor of the other one
    order = ((b, a), (a, b)) if _feeds(a, b) else ((a, b), (b, a))
    for p, q in order:
        e = p.expr
        if isinstance(e, (Conv, Linear)) and e.res is None and e.act2 is None:
            return Sym(p.kind, p.shape, e.replace(res=q))
    return Sym(a.kind, a.shape, Add(a, b))


def _feeds(p: Sym, q: Sym, limit: int = 64) -> bool:
    """True when p's expression is reachable from q through at most `limit` expression nodes"""
    target = p.expr
    stack, seen = [q.expr], 0
    while stack and seen < limit:
        e = stack.pop()
        seen += 1
        if e is target:
            return True
        for name in getattr(e, "__slots__", ()):
            v = getattr(e, name, None)
            if isinstance(v, Sym):
                stack.append(v.expr)
            elif isinstance(v, tuple):
                stack.extend(i.expr for i in v if isinstance(i, Sym))
    return False


def _scale_channels(x: Sym, s) -> Optional[Sym]:
    """per-channel constant gain `s` (C,1,1) on a map produced by a Linear2d (ConvNeXt layer scale, convnext.py:62):
    folded into that Linear's rows on the host, so the GEMM epilogue (and the residual add that follows) is unchanged"""
    inner, rewrap = _under_view(x)
    if inner is None or not isinstance(inner.expr, Linear):
        return None
    e = inner.expr
    if e.act1 is not None or e.res is not None or e.act2 is not None:
        return None
    g = s.detach().float().reshape(-1)
    if g.numel() != e.weight.shape[0]:
        return None
    w = e.weight.detach().float() * g[:, None]
    b = None if e.bias is None else e.bias.detach().float().reshape(-1) * g
    return rewrap(Sym(inner.kind, inner.shape, e.replace(weight=w, bias=b)))


def mul(a, b) -> Sym:
    for x, s in ((a, b), (b, a)):
        if is_sym(x) and isinstance(s, torch.Tensor) and x.kind == "chw" and tuple(s.shape) == (x.shape[0], 1, 1):
            y = _scale_channels(x, s)
            if y is not None:
                return y
            raise NotImplementedError("constant per-channel scale is only built behind a Linear2d (ConvNeXt layer scale)")
    if is_sym(a) and is_sym(b):
        for x, s in ((a, b), (b, a)):
            if x.kind == "chw" and s.kind == "chw" and s.shape == (x.shape[0], 1, 1):
                return Sym("chw", x.shape, ChannelScale(x, s))
    raise TypeError(f"mul: unsupported operands {a!r} * {b!r}")


def _pool_out(size: int, k: int, stride: int, pad: int, ceil: bool) -> int:
    if not ceil:
        return (size + 2 * pad - k) // stride + 1
    # use_ceil=True (squeezenet.py:84, googlenet.py:95): equinox pads the right edge by one more stride when the sweep
    # does not divide evenly; the reference pins these models against torchvision, whose ceil_mode also drops a last
    # window that would start beyond the input. The two rules agree unless that window is pure padding.
    o = -((size + 2 * pad - k) // -stride) + 1
    if (o - 1) * stride >= size + pad:
        raise NotImplementedError("ceil-mode pooling whose last window lies entirely in the padding")
    return o


def pool2d(x: Sym, mode: str, k, stride, pad, ceil: bool = False) -> Sym:
    (kh, kw), (sh, sw), (ph, pw) = _pair(k), _pair(stride), _pair(pad)
    if kh != kw or sh != sw or ph != pw:
        raise NotImplementedError("only square pooling windows are supported")
    if ceil and mode != "max":
        raise NotImplementedError("ceil-mode average pooling is not on the hot path")
    c, h, w = x.shape
    ho, wo = _pool_out(h, kh, sh, ph, ceil), _pool_out(w, kw, sw, pw, ceil)
    return Sym("chw", (c, ho, wo), Pool(x, mode, kh, sh, ph, bool(ceil)))


def adaptive_avg_pool2d(x: Sym, target) -> Sym:
    oh, ow = _pair(target)
    c, h, w = x.shape
    if (oh, ow) == (h, w):
        return x
    if (h % oh or w % ow) and (oh > h or ow > w):
        # equinox splits an uneven axis into `dim % t` leading blocks of dim // t + 1 and the rest of dim // t (the
        # device kernel follows that rule); a target larger than the map would need empty blocks
        raise NotImplementedError(f"adaptive average pooling {h}x{w} -> {oh}x{ow}: target larger than the map")
    return Sym("chw", (c, oh, ow), AdaptiveAvgPool(x, oh, ow))


def ravel(x: Sym) -> Sym:
    if x.kind == "vec":
        return x
    if x.kind != "chw":
        raise ValueError(f"ravel expects (C,H,W), got {x}")
    c, h, w = x.shape
    return Sym("vec", (c * h * w,), Ravel(x))


def to_tokens(x: Sym) -> Sym:
    c, h, w = x.shape
    if isinstance(x.expr, ToMap):  # (T,C) -> (C,H,W) -> (T,C): the same buffer
        return x.expr.x
    return Sym("tokens", (h * w, c), ToTokens(x))


def to_map(x: Sym, h: int, w: int) -> Sym:
    t, d = x.shape
    assert t == h * w
    if isinstance(x.expr, ToTokens) and x.expr.x.shape == (d, h, w):
        return x.expr.x
    return Sym("chw", (d, h, w), ToMap(x, h, w))


def layer_norm(x: Sym, weight, bias, eps) -> Sym:
    if x.kind not in ("tokens", "vec"):
        raise ValueError(f"LayerNorm expects tokens or a vector, got {x}")
    return Sym(x.kind, x.shape, LayerNormE(x, weight, bias, eps))


def attention(qkv: Sym, heads: int, scale: float):
    t, c3 = qkv.shape
    c = c3 // 3
    out = Sym("tokens", (t, c), Attention(qkv, heads, scale))
    probs = Sym("attn", (1, heads, t, t), AttentionProbs(qkv, heads, scale))
    return out, probs


def prepend_cls_add_pos(x: Sym, cls_token, pos_embed) -> Sym:
    t, d = x.shape
    if tuple(pos_embed.shape) != (t + 1, d) or tuple(cls_token.shape) != (1, d):
        raise ValueError("cls_token / pos_embed shape mismatch")
    return Sym("tokens", (t + 1, d), ClsPos(x, cls_token, pos_embed))


def select_row(x: Sym, row: int) -> Sym:
    t, d = x.shape
    if row < 0:
        row += t
    e = x.expr
    if isinstance(e, LayerNormE):  # LayerNorm is row-wise: normalise only the selected row
        return Sym("vec", (d,), LayerNormE(select_row(e.x, row), e.weight, e.bias, e.eps))
    return Sym("vec", (d,), SelectRow(x, row))


def concat_channels(xs: Sequence[Sym], capacity: Optional[int] = None) -> Sym:
    xs = list(xs)
    if len(xs) == 1 and capacity is None:
        return xs[0]
    h, w = xs[0].shape[1:]
    for x in xs:
        if x.kind != "chw" or x.shape[1:] != (h, w):
            raise ValueError("concat_channels: spatial shape mismatch")
    widths = [x.shape[0] for x in xs]
    if capacity is None and any(c % 8 for c in widths[:-1]):
        # members that do not end on an 8-channel (16-byte) boundary (ShuffleNetV2 x1.0: 58 + 58): each member gets
        # an aligned slot in a wider buffer whose pad channels stay zero; the logical tensor is a view that skips them
        idx, off = [], 0
        for c in widths:
            idx.extend(range(off, off + c))
            off += (c + 7) // 8 * 8
        phys = Sym("chw", (off - ((widths[-1] + 7) // 8 * 8) + widths[-1], h, w), Concat(xs, None, align=8))
        return channel_view(phys, idx)
    return Sym("chw", (sum(widths), h, w), Concat(xs, capacity))


def channel_view(x: Sym, idx) -> Sym:
    """out[j] = x[idx[j]] over the channel axis of a (C,H,W) map, lazily (composes; the identity disappears)"""
    if x.kind != "chw":
        raise ValueError("channel_view expects a (C,H,W) map")
    idx = [int(i) for i in idx]
    if isinstance(x.expr, ChannelView):
        inner = x.expr.idx
        idx = [inner[i] for i in idx]
        x = x.expr.x
    if idx == list(range(x.shape[0])):
        return x
    return Sym("chw", (len(idx),) + x.shape[1:], ChannelView(x, idx))


def split_channels(x: Sym, sections: int):
    """jnp.split(x, sections, axis=0) on a (C,H,W) map (shufflenetv2.py:134)"""
    c = x.shape[0]
    if c % sections:
        raise ValueError("array split does not result in an equal division")
    step = c // sections
    return [channel_view(x, range(i * step, (i + 1) * step)) for i in range(sections)]


def channel_shuffle(x: Sym, groups: int) -> Sym:
    """_channel_shuffle (shufflenetv2.py:14-20): (g, c/g, H, W) -> transpose(1, 0) -> flatten"""
    c = x.shape[0]
    cpg = c // groups
    return channel_view(x, [(j % groups) * cpg + j // groups for j in range(c)])


def resize_bilinear(x: Sym, h: int, w: int) -> Sym:
    c = x.shape[0]
    return Sym("chw", (c, h, w), Resize(x, h, w))


def window_attention(qkv: Sym, h: int, w: int, heads: int, window, shift, bias, scale: float,
                     cosine_scale=None) -> Sym:
    """softmax(q*scale k^T + bias + shift_mask) v per window; qkv is the (H*W, 3C) token matrix in
    spatial order, the result is the (H*W, C) matrix in the same order (swin.py:117-253)."""
    t, c3 = qkv.shape
    if t != h * w or c3 % (3 * heads) != 0:
        raise ValueError(f"window_attention: qkv {qkv.shape} does not match a {h}x{w} map with {heads} heads")
    if h % window[0] != 0 or w % window[1] != 0:
        # the reference has the padding commented out (swin.py:107-112): its reshape raises
        raise ValueError(f"feature map {h}x{w} is not a multiple of the window {tuple(window)}")
    return Sym("tokens", (t, c3 // 3), WindowAttention(qkv, h, w, heads, tuple(window), tuple(shift), bias, scale,
                                                      cosine_scale))


def patch_merge(x: Sym) -> Sym:
    c, h, w = x.shape
    if h % 2 or w % 2:
        raise NotImplementedError("patch merging of odd feature maps (zero padding, swin.py:25)")
    return Sym("chw", (4 * c, h // 2, w // 2), PatchMerge(x))


def const_like(t: torch.Tensor) -> Const:
    return Const(t)
