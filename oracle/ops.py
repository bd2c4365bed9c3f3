"""ORACLE (test infrastructure, not product code) — CPU fp32 restatement of the third-party ops the
reference delegates to (equinox.nn / equinox.experimental / jax.nn / jax.image; none of them is
under /root/reference).  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this package.

Parity status: the reference cannot be imported here (jax / equinox are not installed and cannot
be), so these functions are pinned (a) against torchvision on identical state_dicts at the
reference's own tolerance atol=1e-4 (tests/test_oracle.py) for the families whose reference tests
assert exactly that, and (b) by running the reference's OWN, unmodified model files on jax / equinox stand-ins
(oracle/refshim: separate implementations of the third-party ops, numpy + torch CPU) and holding oracle/models.py
to 1e-4 against their output (tests/test_refshim.py, 21 model configurations; golden vectors produced that way are
committed under tests/golden/golden_ref_v1.pt).  ViT has shape-only tests in the reference, so (b) is its pin.

Everything is plain torch on CPU in float32 (float64 where noted), written op by op.
"""
from __future__ import annotations

import math
from typing import Optional, Sequence, Tuple

import torch
import torch.nn.functional as F


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def conv2d(x, weight, bias=None, stride=1, padding=0, dilation=1, groups=1):
    """equinox.nn.Conv2d.__call__: lax.conv_general_dilated(x[None], W, stride, [(p,p),(p,p)],
    rhs_dilation=dil, feature_group_count=groups) + bias(O,1,1).  x: (N,C,H,W) batched here."""
    b = None if bias is None else bias.reshape(-1)
    return F.conv2d(x, weight, b, _pair(stride), _pair(padding), _pair(dilation), groups)


def linear(x, weight, bias=None):
    """equinox.nn.Linear: W @ x + b on the last axis"""
    y = x @ weight.t()
    return y if bias is None else y + bias


def batch_norm_inference(x, weight, bias, mean, var, eps=1e-5):
    """equinox.experimental.BatchNorm(inference): (x - mean)/sqrt(var+eps) * weight + bias, per channel"""
    shape = (1, -1) + (1,) * (x.dim() - 2)
    y = (x - mean.reshape(shape)) / torch.sqrt(var.reshape(shape) + eps)
    if weight is not None:
        y = y * weight.reshape(shape)
    if bias is not None:
        y = y + bias.reshape(shape)
    return y


def layer_norm(x, weight, bias, eps=1e-5):
    """equinox.nn.LayerNorm: biased variance over the last axis"""
    mean = x.mean(-1, keepdim=True)
    var = ((x - mean) ** 2).mean(-1, keepdim=True)
    y = (x - mean) * torch.rsqrt(var + eps)
    if weight is not None:
        y = y * weight
    if bias is not None:
        y = y + bias
    return y


def max_pool2d(x, k, stride, padding=0, ceil_mode=False):
    """equinox.nn.MaxPool2d: reduce_window(max) with -inf padding. use_ceil=True pads the right/bottom edge (with
    -inf) by one extra stride when the window sweep does not divide evenly, i.e. partial windows at the edge; for
    every shape on which the last window still overlaps the input this equals torch's ceil_mode, which is what the
    reference's SqueezeNet / GoogLeNet tests pin at 1e-4."""
    return F.max_pool2d(x, _pair(k), _pair(stride), _pair(padding), ceil_mode=ceil_mode)


def avg_pool2d(x, k, stride):
    return F.avg_pool2d(x, _pair(k), _pair(stride))


def adaptive_avg_pool2d(x, target, uneven: bool = False):
    """equinox.nn.AdaptiveAvgPool2d. Even split (dim % target == 0): block mean. `uneven=True` opts into equinox's
    rule for the other case - the first `dim % target` blocks have dim // target + 1 elements, the rest dim // target -
    which is NOT torch's overlapping-window rule; the device library builds neither, so callers
    must ask for it explicitly (GoogLeNet's auxiliary heads, googlenet.py:265-268)."""
    oh, ow = _pair(target)
    n, c, h, w = x.shape
    if h % oh or w % ow:
        if not uneven:
            raise NotImplementedError("uneven adaptive pooling differs between equinox and torch")
        for axis, t in ((2, oh), (3, ow)):
            size = x.shape[axis]
            head, block = size % t, size // t
            parts = []
            if head:
                parts.append(x.narrow(axis, 0, head * (block + 1)).unflatten(axis, (head, block + 1)).mean(axis + 1))
            parts.append(x.narrow(axis, head * (block + 1), (t - head) * block).unflatten(axis, (t - head, block))
                         .mean(axis + 1))
            x = torch.cat(parts, axis)
        return x
    return x.reshape(n, c, oh, h // oh, ow, w // ow).mean((3, 5))


# jax.nn activations ---------------------------------------------------------------------------
def relu(x):
    return torch.clamp_min(x, 0.0)


def relu6(x):
    return torch.clamp(x, 0.0, 6.0)


def silu(x):
    return x * torch.sigmoid(x)


def gelu_tanh(x):
    """jax.nn.gelu(approximate=True), the default used by vit.py:96 / mlps.py:62"""
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))


def hard_sigmoid(x):
    return relu6(x + 3.0) / 6.0


def hard_swish(x):
    return x * hard_sigmoid(x)


def sigmoid(x):
    return torch.sigmoid(x)


ACTS = {None: lambda v: v, "relu": relu, "relu6": relu6, "silu": silu, "gelu": gelu_tanh,
        "hard_sigmoid": hard_sigmoid, "hard_swish": hard_swish, "sigmoid": sigmoid}


def softmax(x, dim=-1):
    m = x.max(dim, keepdim=True).values
    e = torch.exp(x - m)
    return e / e.sum(dim, keepdim=True)


def resize_bilinear(x, h, w):
    """jax.image.resize(method='bilinear') when upsampling == half-pixel centres, edge clamped
    (== F.interpolate(mode='bilinear', align_corners=False); asserted by test_deeplabv3.py:27)"""
    return F.interpolate(x, size=(h, w), mode="bilinear", align_corners=False)


# ------------------------------------------------------------------------------------------------
# fused-op helpers + optional bf16 emulation
# ------------------------------------------------------------------------------------------------
# The device path stores every activation in bf16 and rounds exactly once per fused kernel
# (conv + folded BN + activation (+ residual)), with the BatchNorm scale folded into the bf16
# filter.  With `emulate_bf16(True)` the helpers below round at the same points, so the oracle
# becomes a bit-faithful model of the device arithmetic up to fp32 summation order; the tests use it
# to separate implementation errors (tight tolerance vs the emulation) from the precision gap
# (stated tolerance vs the plain fp32 oracle).
_EMULATE = False
_ROUND_ACTIVATIONS = True


class emulate_bf16:
    """`activations=False` keeps the bf16 filters (BatchNorm folded before the rounding, as the device packs them)
    but leaves every activation in fp32: the mode tests/test_plan_lowering.py compares the fp32 replay of a launch
    plan against, where the only remaining difference is fp32 summation order."""

    def __init__(self, on: bool = True, activations: bool = True):
        self.on, self.activations = on, activations

    def __enter__(self):
        global _EMULATE, _ROUND_ACTIVATIONS
        self.prev, _EMULATE = (_EMULATE, _ROUND_ACTIVATIONS), self.on
        _ROUND_ACTIVATIONS = self.activations
        return self

    def __exit__(self, *exc):
        global _EMULATE, _ROUND_ACTIVATIONS
        _EMULATE, _ROUND_ACTIVATIONS = self.prev


def rnd(x):
    """round an ACTIVATION to bf16 (round-to-nearest-even) when emulating, identity otherwise"""
    return x.to(torch.bfloat16).to(torch.float32) if (_EMULATE and _ROUND_ACTIVATIONS) else x


def rnd_w(w):
    """round a packed FILTER to bf16 when emulating"""
    return w.to(torch.bfloat16).to(torch.float32) if _EMULATE else w


def conv_bn_act(x, weight, bias=None, bn=None, stride=1, padding=0, dilation=1, groups=1, act=None, res=None,
                res_after_act=False, eps=1e-5, round_out=True):
    """Conv2d -> [BatchNorm(inference)] -> act, with the residual added before (default) or after the
    activation: one fused device kernel.  bn = (weight, bias, mean, var)."""
    if _EMULATE:
        w = weight.double()
        b = None if bias is None else bias.double().reshape(-1)
        if bn is not None:
            scale = bn[0].double() / torch.sqrt(bn[3].double() + eps)
            shift = bn[1].double() - bn[2].double() * scale
            w = w * scale.reshape(-1, 1, 1, 1)
            b = shift if b is None else b * scale + shift
        # depthwise filters stay fp32 on the device ([K*K][C] fp32, _pack.pack_depthwise_weight): no rounding there
        depthwise = groups > 1 and groups == w.shape[0] and w.shape[1] == 1
        wq = w.float() if depthwise else rnd_w(w.float())
        y = conv2d(rnd(x), wq, None if b is None else b.float(), stride, padding, dilation, groups)
    else:
        y = conv2d(x, weight, bias, stride, padding, dilation, groups)
        if bn is not None:
            y = batch_norm_inference(y, bn[0], bn[1], bn[2], bn[3], eps)
    if res is not None and not res_after_act:
        y = y + res
    y = ACTS[act](y)
    if res is not None and res_after_act:
        y = y + res
    return rnd(y) if round_out else y


def linear_act(x, weight, bias=None, act=None, res=None, round_out=True):
    """Linear -> act (+ residual after): one fused device GEMM"""
    y = linear(rnd(x), rnd_w(weight), bias)
    y = ACTS[act](y)
    if res is not None:
        y = y + res
    return rnd(y) if round_out else y
