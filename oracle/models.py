"""ORACLE (test infrastructure) — CPU fp32 restatement of the reference's model forwards.

Each function follows the cited reference lines and consumes the checkpoint POSITIONALLY, the way
eqxvision/utils.py:172-218 does: non-"running"/"num_batches" tensors in state_dict order for the
array leaves in field order, (running_mean, running_var) pairs in state_dict order for the
BatchNorm state.  Inputs are batched (N,C,H,W) fp32 tensors (the reference vmaps a per-sample
function; every op used here is batch-independent in inference mode).
See oracle/ops.py for the parity status.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import ops as O


class Stream:
    """positional view of a torch state_dict (utils.py:172-187)"""

    def __init__(self, state_dict: Dict[str, torch.Tensor]):
        self.w = [v.detach().float() if v.is_floating_point() else v.detach()
                  for k, v in state_dict.items() if "running" not in k and "num_batches" not in k]
        means = [v.detach().float() for k, v in state_dict.items() if "running_mean" in k]
        var = [v.detach().float() for k, v in state_dict.items() if "running_var" in k]
        self.s = list(zip(means, var))
        self.wi = 0
        self.si = 0

    def take(self, shape=None):
        t = self.w[self.wi]
        self.wi += 1
        return t if shape is None else t.reshape(shape)

    def take_bn(self):
        """(weight, bias, mean, var) of the next BatchNorm in tree order"""
        w, b = self.take(), self.take()
        m, v = self.s[self.si]
        self.si += 1
        return (w, b, m, v)

    def done(self) -> bool:
        return self.wi == len(self.w) and self.si == len(self.s)


def _bn(x, p, eps=1e-5):
    return O.batch_norm_inference(x, p[0], p[1], p[2], p[3], eps)


# ------------------------------------------------------------------------------------------------
# ResNet (resnet.py)
# ------------------------------------------------------------------------------------------------
_RESNETS = {
    "resnet18": ("basic", [2, 2, 2, 2]), "resnet34": ("basic", [3, 4, 6, 3]),
    "resnet50": ("bottleneck", [3, 4, 6, 3]), "resnet101": ("bottleneck", [3, 4, 23, 3]),
    "resnet152": ("bottleneck", [3, 8, 36, 3]),
}


def _read_resnet_block(s: Stream, kind: str, has_ds: bool):
    # field order: conv1, bn1, conv2, bn2, [conv3, bn3], downsample (resnet.py:38-45, 101-110)
    n = 2 if kind == "basic" else 3
    convs = []
    for _ in range(n):
        w = s.take()
        convs.append((w, s.take_bn()))
    ds = None
    if has_ds:
        w = s.take()
        ds = (w, s.take_bn())
    return convs, ds


def _resnet_block(x, kind, convs, ds, stride, dilation):
    identity = x if ds is None else O.conv_bn_act(x, ds[0], None, ds[1], stride)      # downsample(x)
    if kind == "basic":  # resnet.py:80-92
        out = O.conv_bn_act(x, convs[0][0], None, convs[0][1], stride, 1, act="relu")
        return O.conv_bn_act(out, convs[1][0], None, convs[1][1], 1, 1, act="relu", res=identity)
    # resnet.py:144-162; stride and dilation on conv2, padding == dilation (resnet.py:15-27)
    out = O.conv_bn_act(x, convs[0][0], None, convs[0][1], act="relu")
    out = O.conv_bn_act(out, convs[1][0], None, convs[1][1], stride, dilation, dilation, act="relu")
    return O.conv_bn_act(out, convs[2][0], None, convs[2][1], act="relu", res=identity)  # out += identity; relu


def resnet_features(s: Stream, x, arch: str, replace_stride_with_dilation=(False, False, False)):
    """stem + 4 stages; returns the list of stage outputs (resnet.py:243-282, 344-352)"""
    kind, depths = _RESNETS[arch]
    expansion = 1 if kind == "basic" else 4
    w = s.take()
    x = O.conv_bn_act(x, w, None, s.take_bn(), 2, 3, act="relu")
    x = O.max_pool2d(x, 3, 2, 1)
    inplanes, dilation = 64, 1
    stages = []
    for i, (planes, depth) in enumerate(zip((64, 128, 256, 512), depths)):
        stride = 1 if i == 0 else 2
        prev_dil = dilation  # resnet.py:289-294: first block keeps the previous dilation
        if i > 0 and replace_stride_with_dilation[i - 1]:
            dilation *= stride
            stride = 1
        has_ds = stride != 1 or inplanes != planes * expansion
        convs, ds = _read_resnet_block(s, kind, has_ds)
        x = _resnet_block(x, kind, convs, ds, stride, prev_dil)
        inplanes = planes * expansion
        for _ in range(1, depth):
            convs, ds = _read_resnet_block(s, kind, False)
            x = _resnet_block(x, kind, convs, None, 1, dilation)
        stages.append(x)
    return stages


def resnet(state_dict, x, arch="resnet50"):
    """ResNet.__call__ (resnet.py:335-358) -> logits (N, num_classes)"""
    s = Stream(state_dict)
    feat = resnet_features(s, x, arch)[-1]
    pooled = O.rnd(O.adaptive_avg_pool2d(feat, 1).flatten(1))  # avgpool + ravel
    logits = O.linear_act(pooled, s.take(), s.take(), round_out=False)
    assert s.done(), "checkpoint has tensors the reference would silently ignore"
    return logits


# ------------------------------------------------------------------------------------------------
# Vision Transformer (vit.py, layers/patch_embed.py, layers/mlps.py)
# ------------------------------------------------------------------------------------------------
def vit_attention(x, qkv_w, qkv_b, proj_w, proj_b, heads, res=None):
    """_VitAttention.__call__ (vit.py:56-76) on (B, N, C); returns (proj(out) [+ res], attn)"""
    b, n, c = x.shape
    d = c // heads
    qkv = O.linear_act(x, qkv_w, qkv_b)                                # vit.py:64
    qkv = qkv.reshape(b, n, 3, heads, d).permute(2, 0, 3, 1, 4)        # vit.py:65-66
    q, k, v = qkv[0], qkv[1], qkv[2]
    logits = (q @ k.transpose(-1, -2)) * (d ** -0.5)                   # vit.py:69 (scale after product)
    attn = O.softmax(logits, -1)                                       # vit.py:70
    if O._EMULATE:
        # device kernel: unnormalised probabilities are rounded to bf16 for the P.V MMA, the row sum
        # is accumulated in fp32 from the unrounded values
        p = torch.exp(logits - logits.max(-1, keepdim=True).values)
        out = (O.rnd(p) @ v) / p.sum(-1, keepdim=True)
    else:
        out = attn @ v
    out = O.rnd(out.permute(0, 2, 1, 3).reshape(b, n, c))              # vit.py:73
    return O.linear_act(out, proj_w, proj_b, res=res), attn            # vit.py:74 (+ residual vit.py:153)


def vit(state_dict, x, heads=12, patch=16, eps=1e-5, return_last_attention=False):
    """VisionTransformer.__call__ (vit.py:261-273). DINO/timm key order: cls_token, pos_embed,
    patch_embed.proj.{w,b}, blocks.i.{norm1, attn.qkv, attn.proj, norm2, mlp.fc1, mlp.fc2}, norm, [head]."""
    s = Stream(state_dict)
    cls = s.take()
    cls = cls.reshape(1, cls.shape[-1])                                # (1,1,D) -> (1,D), utils.py:197
    pos = s.take()
    pos = pos.reshape(-1, pos.shape[-1])
    pw, pb = s.take(), s.take()
    b = x.shape[0]
    t = O.conv_bn_act(x, pw, pb, None, patch, 0)                       # patch_embed.py:79
    t = t.flatten(2).transpose(1, 2)                                   # ravel + moveaxis -> (B, np, D)
    t = O.rnd(torch.cat([cls.expand(b, 1, -1), t], 1) + pos)           # vit.py:269
    depth = (len(s.w) - s.wi - 2) // 12
    for i in range(depth):
        n1w, n1b = s.take(), s.take()
        qw, qb, ow, ob = s.take(), s.take(), s.take(), s.take()
        n2w, n2b = s.take(), s.take()
        f1w, f1b, f2w, f2b = s.take(), s.take(), s.take(), s.take()
        y = O.rnd(O.layer_norm(t, n1w, n1b, eps))                      # vit.py:149
        t_new, attn = vit_attention(y, qw, qb, ow, ob, heads, res=t)   # vit.py:150,153
        if return_last_attention and i == depth - 1:
            return attn.unsqueeze(1)                                   # per-sample (1,H,N,N)
        t = t_new
        y = O.rnd(O.layer_norm(t, n2w, n2b, eps))                      # vit.py:154
        y = O.linear_act(y, f1w, f1b, act="gelu")                      # mlps.py:61-62
        t = O.linear_act(y, f2w, f2b, res=t)                           # mlps.py:64, vit.py:156
    # vit.py:272 normalises all tokens, vit.py:273 keeps row 0: LayerNorm is row-wise
    out = O.rnd(O.layer_norm(t[:, 0], s.take(), s.take(), eps))
    if s.wi < len(s.w):                                                # fc when num_classes > 0
        out = O.linear_act(out, s.take(), s.take(), round_out=False)
    assert s.done()
    return out
