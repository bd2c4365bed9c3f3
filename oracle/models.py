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

import math

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
    # resnet.py:440-511: same block lists, groups / width_per_group only change the conv2 weights' shapes
    "resnext50_32x4d": ("bottleneck", [3, 4, 6, 3]), "resnext101_32x8d": ("bottleneck", [3, 4, 23, 3]),
    "wide_resnet50_2": ("bottleneck", [3, 4, 6, 3]), "wide_resnet101_2": ("bottleneck", [3, 4, 23, 3]),
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
    groups = out.shape[1] // convs[1][0].shape[1]     # ResNeXt: conv2 is grouped (resnet.py:83, _conv3x3 :19-23)
    out = O.conv_bn_act(out, convs[1][0], None, convs[1][1], stride, dilation, dilation, groups, act="relu")
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


# ------------------------------------------------------------------------------------------------
# shared pieces for the mobile families
# ------------------------------------------------------------------------------------------------
def _make_divisible(v, divisor, min_value=None):
    """utils.py:104-117"""
    if min_value is None:
        min_value = divisor
    new_v = max(min_value, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v


def squeeze_excitation(s: Stream, x, act, gate):
    """SqueezeExcitation.__call__ (layers/squeeze.py:51-61); fc1/fc2 are 1x1 convs WITH bias"""
    w1, b1, w2, b2 = s.take(), s.take(), s.take(), s.take()
    scale = O.rnd(O.adaptive_avg_pool2d(x, 1))
    scale = O.conv_bn_act(scale, w1, b1, None, act=act)
    scale = O.conv_bn_act(scale, w2, b2, None, act=gate)
    return O.rnd(x * scale)


def _cna(s: Stream, x, stride=1, padding=0, dilation=1, groups=1, act=None, eps=1e-5, res=None):
    """ConvNormActivation with BatchNorm and no conv bias (conv_norm_activation.py:56-85)"""
    w = s.take()
    return O.conv_bn_act(x, w, None, s.take_bn(), stride, padding, dilation, groups, act=act, res=res, eps=eps)


# ------------------------------------------------------------------------------------------------
# EfficientNet (efficientnet.py)
# ------------------------------------------------------------------------------------------------
_EFFNET_V1 = [(1, 3, 1, 32, 16, 1), (6, 3, 2, 16, 24, 2), (6, 5, 2, 24, 40, 2), (6, 3, 2, 40, 80, 3),
              (6, 5, 1, 80, 112, 3), (6, 5, 2, 112, 192, 4), (6, 3, 1, 192, 320, 1)]  # efficientnet.py:434-442
_EFFNET_MULT = {"efficientnet_b0": (1.0, 1.0, 1e-5), "efficientnet_b1": (1.0, 1.1, 1e-5),
                "efficientnet_b2": (1.1, 1.2, 1e-5), "efficientnet_b3": (1.2, 1.4, 1e-5),
                "efficientnet_b4": (1.4, 1.8, 1e-5), "efficientnet_b5": (1.6, 2.2, 1e-3),
                "efficientnet_b6": (1.8, 2.6, 1e-3), "efficientnet_b7": (2.0, 3.1, 1e-3)}
_EFFNET_V2 = {"efficientnet_v2_s": [("F", 1, 3, 1, 24, 24, 2), ("F", 4, 3, 2, 24, 48, 4), ("F", 4, 3, 2, 48, 64, 4),
                                    ("M", 4, 3, 2, 64, 128, 6), ("M", 6, 3, 1, 128, 160, 9),
                                    ("M", 6, 3, 2, 160, 256, 15)]}  # efficientnet.py:445-453


def efficientnet(state_dict, x, arch="efficientnet_b4"):
    """EfficientNet.__call__ (efficientnet.py:392-403) with _MBConv (101-186) / _FusedMBConv (195-266)"""
    import math

    s = Stream(state_dict)
    if arch in _EFFNET_MULT:
        wm, dm, eps = _EFFNET_MULT[arch]
        rows = [("M", e, k, st, _make_divisible(ci * wm, 8), _make_divisible(co * wm, 8), int(math.ceil(n * dm)))
                for (e, k, st, ci, co, n) in _EFFNET_V1]
    else:
        rows, eps = _EFFNET_V2[arch], 1e-3
    x = _cna(s, x, 2, 1, act="silu", eps=eps)                                  # stem 3x3 s2
    for kind, e, k, st, ci, co, n in rows:
        for i in range(n):
            cin, stride = (ci, st) if i == 0 else (co, 1)
            exp = _make_divisible(cin * e, 8)
            use_res = stride == 1 and cin == co
            inp = x
            if kind == "M":
                if exp != cin:                                                  # expand (efficientnet.py:127)
                    x = _cna(s, x, act="silu", eps=eps)
                x = _cna(s, x, stride, (k - 1) // 2, groups=exp, act="silu", eps=eps)   # depthwise
                x = squeeze_excitation(s, x, "silu", "sigmoid")                  # squeeze = max(1, cin//4)
                x = _cna(s, x, eps=eps, res=inp if use_res else None)            # project (+ x)
            else:
                if exp != cin:
                    x = _cna(s, x, stride, (k - 1) // 2, act="silu", eps=eps)
                    x = _cna(s, x, eps=eps, res=inp if use_res else None)
                else:                                                            # single conv, act then + x
                    w = s.take()
                    x = O.conv_bn_act(x, w, None, s.take_bn(), stride, (k - 1) // 2, act="silu", eps=eps,
                                      res=inp if use_res else None, res_after_act=True)
    x = _cna(s, x, act="silu", eps=eps)                                         # head 1x1
    x = O.rnd(O.adaptive_avg_pool2d(x, 1).flatten(1))
    out = O.linear_act(x, s.take(), s.take(), round_out=False)                  # Dropout is a no-op
    assert s.done()
    return out


# ------------------------------------------------------------------------------------------------
# MobileNetV3 (mobilenetv3.py)
# ------------------------------------------------------------------------------------------------
_MBV3 = {
    "mobilenet_v3_large": ([(16, 3, 16, 16, False, "RE", 1), (16, 3, 64, 24, False, "RE", 2),
                            (24, 3, 72, 24, False, "RE", 1), (24, 5, 72, 40, True, "RE", 2),
                            (40, 5, 120, 40, True, "RE", 1), (40, 5, 120, 40, True, "RE", 1),
                            (40, 3, 240, 80, False, "HS", 2), (80, 3, 200, 80, False, "HS", 1),
                            (80, 3, 184, 80, False, "HS", 1), (80, 3, 184, 80, False, "HS", 1),
                            (80, 3, 480, 112, True, "HS", 1), (112, 3, 672, 112, True, "HS", 1),
                            (112, 5, 672, 160, True, "HS", 2), (160, 5, 960, 160, True, "HS", 1),
                            (160, 5, 960, 160, True, "HS", 1)], 1280),
    "mobilenet_v3_small": ([(16, 3, 16, 16, True, "RE", 2), (16, 3, 72, 24, False, "RE", 2),
                            (24, 3, 88, 24, False, "RE", 1), (24, 5, 96, 40, True, "HS", 2),
                            (40, 5, 240, 40, True, "HS", 1), (40, 5, 240, 40, True, "HS", 1),
                            (40, 5, 120, 48, True, "HS", 1), (48, 5, 144, 48, True, "HS", 1),
                            (48, 5, 288, 96, True, "HS", 2), (96, 5, 576, 96, True, "HS", 1),
                            (96, 5, 576, 96, True, "HS", 1)], 1024),
}  # mobilenetv3.py:265-336 (width_mult 1, not dilated, not reduced)


def mobilenet_v3_features(s: Stream, x, arch="mobilenet_v3_small", dilated=False, taps=()):
    """MobileNetV3.features (mobilenetv3.py:190-222): stem, _InvertedResidual stack (62-132), last 1x1 conv;
    BN eps 1e-3 (:189). `dilated` (segmentation backbone, mobilenetv3.py:265-336 with dilated=True): the last three
    blocks use dilation 2 and the stride-2 one of them stride 1 (:88). Returns (out, [outputs of features[i] for i in taps])."""
    rows, _ = _MBV3[arch]
    eps = 1e-3
    feats = []
    x = _cna(s, x, 2, 1, act="hard_swish", eps=eps)
    feats.append(x)
    for i, (cin, k, exp, cout, use_se, a, stride) in enumerate(rows):
        act = "hard_swish" if a == "HS" else "relu"
        dil = 2 if (dilated and i >= len(rows) - 3) else 1
        st = 1 if dil > 1 else stride                                           # mobilenetv3.py:88
        inp = x
        if exp != cin:
            x = _cna(s, x, act=act, eps=eps)
        x = _cna(s, x, st, (k - 1) // 2 * dil, dil, groups=exp, act=act, eps=eps)
        if use_se:
            x = squeeze_excitation(s, x, "relu", "hard_sigmoid")                # mobilenetv3.py:56-58,103
        x = _cna(s, x, eps=eps, res=inp if (stride == 1 and cin == cout) else None)
        feats.append(x)
    x = _cna(s, x, act="hard_swish", eps=eps)                                   # 6 * last channels
    feats.append(x)
    return x, [feats[t] for t in taps]


def lraspp_mobilenet_v3_large(state_dict, x):
    """LRASPP.__call__ (lraspp.py:56-68) with LRASPPHead (lraspp.py:71-116) on the dilated MobileNetV3-Large
    backbone, taps features[4] (low, C2) and features[16] (high) (lraspp.py:160-166). Returns `out` (N, classes, H, W);
    the reference returns (None, out)."""
    s = Stream(state_dict)
    h, w = x.shape[-2:]
    _, (low, high) = mobilenet_v3_features(s, x, "mobilenet_v3_large", dilated=True, taps=(4, 16))
    y = _cna(s, high, act="relu")                                               # cbr: conv1x1 -> BN -> relu
    ws = s.take()                                                               # scale: pool -> conv1x1 -> sigmoid
    g = O.conv_bn_act(O.rnd(O.adaptive_avg_pool2d(high, 1)), ws, None, None, act="sigmoid")
    y = O.rnd(y * g)                                                            # lraspp.py:113
    y = O.rnd(O.resize_bilinear(y, low.shape[-2], low.shape[-1]))               # lraspp.py:114
    wl, bl, wh, bh = s.take(), s.take(), s.take(), s.take()
    lo = O.conv_bn_act(low, wl, bl, None)                                       # low_classifier
    out = O.conv_bn_act(y, wh, bh, None, res=lo)                                # + high_classifier (lraspp.py:116)
    assert s.done()
    return O.resize_bilinear(out, h, w)                                         # lraspp.py:67


def mobilenet_v3(state_dict, x, arch="mobilenet_v3_small"):
    """MobileNetV3.__call__ (mobilenetv3.py:236-247) with _InvertedResidual (62-132); BN eps 1e-3 (:189)"""
    s = Stream(state_dict)
    x, _ = mobilenet_v3_features(s, x, arch)
    x = O.rnd(O.adaptive_avg_pool2d(x, 1).flatten(1))
    x = O.linear_act(x, s.take(), s.take(), act="hard_swish")                   # Linear -> hswish -> Dropout
    out = O.linear_act(x, s.take(), s.take(), round_out=False)
    assert s.done()
    return out


# ------------------------------------------------------------------------------------------------
# MobileNetV2 (mobilenetv2.py)
# ------------------------------------------------------------------------------------------------
_MBV2 = [(1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2),
         (6, 320, 1, 1)]  # t, c, n, s: mobilenetv2.py:140-149


def mobilenet_v2(state_dict, x, arch="mobilenet_v2"):
    """MobileNetV2.__call__ (mobilenetv2.py:217-227) over _InvertedResidual (16-88): [expand 1x1 CNA if t != 1],
    depthwise 3x3 CNA, 1x1 conv (no bias) -> BN, `x + conv(x)` when stride 1 and inp == oup (:42,:85-88).
    REFERENCE QUIRK: activations are jnn.relu (:54,:67,:176,:200), not torchvision's ReLU6."""
    s = Stream(state_dict)
    x = _cna(s, x, 2, 1, act="relu")                                            # mobilenetv2.py:170-178
    cin = 32
    for t, c, n, st in _MBV2:
        for i in range(n):
            stride = st if i == 0 else 1
            inp = x
            hidden = int(round(cin * t))
            if t != 1:
                x = _cna(s, x, act="relu")
            x = _cna(s, x, stride, 1, groups=hidden, act="relu")
            x = _cna(s, x, res=inp if (stride == 1 and cin == c) else None)
            cin = c
    x = _cna(s, x, act="relu")                                                  # 1x1 -> 1280
    x = O.rnd(O.adaptive_avg_pool2d(x, 1).flatten(1))
    out = O.linear_act(x, s.take(), s.take(), round_out=False)                  # Dropout (no-op) -> Linear
    assert s.done()
    return out


# ------------------------------------------------------------------------------------------------
# RegNet (regnet.py)
# ------------------------------------------------------------------------------------------------
def regnet(state_dict, x, arch="regnet_y_400mf", stages=None, group_width=None, se_ratio=None, stem_width=32):
    """RegNet.__call__ (regnet.py:420-430). `stages` = [(width, depth)], per-stage stride 2 (regnet.py:254);
    ResBottleneckBlock (regnet.py:113-167): relu(proj(x) + f(x)), f = 1x1 CNA, grouped 3x3 CNA (stride here),
    [SE with squeeze width round(se_ratio * width_in), ReLU inside, sigmoid gate], 1x1 CNA without activation.
    Checkpoint order per block = field order (proj, f): the projection's conv+BN first when present."""
    if stages is None:
        stages, group_width, se_ratio = _REGNET[arch]
    s = Stream(state_dict)
    x = _cna(s, x, 2, 1, act="relu")                                            # SimpleStemIN
    cin = stem_width
    for width, depth in stages:
        for i in range(depth):
            stride = 2 if i == 0 else 1
            w_in = cin if i == 0 else width
            proj = _cna(s, x, stride) if (w_in != width or stride != 1) else x  # 1x1, stride on the conv
            gw = min(group_width, width)
            y = _cna(s, x, act="relu")
            y = _cna(s, y, stride, 1, groups=width // gw, act="relu")
            if se_ratio:
                y = squeeze_excitation(s, y, "relu", "sigmoid")
            x = _cna(s, y, act="relu", res=proj)                                # relu(proj + f): regnet.py:166-167
        cin = width
    x = O.rnd(O.adaptive_avg_pool2d(x, 1).flatten(1))
    out = O.linear_act(x, s.take(), s.take(), round_out=False)
    assert s.done()
    return out


# (stage (width, depth) list, group width, se_ratio) as BlockParams.from_init_params yields them (regnet.py:222-297)
_REGNET = {
    "regnet_y_400mf": ([(48, 1), (104, 3), (208, 6), (440, 6)], 8, 0.25),
    "regnet_x_400mf": ([(32, 1), (64, 2), (160, 7), (400, 12)], 16, None),
    "regnet_y_800mf": ([(64, 1), (144, 3), (320, 8), (784, 2)], 16, 0.25),
}


# ------------------------------------------------------------------------------------------------
# SqueezeNet (squeezenet.py)
# ------------------------------------------------------------------------------------------------
_SQUEEZE = {  # "P" = max-pool 3x3/2 use_ceil=True, tuples = _Fire(in, squeeze, expand1x1, expand3x3): squeezenet.py:83-118
    "squeezenet1_0": ((7, 96), ["P", (96, 16, 64, 64), (128, 16, 64, 64), (128, 32, 128, 128), "P",
                                (256, 32, 128, 128), (256, 48, 192, 192), (384, 48, 192, 192), (384, 64, 256, 256), "P",
                                (512, 64, 256, 256)]),
    "squeezenet1_1": ((3, 64), ["P", (64, 16, 64, 64), (128, 16, 64, 64), "P", (128, 32, 128, 128),
                                (256, 32, 128, 128), "P", (256, 48, 192, 192), (384, 48, 192, 192),
                                (384, 64, 256, 256), (512, 64, 256, 256)]),
}


def squeezenet(state_dict, x, arch="squeezenet1_0"):
    """SqueezeNet.__call__ (squeezenet.py:137-142): features -> [Dropout, conv1x1, ReLU, global avgpool] -> ravel.
    _Fire.__call__ (squeezenet.py:47-55): relu(squeeze(x)) -> concat(relu(expand1x1), relu(expand3x3)) on channels."""
    (k, _), layers = _SQUEEZE[arch]
    s = Stream(state_dict)
    x = O.conv_bn_act(x, s.take(), s.take(), None, 2, 0, act="relu")
    for item in layers:
        if item == "P":
            x = O.max_pool2d(x, 3, 2, ceil_mode=True)
            continue
        y = O.conv_bn_act(x, s.take(), s.take(), None, act="relu")
        e1 = O.conv_bn_act(y, s.take(), s.take(), None, act="relu")
        e3 = O.conv_bn_act(y, s.take(), s.take(), None, 1, 1, act="relu")
        x = torch.cat([e1, e3], 1)
    x = O.conv_bn_act(x, s.take(), s.take(), None, act="relu")                  # classifier conv (with bias)
    out = O.rnd(O.adaptive_avg_pool2d(x, 1)).flatten(1)
    assert s.done()
    return out


# ------------------------------------------------------------------------------------------------
# GoogLeNet (googlenet.py)
# ------------------------------------------------------------------------------------------------
_INCEPTION = {  # in, 1x1, 3x3red, 3x3, 5x5red, 5x5, pool_proj: googlenet.py:97-112
    "3a": (192, 64, 96, 128, 16, 32, 32), "3b": (256, 128, 128, 192, 32, 96, 64),
    "4a": (480, 192, 96, 208, 16, 48, 64), "4b": (512, 160, 112, 224, 24, 64, 64),
    "4c": (512, 128, 128, 256, 24, 64, 64), "4d": (512, 112, 144, 288, 32, 64, 64),
    "4e": (528, 256, 160, 320, 32, 128, 128), "5a": (832, 256, 160, 320, 32, 128, 128),
    "5b": (832, 384, 192, 384, 48, 128, 128)}


def _inception(s: Stream, x):
    """_Inception.__call__ (googlenet.py:232-240); BasicConv2d = conv(no bias) -> BN(eps 1e-3) -> relu (:287-311)"""
    e = 1e-3
    b1 = _cna(s, x, act="relu", eps=e)
    b2 = _cna(s, _cna(s, x, act="relu", eps=e), 1, 1, act="relu", eps=e)
    b3 = _cna(s, _cna(s, x, act="relu", eps=e), 1, 1, act="relu", eps=e)        # 3x3, not 5x5 (googlenet.py:214-216)
    b4 = _cna(s, O.max_pool2d(x, 3, 1, 1, ceil_mode=True), act="relu", eps=e)
    return torch.cat([b1, b2, b3, b4], 1)


def _inception_aux(x, p):
    """InceptionAux.__call__ (googlenet.py:263-279): adaptive pool to 4x4 (equinox's uneven rule on 14x14 maps),
    BasicConv2d 1x1 -> ravel (C,H,W order) -> fc1 -> relu -> Dropout (no-op) -> fc2"""
    wc, bn, w1, b1, w2, b2 = p
    y = O.rnd(O.adaptive_avg_pool2d(x, 4, uneven=True))
    y = O.conv_bn_act(y, wc, None, bn, act="relu", eps=1e-3).flatten(1)
    y = O.linear_act(y, w1, b1, act="relu")
    return O.linear_act(y, w2, b2, round_out=False)


def googlenet(state_dict, x, arch="googlenet", aux_logits=False):
    """GoogLeNet.__call__ (googlenet.py:108-177). A checkpoint that carries the auxiliary heads (torchvision saves
    them; the reference loads them positionally, googlenet.py:322-327) has them skipped when `aux_logits` is False:
    they sit between inception5b and fc in field order. With `aux_logits=True` returns (logits, aux2, aux1)
    (googlenet.py:174-175) - CPU oracle only, the device library does not build the uneven 14x14 -> 4x4 pooling."""
    e = 1e-3
    s = Stream(state_dict)
    if aux_logits:
        return _googlenet_with_aux(s, x)
    x = _cna(s, x, 2, 3, act="relu", eps=e)
    x = O.max_pool2d(x, 3, 2, ceil_mode=True)
    x = _cna(s, x, act="relu", eps=e)
    x = _cna(s, x, 1, 1, act="relu", eps=e)
    x = O.max_pool2d(x, 3, 2, ceil_mode=True)
    x = _inception(s, _inception(s, x))
    x = O.max_pool2d(x, 3, 2, ceil_mode=True)
    for _ in range(5):
        x = _inception(s, x)
    x = O.max_pool2d(x, 2, 2, ceil_mode=True)
    x = _inception(s, _inception(s, x))
    if any(k.startswith("aux1.") for k in state_dict):                          # 2 x (conv, bn.w, bn.b, fc1 w/b, fc2 w/b)
        for _ in range(2):
            s.take(), s.take_bn(), s.take(), s.take(), s.take(), s.take()
    x = O.rnd(O.adaptive_avg_pool2d(x, 1).flatten(1))
    out = O.linear_act(x, s.take(), s.take(), round_out=False)
    assert s.done()
    return out


# ------------------------------------------------------------------------------------------------
# ConvNeXt (convnext.py)
# ------------------------------------------------------------------------------------------------
_CONVNEXT = {"convnext_tiny": [(96, 3), (192, 3), (384, 9), (768, 3)], "convnext_small": [(96, 3), (192, 3), (384, 27), (768, 3)],
             "convnext_base": [(128, 3), (256, 3), (512, 27), (1024, 3)], "convnext_large": [(192, 3), (384, 3), (768, 27), (1536, 3)]}


def _ln2d(x, w, b, eps):
    """LayerNorm2d (layers/extensions_2d.py:24-28): per-pixel LayerNorm over channels"""
    return O.rnd(O.layer_norm(x.permute(0, 2, 3, 1), w, b, eps).permute(0, 3, 1, 2))


def convnext(state_dict, x, arch="convnext_tiny", block_eps=1e-5, eps=1e-6):
    """ConvNeXt.__call__ (convnext.py:198-204) over CNBlock (convnext.py:16-66):
    x + layer_scale * Linear2d(gelu_tanh(Linear2d(LayerNorm2d(dwconv7x7(x))))).
    REFERENCE QUIRKS: jnn.gelu = tanh approximation (convnext.py:52); the block norm is `LayerNorm2d(dim)` with the
    equinox default eps 1e-5 (convnext.py:24,39) while the other norms get 1e-6 (convnext.py:120).
    Checkpoint order per block: layer_scale, dw w/b, norm w/b, fc1 w/b, fc2 w/b (field order convnext.py:17-19)."""
    s = Stream(state_dict)
    x = O.conv_bn_act(x, s.take(), s.take(), None, 4, 0)                        # stem conv 4x4/4 with bias
    x = _ln2d(x, s.take(), s.take(), eps)
    stages = _CONVNEXT[arch]
    for si, (dim, depth) in enumerate(stages):
        for _ in range(depth):
            gamma = s.take().reshape(-1)
            wd, bd = s.take(), s.take()
            y = O.conv_bn_act(x, wd, bd, None, 1, 3, groups=dim)
            y = _ln2d(y, s.take(), s.take(), block_eps)
            t = y.permute(0, 2, 3, 1)
            t = O.linear_act(t, s.take(), s.take(), act="gelu")
            w2, b2 = s.take(), s.take()
            # the device folds layer_scale into fc2's rows (fp32 product, one bf16 rounding of the weight)
            t = O.linear_act(t, w2 * gamma[:, None], b2 * gamma, res=x.permute(0, 2, 3, 1))
            x = t.permute(0, 3, 1, 2)
        if si + 1 < len(stages):
            x = _ln2d(x, s.take(), s.take(), eps)
            x = O.conv_bn_act(x, s.take(), s.take(), None, 2, 0)                # downsample conv 2x2/2 with bias
    x = O.rnd(O.adaptive_avg_pool2d(x, 1))
    x = _ln2d(x, s.take(), s.take(), eps).flatten(1)
    out = O.linear_act(x, s.take(), s.take(), round_out=False)
    assert s.done()
    return out


# ------------------------------------------------------------------------------------------------
# ShuffleNetV2 (shufflenetv2.py)
# ------------------------------------------------------------------------------------------------
_SHUFFLENET = {"shufflenet_v2_x0_5": [24, 48, 96, 192, 1024], "shufflenet_v2_x1_0": [24, 116, 232, 464, 1024],
               "shufflenet_v2_x1_5": [24, 176, 352, 704, 1024], "shufflenet_v2_x2_0": [24, 244, 488, 976, 2048]}


def channel_shuffle(x, groups):
    """_channel_shuffle (shufflenetv2.py:14-20)"""
    n, c, h, w = x.shape
    return x.reshape(n, groups, c // groups, h, w).transpose(1, 2).reshape(n, c, h, w)


def shufflenet_v2(state_dict, x, arch="shufflenet_v2_x1_0"):
    """ShuffleNetV2.__call__ (shufflenetv2.py:252-262) over _InvertedResidual (shufflenetv2.py:132-140).
    The split / concat / shuffle only move bf16 values around, so they add no rounding in the emulating mode."""
    ch = _SHUFFLENET[arch]
    s = Stream(state_dict)
    x = _cna(s, x, 2, 1, act="relu")
    x = O.max_pool2d(x, 3, 2, 1)
    for repeats, cout in zip([4, 8, 4], ch[1:4]):
        for i in range(repeats):
            if i == 0:                                                          # stride 2: both branches see x
                b1 = _cna(s, x, 2, 1, groups=x.shape[1])
                b1 = _cna(s, b1, act="relu")
                b2 = _cna(s, x, act="relu")
                b2 = _cna(s, b2, 2, 1, groups=cout // 2)
                b2 = _cna(s, b2, act="relu")
                x = torch.cat([b1, b2], 1)
            else:                                                               # stride 1: x1 passes through
                x1, x2 = x.chunk(2, 1)
                b2 = _cna(s, x2, act="relu")
                b2 = _cna(s, b2, 1, 1, groups=cout // 2)
                b2 = _cna(s, b2, act="relu")
                x = torch.cat([x1, b2], 1)
            x = channel_shuffle(x, 2)
    x = _cna(s, x, act="relu")
    x = O.rnd(O.adaptive_avg_pool2d(x, 1).flatten(1))
    out = O.linear_act(x, s.take(), s.take(), round_out=False)
    assert s.done()
    return out


def _googlenet_with_aux(s: Stream, x):
    e = 1e-3
    x = _cna(s, x, 2, 3, act="relu", eps=e)
    x = O.max_pool2d(x, 3, 2, ceil_mode=True)
    x = _cna(s, x, act="relu", eps=e)
    x = _cna(s, x, 1, 1, act="relu", eps=e)
    x = O.max_pool2d(x, 3, 2, ceil_mode=True)
    x = _inception(s, _inception(s, x))
    x = O.max_pool2d(x, 3, 2, ceil_mode=True)
    x4a = _inception(s, x)
    x = _inception(s, _inception(s, _inception(s, x4a)))                        # 4b, 4c, 4d
    x4d = x
    x = _inception(s, x)
    x = O.max_pool2d(x, 2, 2, ceil_mode=True)
    x = _inception(s, _inception(s, x))
    heads = [(s.take(), s.take_bn(), s.take(), s.take(), s.take(), s.take()) for _ in range(2)]
    aux1, aux2 = _inception_aux(x4a, heads[0]), _inception_aux(x4d, heads[1])
    x = O.rnd(O.adaptive_avg_pool2d(x, 1).flatten(1))
    out = O.linear_act(x, s.take(), s.take(), round_out=False)
    assert s.done()
    return out, aux2, aux1


# ------------------------------------------------------------------------------------------------
# VGG (vgg.py)
# ------------------------------------------------------------------------------------------------
_VGG = {"A": [64, "M", 128, "M", 256, 256, "M", 512, 512, "M", 512, 512, "M"],
        "B": [64, 64, "M", 128, 128, "M", 256, 256, "M", 512, 512, "M", 512, 512, "M"],
        "D": [64, 64, "M", 128, 128, "M", 256, 256, 256, "M", 512, 512, 512, "M", 512, 512, 512, "M"],
        "E": [64, 64, "M", 128, 128, "M", 256, 256, 256, 256, "M", 512, 512, 512, 512, "M", 512, 512, 512, 512,
              "M"]}
_VGG_ARCH = {"vgg11": "A", "vgg13": "B", "vgg16": "D", "vgg19": "E"}


def vgg_features(s: Stream, x, cfg, batch_norm):
    """_make_layers (vgg.py:122-150): conv3x3 pad 1 WITH bias, [BN], ReLU; 'M' = maxpool 2x2/2"""
    for v in cfg:
        if v == "M":
            x = O.max_pool2d(x, 2, 2)
        else:
            w, b = s.take(), s.take()
            x = O.conv_bn_act(x, w, b, s.take_bn() if batch_norm else None, 1, 1, act="relu")
    return x


def vgg(state_dict, x, arch="vgg11", features_only=False):
    """VGG.__call__ (vgg.py:108-119). Classifier = Linear, Dropout, Linear, ReLU, Dropout, Linear: NO ReLU
    after the first Linear (vgg.py:97-106) - the reference's deviation from torchvision."""
    bn = arch.endswith("_bn")
    s = Stream(state_dict)
    x = vgg_features(s, x, _VGG[_VGG_ARCH[arch.replace("_bn", "")]], bn)
    if features_only:
        return x
    x = O.rnd(O.adaptive_avg_pool2d(x, 7)).flatten(1)                            # ravel in C,H,W order
    x = O.linear_act(x, s.take(), s.take())
    x = O.linear_act(x, s.take(), s.take(), act="relu")
    out = O.linear_act(x, s.take(), s.take(), round_out=False)
    assert s.done()
    return out


# ------------------------------------------------------------------------------------------------
# AlexNet (alexnet.py) - BASELINE config 0 (README example)
# ------------------------------------------------------------------------------------------------
def alexnet(state_dict, x, arch="alexnet", features_only=False):
    """AlexNet.__call__ (alexnet.py:72-85). features (alexnet.py:42-58): conv11x11/4 p2, conv5x5 p2, 3 x conv3x3 p1,
    all WITH bias and followed by ReLU, max-pool 3x3/2 (no padding) after convs 1, 2 and 5; adaptive average pool to
    6x6 (alexnet.py:59); classifier (alexnet.py:60-70): Dropout, Linear, ReLU, Dropout, Linear, ReLU, Linear -
    identical to torchvision's, which is what the reference's own test pins at 1e-4 (tests/test_models/test_alexnet.py)."""
    s = Stream(state_dict)
    for stride, pad, pool in ((4, 2, True), (1, 2, True), (1, 1, False), (1, 1, False), (1, 1, True)):
        w, b = s.take(), s.take()
        x = O.conv_bn_act(x, w, b, None, stride, pad, act="relu")
        if pool:
            x = O.max_pool2d(x, 3, 2)
    if features_only:
        return x
    x = O.rnd(O.adaptive_avg_pool2d(x, 6)).flatten(1)                            # jnp.ravel: C,H,W order
    x = O.linear_act(x, s.take(), s.take(), act="relu")
    x = O.linear_act(x, s.take(), s.take(), act="relu")
    out = O.linear_act(x, s.take(), s.take(), round_out=False)
    assert s.done()
    return out


# ------------------------------------------------------------------------------------------------
# DenseNet (densenet.py)
# ------------------------------------------------------------------------------------------------
_DENSENET = {"densenet121": (32, (6, 12, 24, 16), 64), "densenet161": (48, (6, 12, 36, 24), 96),
             "densenet169": (32, (6, 12, 32, 32), 64), "densenet201": (32, (6, 12, 48, 32), 64)}


def _bn_relu(x, p):
    return O.rnd(O.relu(_bn(x, p)))


def densenet(state_dict, x, arch="densenet121"):
    """DenseNet.__call__ (densenet.py:220-229): pre-activation dense layers (55-67), transitions (106-133)"""
    growth, blocks, init = _DENSENET[arch]
    s = Stream(state_dict)
    w = s.take()
    x = O.conv_bn_act(x, w, None, s.take_bn(), 2, 3, act="relu")
    x = O.max_pool2d(x, 3, 2, 1)
    for bi, depth in enumerate(blocks):
        feats = [x]
        for _ in range(depth):
            n1 = s.take_bn()
            w1 = s.take()
            n2 = s.take_bn()
            w2 = s.take()
            cat = torch.cat(feats, 1)                                            # densenet.py:63
            y = O.conv_bn_act(_bn_relu(cat, n1), w1, None, n2, act="relu")       # norm1,relu,conv1,norm2,relu
            feats.append(O.conv_bn_act(y, w2, None, None, 1, 1))                 # conv2 (growth channels)
        x = torch.cat(feats, 1)
        if bi != len(blocks) - 1:                                                # transition
            n = s.take_bn()
            w = s.take()
            x = O.rnd(O.avg_pool2d(O.conv_bn_act(_bn_relu(x, n), w), 2, 2))
    x = _bn_relu(x, s.take_bn())                                                 # norm5 + relu
    x = O.rnd(O.adaptive_avg_pool2d(x, 1).flatten(1))
    out = O.linear_act(x, s.take(), s.take(), round_out=False)
    assert s.done()
    return out


# ------------------------------------------------------------------------------------------------
# DeepLabV3-ResNet50 (segmentation/deeplabv3.py, fcn.py, _utils.py)
# ------------------------------------------------------------------------------------------------
def deeplabv3_resnet50(state_dict, x, rates=(12, 24, 36)):
    """_SimpleSegmentationModel.__call__ (_utils.py:36-60) -> (aux, out), both (N, classes, H, W).
    Checkpoint order: backbone (no fc), classifier.0.convs.{0..4}, classifier.0.project, classifier.1/2/4,
    aux_classifier.0/1/4."""
    s = Stream(state_dict)
    h, w = x.shape[-2:]
    stages = resnet_features(s, x, "resnet50", (False, True, True))              # deeplabv3.py:191-194
    c3, c4 = stages[2], stages[3]                                               # taps layer3 / layer4
    # ASPP (deeplabv3.py:77-135)
    branches = [_cna(s, c4, act="relu")]
    for r in rates:
        branches.append(_cna(s, c4, 1, r, r, act="relu"))                       # padding == dilation
    pooled = O.rnd(O.adaptive_avg_pool2d(c4, 1))
    pooled = _cna(s, pooled, act="relu")
    branches.append(O.rnd(O.resize_bilinear(pooled, c4.shape[-2], c4.shape[-1])))   # deeplabv3.py:74
    y = _cna(s, torch.cat(branches, 1), act="relu")                             # project (+Dropout no-op)
    y = _cna(s, y, 1, 1, act="relu")                                            # head 3x3
    wcls, bcls = s.take(), s.take()
    y = O.conv_bn_act(y, wcls, bcls, None)                                      # 1x1 with bias
    out = O.resize_bilinear(y, h, w)                                            # _utils.py:51-52
    # aux FCNHead on layer3 (fcn.py:19-34)
    a = _cna(s, c3, 1, 1, act="relu")
    wa, ba = s.take(), s.take()
    a = O.conv_bn_act(a, wa, ba, None)
    aux = O.resize_bilinear(a, h, w)
    assert s.done()
    return aux, out


def fcn_resnet50(state_dict, x):
    """FCN (fcn.py:37-120) over the dilated ResNet-50 backbone, taps layer3 (aux) / layer4: FCNHead (fcn.py:19-34) =
    conv3x3 (no bias) -> BN -> ReLU -> Dropout(0.1, no-op) -> conv1x1 (bias); both outputs resized bilinearly to the
    input size; returns (aux, out) as _SimpleSegmentationModel.__call__ does (_utils.py:58)."""
    s = Stream(state_dict)
    h, w = x.shape[-2:]
    stages = resnet_features(s, x, "resnet50", (False, True, True))
    c3, c4 = stages[2], stages[3]
    y = _cna(s, c4, 1, 1, act="relu")
    wc, bc = s.take(), s.take()
    out = O.resize_bilinear(O.conv_bn_act(y, wc, bc, None), h, w)
    a = _cna(s, c3, 1, 1, act="relu")
    wa, ba = s.take(), s.take()
    aux = O.resize_bilinear(O.conv_bn_act(a, wa, ba, None), h, w)
    assert s.done()
    return aux, out


# ------------------------------------------------------------------------------------------------
# Swin Transformer v1 (swin.py)
# ------------------------------------------------------------------------------------------------
_SWINS = {
    "swin_t": (96, [2, 2, 6, 2], [3, 6, 12, 24], 7),
    "swin_s": (96, [2, 2, 18, 2], [3, 6, 12, 24], 7),
    "swin_b": (128, [2, 2, 18, 2], [4, 8, 16, 32], 7),
}


def swin_shift_mask(h, w, ws, shift):
    """the -100 mask of swin.py:184-229, built the way the reference builds it: 3x3 region labels over
    the ROLLED map, window partition, pairwise label difference. Returns (num_windows, ws*ws, ws*ws)."""
    labels = torch.zeros(h, w)
    bands_h = ((0, h - ws), (h - ws, h - shift[0]), (h - shift[0], h))
    bands_w = ((0, w - ws), (w - ws, w - shift[1]), (w - shift[1], w))
    for i, (h0, h1) in enumerate(bands_h):
        for j, (w0, w1) in enumerate(bands_w):
            labels[h0:h1, w0:w1] = i * 3 + j
    lab = labels.reshape(h // ws, ws, w // ws, ws).permute(0, 2, 1, 3).reshape(-1, ws * ws)
    diff = lab[:, None, :] - lab[:, :, None]
    return torch.where(diff == 0, 0.0, -100.0)


def swin_attention(x, qkv_w, qkv_b, proj_w, proj_b, table, index, heads, ws, shift, res=None):
    """_shifted_window_attention (swin.py:90-255) on a batch of channels-last maps (B,H,W,C).
    Returns proj(...) [+ res] with the windows reversed and the roll undone."""
    b, h, w, c = x.shape
    d = c // heads
    shift = [0 if ws >= h else shift[0], 0 if ws >= w else shift[1]]          # swin.py:115-119
    if sum(shift) > 0:
        x = torch.roll(x, shifts=(-shift[0], -shift[1]), dims=(1, 2))          # swin.py:122-123
    nw = (h // ws) * (w // ws)
    xw = x.reshape(b, h // ws, ws, w // ws, ws, c).permute(0, 1, 3, 2, 4, 5).reshape(b * nw, ws * ws, c)
    qkv = O.linear_act(xw, qkv_w, qkv_b)                                       # swin.py:155-157
    qkv = qkv.reshape(b * nw, ws * ws, 3, heads, d).permute(2, 0, 3, 1, 4)     # swin.py:158-161
    q, k, v = qkv[0], qkv[1], qkv[2]
    n = ws * ws
    bias = table[index.long()].reshape(n, n, -1).permute(2, 0, 1)              # swin.py:36-46
    logits = (q * d ** -0.5) @ k.transpose(-1, -2) + bias                      # swin.py:180-183
    if sum(shift) > 0:
        mask = swin_shift_mask(h, w, ws, shift)                                # swin.py:185-229
        logits = (logits.reshape(b, nw, heads, n, n) + mask[None, :, None]).reshape(b * nw, heads, n, n)
    attn = O.softmax(logits, -1)                                               # swin.py:231
    out = (attn @ v).permute(0, 2, 1, 3).reshape(b * nw, n, c)                 # swin.py:235-237
    out = O.rnd(out)
    # the device adds the residual (in spatial order) inside the proj GEMM; un-window first
    out = out.reshape(b, h // ws, w // ws, ws, ws, c).permute(0, 1, 3, 2, 4, 5).reshape(b, h, w, c)
    if sum(shift) > 0:
        out = torch.roll(out, shifts=(shift[0], shift[1]), dims=(1, 2))        # swin.py:249-250
    return O.linear_act(out, proj_w, proj_b, res=res)                          # swin.py:238 (+ :572)


def swin(state_dict, x, arch="swin_t", eps=1e-5):
    """SwinTransformer.__call__ (swin.py:762-772), torchvision key order: features.0.{0 conv, 2 norm},
    per block norm1, attn.{relative_position_bias_table, relative_position_index, qkv, proj}, norm2,
    mlp.{0,3}; per merge reduction.weight, norm; norm; head."""
    dim, depths, heads, ws = _SWINS[arch]
    s = Stream(state_dict)
    pw, pb = s.take(), s.take()
    t = O.conv_bn_act(x, pw, pb, None, pw.shape[-1], 0)                        # swin.py:705-713
    t = t.permute(0, 2, 3, 1)                                                  # channels-last (B,H,W,C)
    t = O.rnd(O.layer_norm(t, s.take(), s.take(), eps))                        # LayerNorm2d, extensions_2d.py:24-28
    for i_stage, depth in enumerate(depths):
        for i_layer in range(depth):
            n1w, n1b = s.take(), s.take()
            table, index = s.take(), s.take()
            qw, qb, ow, ob = s.take(), s.take(), s.take(), s.take()
            n2w, n2b = s.take(), s.take()
            f1w, f1b, f2w, f2b = s.take(), s.take(), s.take(), s.take()
            shift = [0, 0] if i_layer % 2 == 0 else [ws // 2, ws // 2]         # swin.py:733-735
            y = O.rnd(O.layer_norm(t, n1w, n1b, eps))
            t = swin_attention(y, qw, qb, ow, ob, table, index, heads[i_stage], ws, shift, res=t)  # swin.py:572-574
            y = O.rnd(O.layer_norm(t, n2w, n2b, eps))
            y = O.linear_act(y, f1w, f1b, act="gelu")                          # mlps.py:61-62, tanh GELU
            t = O.linear_act(y, f2w, f2b, res=t)                               # swin.py:575
        if i_stage < len(depths) - 1:
            rw = s.take()                                                      # reduction (no bias) comes first
            nw_, nb_ = s.take(), s.take()
            t = torch.cat([t[:, 0::2, 0::2], t[:, 1::2, 0::2], t[:, 0::2, 1::2], t[:, 1::2, 1::2]], -1)  # swin.py:26-31
            t = O.rnd(O.layer_norm(t, nw_, nb_, eps))                          # swin.py:62
            t = O.linear_act(t, rw, None)                                      # swin.py:63
    t = O.rnd(O.layer_norm(t, s.take(), s.take(), eps))                        # swin.py:767
    pooled = O.rnd(t.mean(dim=(1, 2)))                                         # avgpool + ravel
    logits = O.linear_act(pooled, s.take(), s.take(), round_out=False)
    assert s.done()
    return logits


# ------------------------------------------------------------------------------------------------
# Swin Transformer v2 (swin.py:369-522 attention, :583-636 block, :68-87 patch merging)
# ------------------------------------------------------------------------------------------------
_SWINS_V2 = {
    "swin_v2_t": (96, [2, 2, 6, 2], [3, 6, 12, 24], 8),
    "swin_v2_s": (96, [2, 2, 18, 2], [3, 6, 12, 24], 8),
    "swin_v2_b": (128, [2, 2, 18, 2], [4, 8, 16, 32], 8),
}


def swin_v2_position_bias(coords_table, index, w1, b1, w2, heads, ws):
    """_ShiftedWindowAttentionV2.get_relative_position_bias (swin.py:494-504): cpb_mlp = [transpose (2,0,1),
    Linear2d(2,512), relu, Linear2d(512,heads,no bias), identity transpose] maps the (2Wh-1, 2Ww-1, 2) coordinate table
    to a CHANNEL-FIRST (heads, 2Wh-1, 2Ww-1) array, which `jnp.reshape(..., (-1, heads))` then reads row-major without
    moving the head axis last (torchvision's cpb_mlp is channels-last, so its view(-1, heads) is a true transpose-free
    flatten; the reference's is a scramble and is reproduced as such). 16 * sigmoid(table[index]) -> (heads, N, N)."""
    a, b, _ = coords_table.shape
    tok = coords_table.reshape(a * b, 2)
    hid = O.relu(O.linear(tok, w1, b1))
    out_chw = O.linear(hid, w2).t().contiguous()               # Linear2d returns (heads, A, B)
    flat = out_chw.reshape(-1, heads)
    n = ws * ws
    bias = flat[index.long()].reshape(n, n, -1).permute(2, 0, 1)
    return 16 * torch.sigmoid(bias)


def swin_attention_v2(x, logit_scale, bias, qkv_w, qkv_b, proj_w, proj_b, heads, ws, shift):
    """_shifted_window_attention with logit_scale (swin.py:146-166): k bias zeroed, cosine attention whose L2 norms run
    over AXIS 0 of the (num_windows, heads, tokens, d) arrays of ONE sample (the reference's quirk: torchvision
    normalises over d), times exp(min(logit_scale, log 100)); then bias, shift mask, softmax, PV, proj as in v1.
    x: (B,H,W,C) channels-last; returns proj(...) in spatial order."""
    b, h, w, c = x.shape
    d = c // heads
    shift = [0 if ws >= h else shift[0], 0 if ws >= w else shift[1]]
    if sum(shift) > 0:
        x = torch.roll(x, shifts=(-shift[0], -shift[1]), dims=(1, 2))
    nw = (h // ws) * (w // ws)
    n = ws * ws
    xw = x.reshape(b, h // ws, ws, w // ws, ws, c).permute(0, 1, 3, 2, 4, 5).reshape(b * nw, n, c)
    qb = qkv_b.clone()
    qb[c:2 * c] = 0                                                            # swin.py:146-153
    qkv = O.linear_act(xw, qkv_w, qb)
    qkv = qkv.reshape(b, nw, n, 3, heads, d).permute(3, 0, 1, 4, 2, 5)          # (3, B, nW, heads, N, d)
    q, k, v = qkv[0], qkv[1], qkv[2]
    scale = torch.exp(torch.clamp(logit_scale.reshape(1, 1, heads, 1, 1), max=math.log(100.0)))
    qn = O.rnd(q / torch.linalg.norm(q, ord=2, dim=1, keepdim=True) * scale)   # dim 1 = the sample's axis 0 (windows)
    kn = O.rnd(k / torch.linalg.norm(k, ord=2, dim=1, keepdim=True))
    logits = qn @ kn.transpose(-1, -2) + bias                                  # (B, nW, heads, N, N)
    if sum(shift) > 0:
        logits = logits + swin_shift_mask(h, w, ws, shift)[None, :, None]
    attn = O.softmax(logits, -1)
    out = (attn @ v).permute(0, 1, 3, 2, 4).reshape(b * nw, n, c)
    out = O.rnd(out)
    out = out.reshape(b, h // ws, w // ws, ws, ws, c).permute(0, 1, 3, 2, 4, 5).reshape(b, h, w, c)
    if sum(shift) > 0:
        out = torch.roll(out, shifts=(shift[0], shift[1]), dims=(1, 2))
    return O.linear_act(out, proj_w, proj_b)


def swin_v2(state_dict, x, arch="swin_v2_t", eps=1e-5):
    """SwinTransformer.__call__ with block=_SwinTransformerBlockV2, downsample_layer=_PatchMergingV2 (swin.py:874-946),
    torchvision swin_v2 key order: per block norm1, attn.{logit_scale, relative_coords_table, relative_position_index,
    qkv, proj, cpb_mlp.0.{weight,bias}, cpb_mlp.2.weight}, norm2, mlp.{0,3}; per merge reduction.weight, norm."""
    dim, depths, heads, ws = _SWINS_V2[arch] if isinstance(arch, str) else arch
    s = Stream(state_dict)
    pw, pb = s.take(), s.take()
    t = O.conv_bn_act(x, pw, pb, None, pw.shape[-1], 0)
    t = t.permute(0, 2, 3, 1)
    t = O.rnd(O.layer_norm(t, s.take(), s.take(), eps))
    for i_stage, depth in enumerate(depths):
        for i_layer in range(depth):
            n1w, n1b = s.take(), s.take()
            logit_scale, table, index = s.take(), s.take(), s.take()
            qw, qb, ow, ob = s.take(), s.take(), s.take(), s.take()
            c1w, c1b, c2w = s.take(), s.take(), s.take()
            n2w, n2b = s.take(), s.take()
            f1w, f1b, f2w, f2b = s.take(), s.take(), s.take(), s.take()
            shift = [0, 0] if i_layer % 2 == 0 else [ws // 2, ws // 2]
            bias = swin_v2_position_bias(table.reshape(2 * ws - 1, 2 * ws - 1, 2).float(), index, c1w, c1b, c2w,
                                         heads[i_stage], ws)
            y = swin_attention_v2(t, logit_scale.float(), bias, qw, qb, ow, ob, heads[i_stage], ws, shift)
            t = O.rnd(t + O.rnd(O.layer_norm(y, n1w, n1b, eps)))               # post-norm, swin.py:632-634
            y = O.linear_act(t, f1w, f1b, act="gelu")
            y = O.linear_act(y, f2w, f2b)
            t = O.rnd(t + O.rnd(O.layer_norm(y, n2w, n2b, eps)))               # swin.py:635
        if i_stage < len(depths) - 1:
            rw = s.take()
            nw_, nb_ = s.take(), s.take()
            t = torch.cat([t[:, 0::2, 0::2], t[:, 1::2, 0::2], t[:, 0::2, 1::2], t[:, 1::2, 1::2]], -1)
            t = O.linear_act(t, rw, None)                                      # _PatchMergingV2: reduce, THEN norm
            t = O.rnd(O.layer_norm(t, nw_, nb_, eps))                          # swin.py:84-86
    t = O.rnd(O.layer_norm(t, s.take(), s.take(), eps))
    pooled = O.rnd(t.mean(dim=(1, 2)))
    logits = O.linear_act(pooled, s.take(), s.take(), round_out=False)
    assert s.done()
    return logits
