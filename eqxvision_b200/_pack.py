"""Host-side weight packing (one-off, at plan-build time).

Same positional/reshape contract as the reference keeps its weights in (OIHW conv filters,
(out,in) Linear matrices — utils.py:196-197 never transposes); the device layout is what the
kernels want:
  * conv filter OIHW fp32  ->  [O, kh*kw*I_pad] bf16, K-major (tap-major, channel-minor), with the
    inference BatchNorm scale gamma/sqrt(var+eps) folded into the filter in fp32 before rounding;
  * the remaining per-channel shift  beta - mean*scale (+ conv bias * scale)  stays fp32 and is
    added in the GEMM epilogue.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch


def round_up(v: int, m: int) -> int:
    return (v + m - 1) // m * m


def fold_bn(weight: torch.Tensor, bias: Optional[torch.Tensor], bn) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """weight (O, ...) fp32, bias (O,) or None, bn a nn.BatchNorm or None -> (w', b')"""
    w = weight.detach().double()
    b = None if bias is None else bias.detach().double().reshape(-1)
    if bn is not None:
        scale, shift = bn.folded()
        scale, shift = scale.double(), shift.double()
        w = w * scale.reshape(-1, *([1] * (w.dim() - 1)))
        b = shift if b is None else b * scale + shift
    return w.float(), (None if b is None else b.float())


def tail_shift_applies(cin_pad: int) -> bool:
    """EQXV_FLAG_K_TAIL_SHIFT (include/eqxv_b200.h) is defined for dense filters with cin > 64, cin % 64 != 0"""
    return cin_pad > 64 and cin_pad % 64 != 0


def pack_conv_weight(w_oihw: torch.Tensor, cin_pad: int, tail_shift: bool = False) -> torch.Tensor:
    """OIHW fp32 -> [O, kh*kw*cin_pad] bf16 (zero-padded input channels).
    tail_shift (EQXV_FLAG_K_TAIL_SHIFT): [O, kh*kw*64*kc], kc = ceil(cin_pad / 64); per tap the last 64-wide chunk holds
    channels [cin_pad - 64, cin_pad) - the chunk the kernel then fetches, entirely inside the tensor - with the columns
    that repeat the previous chunk zeroed, so every product is still counted exactly once."""
    o, i, kh, kw = w_oihw.shape
    w = w_oihw.permute(0, 2, 3, 1).contiguous()  # O, kh, kw, I
    if cin_pad != i:
        wp = torch.zeros(o, kh, kw, cin_pad, dtype=w.dtype)
        wp[..., :i] = w
        w = wp
    if tail_shift:
        assert tail_shift_applies(cin_pad)
        kc = -(-cin_pad // 64)
        head = 64 * (kc - 1)
        off = 64 * kc - cin_pad                      # columns of the last chunk that overlap the previous one
        ws = torch.zeros(o, kh, kw, 64 * kc, dtype=w.dtype)
        ws[..., :head] = w[..., :head]
        ws[..., head + off:] = w[..., head:]
        return ws.reshape(o, kh * kw * 64 * kc).to(torch.bfloat16).contiguous()
    return w.reshape(o, kh * kw * cin_pad).to(torch.bfloat16).contiguous()


def unshift_tail(wp: torch.Tensor, taps: int, cin_pad: int) -> torch.Tensor:
    """inverse of pack_conv_weight(tail_shift=True) on the packed matrix: [O, taps*64*kc] -> [O, taps*cin_pad]"""
    o = wp.shape[0]
    kc = -(-cin_pad // 64)
    head, off = 64 * (kc - 1), 64 * kc - cin_pad
    ws = wp.reshape(o, taps, 64 * kc)
    return torch.cat([ws[..., :head], ws[..., head + off:]], dim=-1).reshape(o, taps * cin_pad).contiguous()


def pack_grouped_weight(w: torch.Tensor, groups: int) -> torch.Tensor:
    """Grouped filter [O, I/G, kh, kw] (equinox.nn.Conv2d(groups=G), resnet.py:19-23) -> block-diagonal
    [O, kh*kw*64] bf16: output channel o lives in the 64-channel block b = o // 64 and its K axis spans the
    input channels [64b, 64b+64) of every tap; only the channels of o's own group are non-zero.
    Requires I == O (per group), O % 64 == 0 and 64 % (I/G) == 0 (EQXV_FLAG_GROUPED_BLOCK64)."""
    o, ig, kh, kw = w.shape
    og = o // groups
    assert og == ig and o % 64 == 0 and 64 % ig == 0, "unsupported group geometry"
    out = torch.zeros(o, kh, kw, 64, dtype=torch.float32)
    oc = torch.arange(o)
    first = (oc // og) * ig % 64          # first input channel of o's group inside its 64-block
    wt = w.permute(0, 2, 3, 1).float()    # O, kh, kw, I/G
    for j in range(ig):
        out[oc, :, :, first + j] = wt[:, :, :, j]
    return out.reshape(o, kh * kw * 64).to(torch.bfloat16).contiguous()


def expand_grouped_weight(w: torch.Tensor, groups: int) -> torch.Tensor:
    """Grouped filter [O, I/G, kh, kw] -> dense [O, I, kh, kw] with zeros outside each output channel's own group
    (the generic fallback for group geometries the 64-channel block layout does not cover)."""
    o, ig, kh, kw = w.shape
    og = o // groups
    dense = torch.zeros(o, ig * groups, kh, kw, dtype=w.dtype)
    for g in range(groups):
        dense[g * og:(g + 1) * og, g * ig:(g + 1) * ig] = w[g * og:(g + 1) * og]
    return dense


def pack_stem_weight(w_oihw: torch.Tensor) -> torch.Tensor:
    """[O, I<=8, kh<=8, kw<=8] -> [O, kh(r), 8(s), 8(c)] bf16 with zero taps/channels
    (see eqxv_conv_stem_bf16)"""
    o, i, kh, kw = w_oihw.shape
    assert i <= 8 and kh <= 8 and kw <= 8
    wp = torch.zeros(o, kh, 8, 8, dtype=torch.float32)
    wp[:, :, :kw, :i] = w_oihw.permute(0, 2, 3, 1)
    return wp.reshape(o, kh * 64).to(torch.bfloat16).contiguous()


def pack_stem_weight_c4(w_oihw: torch.Tensor) -> torch.Tensor:
    """[O, I<=4, kh<=8, kw<=8] -> [O, kh(r), 64] bf16, K index 4*s + c (zeros beyond kw / cin and in the upper 32
    columns): the filter of eqxv_conv_stem_c4_bf16 (pixel-pair layout)"""
    o, i, kh, kw = w_oihw.shape
    assert i <= 4 and kh <= 8 and kw <= 8
    wp = torch.zeros(o, kh, 16, 4, dtype=torch.float32)
    wp[:, :, :kw, :i] = w_oihw.permute(0, 2, 3, 1)
    return wp.reshape(o, kh * 64).to(torch.bfloat16).contiguous()


def pack_linear_weight(w: torch.Tensor, in_pad: int) -> torch.Tensor:
    o, i = w.shape
    if in_pad != i:
        wp = torch.zeros(o, in_pad, dtype=w.dtype)
        wp[:, :i] = w
        w = wp
    return w.to(torch.bfloat16).contiguous()


def pack_depthwise_weight(w: torch.Tensor, c_pad: int) -> torch.Tensor:
    """[C,1,kh,kw] fp32 -> [kh*kw, c_pad] fp32 (tap-major so a thread reads 8 contiguous channels)"""
    c, one, kh, kw = w.shape
    assert one == 1
    out = torch.zeros(kh * kw, c_pad, dtype=torch.float32)
    out[:, :c] = w.reshape(c, kh * kw).t()
    return out.contiguous()
