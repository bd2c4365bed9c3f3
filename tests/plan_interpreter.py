"""TEST INFRASTRUCTURE - a torch-CPU executor for the launch plans `_engine.Plan` records.

The product has no CPU path: `ops.*` raise without a CUDA device. What this file checks is the HOST logic of the engine
without a GPU - weight packing, BatchNorm folding, residual order, concat slots, channel views, flatten permutations:
every recorded step `(ops.<fn>, kwargs)` is replayed with a plain torch implementation of that C entry's documented
contract (include/eqxv_b200.h), writing into the plan's own (CPU-resident) buffers. Outputs are compared with the
bf16-emulating oracle in tests/test_plan_lowering.py. Only tests import this module.
"""
import torch
import torch.nn.functional as F

BF16 = torch.bfloat16


def _act(y, code):
    from eqxvision_b200 import _lib as L

    if code == L.ACT_NONE:
        return y
    if code == L.ACT_RELU:
        return F.relu(y)
    if code == L.ACT_SILU:
        return F.silu(y)
    if code == L.ACT_GELU_TANH:
        return F.gelu(y, approximate="tanh")
    if code == L.ACT_HARDSWISH:
        return F.hardswish(y)
    if code == L.ACT_SIGMOID:
        return torch.sigmoid(y)
    if code == L.ACT_HARDSIGMOID:
        return F.hardsigmoid(y)
    if code == L.ACT_RELU6:
        return F.relu6(y)
    raise ValueError(code)


def _epilogue(y, bias, act, residual, res_after_act):
    """y fp32 [..., C] channels-last: + bias, then (residual, act) in the order the flag says"""
    if bias is not None:
        y = y + bias.float()[: y.shape[-1]]
    if residual is not None and not res_after_act:
        y = y + residual.float()[..., : y.shape[-1]]
    y = _act(y, act)
    if residual is not None and res_after_act:
        y = y + residual.float()[..., : y.shape[-1]]
    return y


def _store(out, y):
    out[..., : y.shape[-1]].copy_(y.to(out.dtype))


def _check_operand(t, what):
    """the alignment rules the C entries enforce on activation views (EQXV_CHECK_ARG in csrc/*.cu): 16-byte aligned
    base address and row pitch, i.e. multiples of 8 bf16 elements - checked on element offsets so that the rule is the
    same whether the replay keeps activations in bf16 or in fp32"""
    if t is None:
        return
    assert t.stride(-1) == 1, f"{what}: innermost stride must be 1"
    assert t.storage_offset() % 8 == 0, f"{what}: view starts at element {t.storage_offset()}, not 16-byte aligned"
    if t.dim() >= 2:
        assert t.stride(-2) % 8 == 0, f"{what}: pitch {t.stride(-2)} is not a multiple of 8 channels"


def nchw_to_nhwc(x, c_pad, out, **_):
    out.zero_()
    out[..., : x.shape[1]].copy_(x.permute(0, 2, 3, 1).to(out.dtype))


def nhwc_to_nchw(x, c, out, **_):
    out.copy_(x[..., :c].float().permute(0, 3, 1, 2))


def pack_stem_input(x_nchw, pad, out, **_):
    n, c, h, w = x_nchw.shape
    out.zero_()
    out[:, pad:pad + h, pad:pad + w, :c].copy_(x_nchw.permute(0, 2, 3, 1).to(out.dtype))


def _unpair(xpad4, wgt4, cout, kh):
    """pixel-pair layout (include/eqxv_b200.h, eqxv_conv_stem_c4_bf16) -> the 8-channel layout conv_stem() below reads"""
    n, hp, wu, _ = xpad4.shape
    x8 = torch.zeros((n, hp, 2 * wu, 8), dtype=xpad4.dtype)
    x8[..., :4] = xpad4.reshape(n, hp, wu, 2, 4).reshape(n, hp, 2 * wu, 4)
    w4 = wgt4.reshape(cout, kh, 16, 4)
    assert w4[:, :, 8:].abs().max() == 0, "stem(c4): the upper 32 K columns of every filter row must be zero"
    w8 = torch.zeros((cout, kh, 8, 8), dtype=wgt4.dtype)
    w8[..., :4] = w4[:, :, :8]
    return x8, w8.reshape(cout, kh * 64)


def pack_stem_input_c4(x_nchw, pad, out, **_):
    n, c, h, w = x_nchw.shape
    assert c <= 4 and w % 2 == 0
    x8 = torch.zeros((n, h + 2 * pad, w + 8, 8), dtype=out.dtype)
    pack_stem_input(x_nchw, pad, x8)
    out.copy_(x8[..., :4].reshape(n, h + 2 * pad, (w + 8) // 2, 8))


def u8_pack_stem_input_c4(x, lut, pad, out, **_):
    n, h, w, c = x.shape
    x8 = torch.zeros((n, h + 2 * pad, w + 8, 8), dtype=out.dtype)
    u8_pack_stem_input(x, lut, pad, x8)
    out.copy_(x8[..., :4].reshape(n, h + 2 * pad, (w + 8) // 2, 8))


def conv_stem(xpad, wgt, bias, n, h, w, cout, kh, kw, stride, pad, act, out, c4=False, **_):
    if c4:
        assert stride == 2 and w % 2 == 0
        xpad, wgt = _unpair(xpad, wgt, cout, kh)
    # xpad [n, h+2p, w+8, 8] with the image at (pad, pad); wgt [cout, kh, 8(s), 8(c)] with zero taps for s >= kw
    img = xpad[:, :, : w + 2 * pad, :].float().permute(0, 3, 1, 2)
    wt = wgt.float().reshape(cout, kh, 8, 8)[:, :, :kw, :].permute(0, 3, 1, 2)
    y = F.conv2d(img, wt, None, stride=stride).permute(0, 2, 3, 1)
    _store(out, _epilogue(y, bias, act, None, False))


def conv_stem_maxpool(xpad, wgt, bias, n, h, w, cout, kh, kw, stride, pad, out, **_):
    """include/eqxv_b200.h eqxv_conv_stem_maxpool_bf16: the stem convolution + ReLU (rounded to the activation dtype, as the
    separate entry stored it), then max-pool 3x3 / 2 / 1"""
    ho, wo = (h + 2 * pad - kh) // stride + 1, (w + 2 * pad - kw) // stride + 1
    assert cout == 64 and ho % 16 == 0 and wo % 8 == 0, "stem+maxpool: geometry the fused entry refuses"
    y = torch.empty((n, ho, wo, cout), dtype=out.dtype)
    conv_stem(xpad, wgt, bias, n, h, w, cout, kh, kw, stride, pad, 1, y)
    maxpool2d(y, 3, 2, 1, out)


def conv2d(x, wgt, bias, cin, cout, kh, kw, stride=1, pad=0, dil=1, act=0, residual=None, res_after_act=False,
           out=None, out_f32=False, grouped_block64=False, k_tail_shift=False, **_):
    assert cin % 8 == 0, "conv: cin must be a multiple of 8"
    if k_tail_shift:   # include/eqxv_b200.h EQXV_FLAG_K_TAIL_SHIFT: the last chunk's overlap columns must be zero
        from eqxvision_b200 import _pack

        assert not grouped_block64 and cin > 64 and cin % 64 != 0 and wgt.shape[1] == kh * kw * 64 * -(-cin // 64)
        kc, off = -(-cin // 64), 64 * -(-cin // 64) - cin
        assert wgt.reshape(cout, kh * kw, 64 * kc)[..., 64 * (kc - 1):64 * (kc - 1) + off].abs().max() == 0
        wgt = _pack.unshift_tail(wgt, kh * kw, cin)
    for t, what in ((x, "conv x"), (out, "conv y"), (residual, "conv residual")):
        _check_operand(t, what)
    assert out.stride(2) >= cout and (residual is None or residual.stride(2) >= cout)
    xin = x[..., :cin].float().permute(0, 3, 1, 2)
    if grouped_block64:
        wt = wgt.float().reshape(cout, kh, kw, 64)
        ys = []
        for b in range(cout // 64):
            ys.append(F.conv2d(xin[:, 64 * b:64 * b + 64], wt[64 * b:64 * b + 64].permute(0, 3, 1, 2), None, stride,
                               pad, dil))
        y = torch.cat(ys, 1)
    else:
        wt = wgt.float().reshape(cout, kh, kw, cin).permute(0, 3, 1, 2)
        y = F.conv2d(xin, wt, None, stride, pad, dil)
    _store(out, _epilogue(y.permute(0, 2, 3, 1), bias, act, residual, res_after_act))


def bottleneck64(t1, w2, b2, w3, b3, residual=None, x0=None, out=None, w1n=None, b1n=None, next_out=None, **_):
    """include/eqxv_b200.h, eqxv_bottleneck64_fused_bf16: conv3x3 -> conv1x1 + shortcut -> (next block's conv1x1), with the
    layer-by-layer path's rounding points (t2, y, next stored in the activation dtype)"""
    assert (residual is None) != (x0 is None) and (w1n is None) == (next_out is None)
    for t, what in ((t1, "bneck t1"), (residual, "bneck residual"), (x0, "bneck x0"), (out, "bneck y"), (next_out, "bneck next")):
        _check_operand(t, what)
    assert w2.shape == (64, 576) and w3.shape == (256, 128 if x0 is not None else 64) and out.shape[-1] >= 256
    t2 = torch.empty(t1.shape[:3] + (64,), dtype=out.dtype)
    conv2d(t1, w2, b2, 64, 64, 3, 3, 1, 1, 1, 1, None, False, t2)
    a = t2 if x0 is None else torch.cat([t2, x0[..., :64].to(t2.dtype)], dim=-1)
    y = a.float() @ w3.float().t()
    _store(out, _epilogue(y, b3, 1, residual, False))
    if w1n is not None:
        assert w1n.shape in ((64, 256), (128, 256)) and next_out.shape[-1] >= w1n.shape[0]
        _store(next_out, _epilogue(out[..., :256].float() @ w1n.float().t(), b1n, 1, None, False))


def gemm(a, wgt, bias, act=0, residual=None, res_after_act=False, out=None, out_f32=False, **_):
    assert a.shape[1] % 8 == 0 and wgt.shape[1] == a.shape[1], "gemm: k must be a multiple of 8 and match the filter"
    for t, what in ((a, "gemm a"), (residual, "gemm residual")) + (() if out_f32 else ((out, "gemm out"),)):
        _check_operand(t, what)
    y = a.float() @ wgt.float().t()
    _store(out, _epilogue(y, bias, act, residual, res_after_act))


def gemm_rowstats(a, wgt, bias, residual, stats, out, **_):
    """include/eqxv_b200.h K5 + K7, producer: the GEMM, plus (sum, sum of squares) of each STORED row per 64-column chunk"""
    gemm(a, wgt, bias, 0, residual, False, out)
    n = wgt.shape[0]
    y = out[:, :n].float()
    for c in range(stats.shape[1]):
        blk = y[:, 64 * c:64 * c + 64]
        stats[:, c, 0] = blk.sum(1)
        stats[:, c, 1] = (blk * blk).sum(1)


def gemm_ln(a, wgt, bias, wsum, stats, eps, act, out, **_):
    """consumer: act(rstd * (a @ w'^T - mean * wsum) + bias') with mean / rstd from the producer's statistics"""
    k = a.shape[1]
    mean = stats[:, :, 0].sum(1) / k
    var = (stats[:, :, 1].sum(1) / k - mean * mean).clamp_min(0)
    rstd = torch.rsqrt(var + eps)
    acc = a.float() @ wgt.float().t()
    y = rstd[:, None] * (acc - mean[:, None] * wsum.float()[None, :]) + bias.float()[None, :]
    _store(out, _act(y, act))


def gemm_gated(a, gate, wgt, bias, rows_per_image, residual=None, out=None, k_tail_shift=False, **_):
    """include/eqxv_b200.h K5 + K11: the SE gate multiplies the A operand (product rounded to the activation dtype, as
    the separate gate pass did), then the plain GEMM + bias (+ residual)"""
    k = a.shape[1]
    if k_tail_shift:
        from eqxvision_b200 import _pack

        wgt = _pack.unshift_tail(wgt, 1, k)
    g = gate[:, :k].float().repeat_interleave(rows_per_image, 0)
    ag = (a.float() * g).to(a.dtype)
    gemm(ag, wgt, bias, 0, residual, False, out)


def dwconv(x, wgt, bias, k, stride, pad, dil, act, out, **_):
    c = x.shape[-1]
    assert c % 8 == 0 and wgt.shape[1] >= c and bias.numel() >= c, "dwconv: channels / filter pitch"
    _check_operand(x, "dwconv x")
    _check_operand(out, "dwconv y")
    wt = wgt.float()[:, :c].t().reshape(c, 1, k, k)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), wt, None, stride, pad, dil, groups=c).permute(0, 2, 3, 1)
    _store(out, _epilogue(y, bias, act, None, False))


def dwconv_pool(x, wgt, bias, k, stride, pad, act, out, pooled, workspace=None, **_):
    """include/eqxv_b200.h K3 + K11: the depthwise conv and, from the same kernel, the mean of its (stored) output"""
    dwconv(x, wgt, bias, k, stride, pad, 1, act, out)
    c = x.shape[-1]
    pooled[..., :c].copy_(out[..., :c].float().mean((1, 2)).to(pooled.dtype))


def maxpool2d(x, k, stride, pad, out, ceil_mode=False, **_):
    assert x.shape[-1] % 8 == 0 and 2 * pad <= k, "maxpool: channels must be a multiple of 8, 2*pad <= k"
    _check_operand(x, "maxpool x")
    _check_operand(out, "maxpool y")
    y = F.max_pool2d(x.float().permute(0, 3, 1, 2), k, stride, pad, ceil_mode=ceil_mode).permute(0, 2, 3, 1)
    _store(out, y)


def avgpool2d(x, k, stride, out, **_):
    _store(out, F.avg_pool2d(x.float().permute(0, 3, 1, 2), k, stride).permute(0, 2, 3, 1))


def adaptive_avgpool(x, oh, ow, out, **_):
    n, h, w, c = x.shape
    if h % oh or w % ow:   # Equinox's uneven rule (include/eqxv_b200.h K10): leading blocks one element longer
        def blocks(dim, t):
            head, blk = dim % t, dim // t
            return [(i * (blk + 1), blk + 1) if i < head else (head * (blk + 1) + (i - head) * blk, blk) for i in range(t)]
        y = torch.stack([torch.stack([x[:, y0:y0 + hy, x0:x0 + wx].float().mean((1, 2)) for x0, wx in blocks(w, ow)], 1)
                         for y0, hy in blocks(h, oh)], 1)
        _store(out, y)
        return
    y = x.float().reshape(n, oh, h // oh, ow, w // ow, c).mean((2, 4))
    _store(out, y)


def eltwise(x, scale=None, shift=None, other=None, gate=None, rows_per_image=1, act=0, out=None, **_):
    y = x.float()
    c = y.shape[-1]
    if scale is not None:
        y = y * scale.float()[:c] + shift.float()[:c]
    if other is not None:
        y = y + other.float()[..., :c]
    y = _act(y, act)
    if gate is not None:
        y = y * gate.float()[..., :c].repeat_interleave(rows_per_image, 0)
    _store(out, y)


def layernorm(x, gamma, beta, eps, out, **_):
    _store(out, F.layer_norm(x.float(), (x.shape[-1],), gamma.float(), beta.float(), eps))


def copy2d(dst, src, **_):
    dst[:, : src.shape[1]].copy_(src)


def patchify(x_nchw, p, out, **_):
    # rows = patches in raster order, columns in (c, py, px) order = the OIHW filter flattened (patch_embed.py:79)
    n, c, h, w = x_nchw.shape
    rows = x_nchw.reshape(n, c, h // p, p, w // p, p).permute(0, 2, 4, 1, 3, 5).reshape(n * (h // p) * (w // p), c * p * p)
    out.copy_(rows.to(out.dtype))


def vit_assemble_tokens(patches, cls, pos, n, np_, d, out, **_):
    tok = torch.cat([cls.float().reshape(1, 1, d).expand(n, 1, d), patches.float().reshape(n, np_, d)], 1)
    _store(out, (tok + pos.float().reshape(1, np_ + 1, d)).reshape(n * (np_ + 1), d))


def _split_qkv(qkv, groups, tokens, heads, head_dim):
    """[groups*tokens, 3*heads*head_dim], columns ordered (3, heads, head_dim) -> q, k, v [groups, heads, tokens, d]"""
    t = qkv.float().reshape(groups, tokens, 3, heads, head_dim).permute(2, 0, 3, 1, 4)
    return t[0], t[1], t[2]


def attention(qkv, images, tokens, heads, head_dim, scale, out, **_):
    q, k, v = _split_qkv(qkv, images, tokens, heads, head_dim)
    p = torch.softmax((q @ k.transpose(-1, -2)) * scale, -1)          # scale after the product (vit.py:69)
    _store(out, (p @ v).permute(0, 2, 1, 3).reshape(images * tokens, heads * head_dim))


def attention_probs(qkv, images, tokens, heads, head_dim, scale, out, **_):
    q, k, _v = _split_qkv(qkv, images, tokens, heads, head_dim)
    out.copy_(torch.softmax((q @ k.transpose(-1, -2)) * scale, -1).reshape(out.shape))


def gather_rows(x, n, tokens, row, out, **_):
    _store(out, x.float().reshape(n, tokens, -1)[:, row])


def resize_bilinear(x, oh, ow, out, **_):
    y = F.interpolate(x.float().permute(0, 3, 1, 2), size=(oh, ow), mode="bilinear", align_corners=False)
    _store(out, y.permute(0, 2, 3, 1))


def resize_bilinear_to_nchw(x, c, oh, ow, out, **_):
    out.copy_(F.interpolate(x[..., :c].float().permute(0, 3, 1, 2), size=(oh, ow), mode="bilinear",
                            align_corners=False))


def patch_merge(x, out, **_):
    _store(out, torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1).float())


def swin_v2_qk_normalize(qkv, scale_q, n, h, w, heads, head_dim, window, shift, **_):
    """include/eqxv_b200.h: in place, q and k divided by their L2 norm over the windows of each image (per head,
    window token and channel; windows of the map rolled by -shift), q times scale_q[head]"""
    c = heads * head_dim
    ws = window
    t = qkv.float().reshape(n, h, w, 3 * c)
    t = torch.roll(t, shifts=(-shift[0], -shift[1]), dims=(1, 2))
    t = t.reshape(n, h // ws, ws, w // ws, ws, 3 * c)
    qk = t[..., : 2 * c]
    norm = qk.pow(2).sum(dim=(1, 3), keepdim=True).sqrt()
    mult = torch.ones(2 * c)
    mult[:c] = scale_q.float().repeat_interleave(head_dim)
    t = torch.cat([qk / norm * mult, t[..., 2 * c:]], -1).reshape(n, h, w, 3 * c)
    t = torch.roll(t, shifts=(shift[0], shift[1]), dims=(1, 2))
    qkv.copy_(t.reshape(n * h * w, 3 * c).to(qkv.dtype))


def window_attention(qkv, bias, n, h, w, heads, head_dim, window, shift, scale, out, **_):
    """include/eqxv_b200.h K16: roll by -shift, partition into window x window tiles, softmax(q*scale k^T + bias
    + shift mask) v per (window, head), reverse the partition and the roll. Region labels are built on the rolled map
    from three bands per axis ([0, size-window), [size-window, size-shift), [size-shift, size)); pairs of tokens
    with different labels get -100."""
    c = heads * head_dim
    ws = window
    sh = [0 if ws >= h else shift[0], 0 if ws >= w else shift[1]]
    t = qkv.float().reshape(n, h, w, 3 * c)
    if sum(sh) > 0:
        t = torch.roll(t, shifts=(-sh[0], -sh[1]), dims=(1, 2))
    nw = (h // ws) * (w // ws)
    t = t.reshape(n, h // ws, ws, w // ws, ws, 3 * c).permute(0, 1, 3, 2, 4, 5).reshape(n * nw, ws * ws, 3 * c)
    q, k, v = _split_qkv(t.reshape(n * nw * ws * ws, 3 * c), n * nw, ws * ws, heads, head_dim)
    logits = (q * scale) @ k.transpose(-1, -2) + bias.float().reshape(1, heads, ws * ws, ws * ws)
    if sum(sh) > 0:
        lab = torch.zeros(h, w)
        for i, (h0, h1) in enumerate(((0, h - ws), (h - ws, h - sh[0]), (h - sh[0], h))):
            for j, (w0, w1) in enumerate(((0, w - ws), (w - ws, w - sh[1]), (w - sh[1], w))):
                lab[h0:h1, w0:w1] = 3 * i + j
        lab = lab.reshape(h // ws, ws, w // ws, ws).permute(0, 2, 1, 3).reshape(nw, ws * ws)
        mask = torch.where(lab[:, None, :] == lab[:, :, None], 0.0, -100.0)
        logits = (logits.reshape(n, nw, heads, ws * ws, ws * ws) + mask[None, :, None]).reshape(logits.shape)
    o = (torch.softmax(logits, -1) @ v).permute(0, 2, 1, 3).reshape(n, h // ws, w // ws, ws, ws, c)
    o = o.permute(0, 1, 3, 2, 4, 5).reshape(n, h, w, c)
    if sum(sh) > 0:
        o = torch.roll(o, shifts=(sh[0], sh[1]), dims=(1, 2))
    _store(out, o.reshape(n * h * w, c))


def _u8_norm(x, lut):
    """Normalize(ToTensor(pixels)) through the table (include/eqxv_b200.h, input edge): fp32 [n, h, w, c]"""
    c = x.shape[-1]
    return torch.stack([lut[ch][x[..., ch].long()] for ch in range(c)], -1)


def u8_to_nchw_f32(x, lut, out, **_):
    out.copy_(_u8_norm(x, lut).permute(0, 3, 1, 2))


def u8_pack_stem_input(x, lut, pad, out, **_):
    n, h, w, c = x.shape
    out.zero_()
    out[:, pad:pad + h, pad:pad + w, :c].copy_(_u8_norm(x, lut).to(out.dtype))


def u8_to_nhwc(x, lut, out, **_):
    out.zero_()
    out[..., : x.shape[-1]].copy_(_u8_norm(x, lut).to(out.dtype))


def u8_patchify(x, lut, p, out, **_):
    patchify(_u8_norm(x, lut).permute(0, 3, 1, 2).contiguous(), p, out)


def u8_resize_bilinear(x, oh, ow, out, **_):
    y = F.interpolate(x.permute(0, 3, 1, 2).float(), size=(oh, ow), mode="bilinear", align_corners=False)
    out.copy_(y.round().clamp_(0, 255).to(torch.uint8).permute(0, 2, 3, 1))


IMPLS = {f.__name__: f for f in (bottleneck64, conv_stem_maxpool, pack_stem_input_c4, u8_pack_stem_input_c4, gemm_gated, gemm_rowstats, gemm_ln, dwconv_pool, u8_to_nchw_f32, u8_pack_stem_input, u8_to_nhwc, u8_patchify, u8_resize_bilinear,
                                 nchw_to_nhwc, nhwc_to_nchw, pack_stem_input, conv_stem, conv2d, gemm, dwconv,
                                 maxpool2d, avgpool2d, adaptive_avgpool, eltwise, layernorm, copy2d, patchify,
                                 vit_assemble_tokens, attention, attention_probs, gather_rows, resize_bilinear,
                                 resize_bilinear_to_nchw, patch_merge, window_attention, swin_v2_qk_normalize)}


def run(net, x, method="__call__", fp32_activations=False, arena=False, **kw):
    """lower `net` for the batch `x` on a CPU-device plan and replay the recorded steps with the torch stand-ins.
    `fp32_activations`: allocate every activation buffer in fp32 (filters stay bf16 as packed), which removes the
    bf16 rounding noise from the comparison and leaves a sharp check of the lowering itself."""
    import eqxvision_b200 as eb
    from eqxvision_b200 import _engine as E
    from eqxvision_b200 import _trace as T

    class _F32Plan(E.Plan):
        def alloc(self, rows, c, geom=(), dtype=torch.bfloat16):
            return super().alloc(rows, c, geom, torch.float32)

    saved = E.BF16
    if fp32_activations:
        E.BF16 = torch.float32
    try:
        return _run(_F32Plan if fp32_activations else E.Plan, net, x, method, kw, arena)
    finally:
        E.BF16 = saved


def _run(plan_cls, net, x, method, kw, arena=False):
    import eqxvision_b200 as eb
    from eqxvision_b200 import _engine as E
    from eqxvision_b200 import _trace as T
    from eqxvision_b200.transforms import ImagesU8

    if isinstance(x, ImagesU8):    # the uint8 input edge: same lowering, the first layout kernel reads pixels
        plan = plan_cls(torch.device("cpu"), x.shape[0], tuple(x.shape[1:]),
                        u8={"mean": x.mean, "std": x.std, "raw_hw": tuple(x.pixels.shape[1:3])})
        shape, x = tuple(x.shape), x.pixels
    else:
        plan = plan_cls(torch.device("cpu"), x.shape[0], tuple(x.shape[1:]))
        shape = tuple(x.shape)
    fn = getattr(type(net), method)
    fn = getattr(fn, "__wrapped__", fn)
    kw.setdefault("key", eb.random.PRNGKey(0))
    out = fn(net, T.Sym("chw", shape[1:], T.Input()), **kw)
    syms = []
    plan.out_struct = E._flatten_out(out, syms)
    for s in syms:
        plan.add_output(s)
    if arena:   # the liveness-based activation arena of the engine (Plan.plan_memory), poisoned first
        plan.plan_memory()
        if plan.arena is not None:
            plan.arena.fill_(0x7f)
    plan.x_host_target.copy_(x)
    for step, kwargs in plan.steps:
        impl = IMPLS.get(step.__name__)
        if impl is None:
            raise NotImplementedError(f"plan interpreter: no stand-in for ops.{step.__name__}")
        impl(**kwargs)
    leaves = [o.clone().reshape(shp) for (o, shp) in plan.outputs]
    return E._unflatten_out(plan.out_struct, leaves), plan
