"""Thin host-side wrappers: one Python function per C-ABI kernel entry.

Tensors are torch CUDA tensors used purely as device-memory holders (pointer + shape); all
arithmetic happens in libeqxv_b200.so. Every wrapper takes an explicit `stream` (raw cudaStream_t
handle as int, 0 = legacy default stream) and never synchronises.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import Bottleneck64Desc, ConvDesc, call, ptr

BF16 = torch.bfloat16


def _check_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.EqxvError("eqxvision_b200 ops need CUDA tensors: there is no CPU fallback")


def conv_out_size(h, k, stride, pad, dil):
    return (h + 2 * pad - dil * (k - 1) - 1) // stride + 1


def conv2d(x, wgt, bias, *, cin, cout, kh, kw, stride=1, pad=0, dil=1, act=0, residual=None,
           res_after_act=False, out=None, out_f32=False, grouped_block64=False, k_tail_shift=False, stream=0):
    """x: [N,H,W,x_pitch] bf16 view (last-dim stride 1); wgt: [cout, kh*kw*cin] bf16; bias fp32 [cout].
    grouped_block64: block-diagonal grouped convolution, wgt [cout, kh*kw*64] (_pack.pack_grouped_weight).
    k_tail_shift: wgt packed by _pack.pack_conv_weight(tail_shift=True) (EQXV_FLAG_K_TAIL_SHIFT)."""
    _check_cuda(x, wgt, bias, residual, out)
    n, h, w, _ = x.shape
    x_pitch = x.stride(2)
    ho = conv_out_size(h, kh, stride, pad, dil)
    wo = conv_out_size(w, kw, stride, pad, dil)
    if out is None:
        out = torch.empty((n, ho, wo, cout), dtype=torch.float32 if out_f32 else BF16, device=x.device)
    d = ConvDesc()
    d.x, d.wgt, d.bias, d.residual, d.y = ptr(x), ptr(wgt), ptr(bias), ptr(residual), ptr(out)
    d.n, d.h, d.w, d.cin, d.cout = n, h, w, cin, cout
    d.kh, d.kw, d.stride, d.pad, d.dil = kh, kw, stride, pad, dil
    d.x_pitch, d.y_pitch = x_pitch, out.stride(2)
    d.res_pitch = residual.stride(2) if residual is not None else 0
    d.act = act
    d.flags = (_lib.FLAG_OUT_F32 if out_f32 else 0) | (_lib.FLAG_RES_AFTER_ACT if res_after_act else 0) | \
        (_lib.FLAG_GROUPED_BLOCK64 if grouped_block64 else 0) | (_lib.FLAG_K_TAIL_SHIFT if k_tail_shift else 0)
    call("eqxv_conv2d_igemm_bf16", C.byref(d), stream)
    return out


def bottleneck64(t1, w2, b2, w3, b3, *, residual=None, x0=None, out=None, w1n=None, b1n=None, next_out=None, stream=0):
    """One ResNet bottleneck with a 64-channel trunk behind its first 1x1 convolution (eqxv_bottleneck64_fused_bf16):
    out = relu(conv1x1(relu(conv3x3(t1))) + residual)  or, with x0, relu([t2 | x0] @ [w3 | wd]^T + b3);
    next_out = relu(conv1x1(out; w1n) + b1n) when w1n is given. t1 / x0 / next_out: [N,H,W,64], residual / out: [N,H,W,256]."""
    _check_cuda(t1, w2, b2, w3, b3, residual, x0, out, w1n, b1n, next_out)
    n, h, w, _ = t1.shape
    if out is None:
        out = torch.empty((n, h, w, 256), dtype=BF16, device=t1.device)
    if w1n is not None and next_out is None:
        next_out = torch.empty((n, h, w, w1n.shape[0]), dtype=BF16, device=t1.device)
    d = Bottleneck64Desc()
    d.t1, d.w2, d.b2, d.w3, d.b3 = ptr(t1), ptr(w2), ptr(b2), ptr(w3), ptr(b3)
    d.residual, d.x0, d.y = ptr(residual), ptr(x0), ptr(out)
    d.w1n, d.b1n, d.next = ptr(w1n), ptr(b1n), ptr(next_out)
    d.n, d.h, d.w = n, h, w
    d.t1_pitch, d.y_pitch = t1.stride(2), out.stride(2)
    d.res_pitch = residual.stride(2) if residual is not None else 0
    d.x0_pitch = x0.stride(2) if x0 is not None else 0
    d.next_pitch = next_out.stride(2) if next_out is not None else 0
    d.next_channels = w1n.shape[0] if w1n is not None else 0
    call("eqxv_bottleneck64_fused_bf16", C.byref(d), stream)
    return out if next_out is None else (out, next_out)


def gemm(a, wgt, bias, *, act=0, residual=None, out=None, out_f32=False, res_after_act=False, stream=0):
    """out[m, n] = act(a[m, :k] @ wgt[n, :k]^T + bias (+ residual)); a/out row-major with pitch."""
    _check_cuda(a, wgt, bias, residual, out)
    m, k = a.shape
    n = wgt.shape[0]
    if out is None:
        out = torch.empty((m, n), dtype=torch.float32 if out_f32 else BF16, device=a.device)
    flags = (_lib.FLAG_OUT_F32 if out_f32 else 0) | (_lib.FLAG_RES_AFTER_ACT if res_after_act else 0)
    call("eqxv_gemm_bias_act_res_bf16", ptr(a), a.stride(0), ptr(wgt), ptr(bias), ptr(residual),
         residual.stride(0) if residual is not None else 0, ptr(out), out.stride(0), m, n, k, act, flags,
         stream)
    return out


def gemm_rowstats(a, wgt, bias, *, residual, stats, out=None, act=0, res_after_act=False, out_f32=False, stream=0):
    """out = a @ wgt^T + bias + residual, and stats[m, ceil(n/64), 2] = (sum, sum of squares) of every stored row per
    64-column chunk: the producer half of a LayerNorm folded into its neighbours (eqxv_gemm_res_rowstats_bf16)"""
    _check_cuda(a, wgt, bias, residual, out, stats)
    if act != 0 or res_after_act or out_f32 or residual is None:
        raise _lib.EqxvError("gemm_rowstats: plain bf16 GEMM + bias + residual only")
    m, k = a.shape
    n = wgt.shape[0]
    if out is None:
        out = torch.empty((m, n), dtype=BF16, device=a.device)
    if stats.dtype != torch.float32 or tuple(stats.shape) != (m, (n + 63) // 64, 2) or not stats.is_contiguous():
        raise _lib.EqxvError("gemm_rowstats: stats must be a contiguous fp32 [m, ceil(n/64), 2] tensor")
    call("eqxv_gemm_res_rowstats_bf16", ptr(a), a.stride(0), ptr(wgt), ptr(bias), ptr(residual), residual.stride(0),
         ptr(out), out.stride(0), ptr(stats), m, n, k, stream)
    return out


def gemm_ln(a, wgt, bias, wsum, stats, eps, *, act=0, out=None, stream=0):
    """out = act(LayerNorm(a) @ W^T + b) without materialising the LayerNorm: wgt = W * gamma, bias = b + W @ beta,
    wsum = row sums of wgt, stats from gemm_rowstats of the GEMM that produced `a` (eqxv_gemm_ln_act_bf16)"""
    _check_cuda(a, wgt, bias, wsum, stats, out)
    m, k = a.shape
    n = wgt.shape[0]
    if out is None:
        out = torch.empty((m, n), dtype=BF16, device=a.device)
    if stats.dtype != torch.float32 or tuple(stats.shape) != (m, (k + 63) // 64, 2) or not stats.is_contiguous():
        raise _lib.EqxvError("gemm_ln: stats must be a contiguous fp32 [m, ceil(k/64), 2] tensor")
    call("eqxv_gemm_ln_act_bf16", ptr(a), a.stride(0), ptr(wgt), ptr(bias), ptr(wsum), ptr(stats), stats.shape[1],
         float(eps), ptr(out), out.stride(0), m, n, k, act, stream)
    return out


def gemm_gated(a, gate, wgt, bias, *, rows_per_image, residual=None, out=None, k_tail_shift=False, stream=0):
    """out = (a * gate[row // rows_per_image]) @ wgt^T + bias (+ residual): the SE gate applied to the A operand inside
    the projection GEMM (eqxv_gemm_gated_bf16); a [m, k], gate [images, >= k] bf16"""
    _check_cuda(a, gate, wgt, bias, residual, out)
    m, k = a.shape
    n = wgt.shape[0]
    if out is None:
        out = torch.empty((m, n), dtype=BF16, device=a.device)
    call("eqxv_gemm_gated_bf16", ptr(a), a.stride(0), ptr(gate), gate.stride(0), rows_per_image, ptr(wgt), ptr(bias),
         ptr(residual), residual.stride(0) if residual is not None else 0, ptr(out), out.stride(0), m, n, k,
         _lib.FLAG_K_TAIL_SHIFT if k_tail_shift else 0, stream)
    return out


def pack_stem_input(x_nchw, pad=3, out=None, stream=0):
    _check_cuda(x_nchw, out)
    n, c, h, w = x_nchw.shape
    assert c <= 8 and x_nchw.dtype == torch.float32 and x_nchw.is_contiguous()
    if out is None:
        out = torch.empty((n, h + 2 * pad, w + 8, 8), dtype=BF16, device=x_nchw.device)
    call("eqxv_pack_stem_input", ptr(x_nchw), ptr(out), n, c, h, w, pad, stream)
    return out


def pack_stem_input_c4(x_nchw, pad=3, out=None, stream=0):
    """fp32 NCHW (<= 4 channels, even width) -> the pixel-pair layout [n, h+2*pad, (w+8)/2, 8] of conv_stem(c4=True)"""
    _check_cuda(x_nchw, out)
    n, c, h, w = x_nchw.shape
    assert c <= 4 and w % 2 == 0 and x_nchw.dtype == torch.float32 and x_nchw.is_contiguous()
    if out is None:
        out = torch.empty((n, h + 2 * pad, (w + 8) // 2, 8), dtype=BF16, device=x_nchw.device)
    call("eqxv_pack_stem_input_c4", ptr(x_nchw), ptr(out), n, c, h, w, pad, stream)
    return out


def conv_stem(xpad, wgt, bias, *, n, h, w, cout, kh=7, kw=7, stride=2, pad=3, act=1, out=None, c4=False, stream=0):
    _check_cuda(xpad, wgt, bias, out)
    ho, wo = (h + 2 * pad - kh) // stride + 1, (w + 2 * pad - kw) // stride + 1
    if out is None:
        out = torch.empty((n, ho, wo, cout), dtype=BF16, device=xpad.device)
    call("eqxv_conv_stem_c4_bf16" if c4 else "eqxv_conv_stem_bf16", ptr(xpad), ptr(wgt), ptr(bias), ptr(out), n, h, w, cout,
         kh, kw, stride, pad, out.stride(2), act, stream)
    return out


def conv_stem_maxpool(xpad, wgt, bias, *, n, h, w, cout, kh=7, kw=7, stride=2, pad=3, out=None, stream=0):
    """first-layer convolution + ReLU + max-pool 3x3 / 2 / 1 in one kernel (eqxv_conv_stem_maxpool_bf16); out:
    [n, ho/2, wo/2, cout] with ho, wo the convolution's output extents"""
    _check_cuda(xpad, wgt, bias, out)
    ho, wo = (h + 2 * pad - kh) // stride + 1, (w + 2 * pad - kw) // stride + 1
    if out is None:
        out = torch.empty((n, ho // 2, wo // 2, cout), dtype=BF16, device=xpad.device)
    call("eqxv_conv_stem_maxpool_bf16", ptr(xpad), ptr(wgt), ptr(bias), ptr(out), n, h, w, cout, kh, kw, stride, pad,
         out.stride(2), stream)
    return out


def conv_stem7x7(xpad, wgt, bias, *, n, h, w, cout, act=1, out=None, stream=0):
    return conv_stem(xpad, wgt, bias, n=n, h=h, w=w, cout=cout, kh=7, kw=7, stride=2, pad=3, act=act, out=out,
                     stream=stream)


def nchw_to_nhwc(x, c_pad=None, out=None, stream=0):
    _check_cuda(x, out)
    n, c, h, w = x.shape
    assert x.dtype == torch.float32 and x.is_contiguous()
    c_pad = c_pad or (c + 7) // 8 * 8
    if out is None:
        out = torch.empty((n, h, w, c_pad), dtype=BF16, device=x.device)
    call("eqxv_nchw_f32_to_nhwc_bf16", ptr(x), ptr(out), n, c, h, w, c_pad, stream)
    return out


def nhwc_to_nchw(x, c=None, out=None, stream=0):
    _check_cuda(x, out)
    n, h, w, cc = x.shape
    c = c or cc
    if out is None:
        out = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    call("eqxv_nhwc_bf16_to_nchw_f32", ptr(x), ptr(out), n, c, h, w, x.stride(2), stream)
    return out


def pool_out_size(size: int, k: int, stride: int, pad: int, ceil_mode: bool = False) -> int:
    """output extent of a pooling window sweep; ceil mode follows torch's rule (see eqxv_maxpool2d_ceil_nhwc_bf16)"""
    if not ceil_mode:
        return (size + 2 * pad - k) // stride + 1
    o = -((size + 2 * pad - k) // -stride) + 1
    return o - 1 if (o - 1) * stride >= size + pad else o


def maxpool2d(x, k, stride, pad, out=None, stream=0, ceil_mode=False):
    _check_cuda(x, out)
    n, h, w, c = x.shape
    ho, wo = pool_out_size(h, k, stride, pad, ceil_mode), pool_out_size(w, k, stride, pad, ceil_mode)
    if out is None:
        out = torch.empty((n, ho, wo, c), dtype=BF16, device=x.device)
    elif tuple(out.shape[1:3]) != (ho, wo):
        raise _lib.EqxvError(f"maxpool2d: output map {tuple(out.shape[1:3])} does not match {(ho, wo)}")
    call("eqxv_maxpool2d_ceil_nhwc_bf16" if ceil_mode else "eqxv_maxpool2d_nhwc_bf16", ptr(x), ptr(out), n, h, w, c,
         k, stride, pad, x.stride(2), out.stride(2), stream)
    return out


def avgpool2d(x, k, stride, out=None, stream=0):
    _check_cuda(x, out)
    n, h, w, c = x.shape
    ho, wo = (h - k) // stride + 1, (w - k) // stride + 1
    if out is None:
        out = torch.empty((n, ho, wo, c), dtype=BF16, device=x.device)
    call("eqxv_avgpool2d_nhwc_bf16", ptr(x), ptr(out), n, h, w, c, k, stride, x.stride(2),
         out.stride(2), stream)
    return out


def adaptive_avgpool(x, oh, ow, out=None, stream=0):
    _check_cuda(x, out)
    n, h, w, c = x.shape
    if out is None:
        out = torch.empty((n, oh, ow, c), dtype=BF16, device=x.device)
    call("eqxv_adaptive_avgpool_nhwc_bf16", ptr(x), ptr(out), n, h, w, c, oh, ow, x.stride(2),
         out.stride(2), stream)
    return out


def layernorm(x, gamma, beta, eps=1e-5, out=None, stream=0):
    _check_cuda(x, gamma, beta, out)
    rows, d = x.shape
    if out is None:
        out = torch.empty((rows, d), dtype=BF16, device=x.device)
    call("eqxv_layernorm_bf16", ptr(x), x.stride(0), ptr(gamma), ptr(beta), ptr(out), out.stride(0),
         rows, d, float(eps), stream)
    return out


def attention(qkv, images, tokens, heads, head_dim, scale, out=None, stream=0):
    _check_cuda(qkv, out)
    if out is None:
        out = torch.empty((images * tokens, heads * head_dim), dtype=BF16, device=qkv.device)
    call("eqxv_attention_fwd_bf16", ptr(qkv), ptr(out), None, images, tokens, heads, head_dim,
         float(scale), stream)
    return out


def attention_probs(qkv, images, tokens, heads, head_dim, scale, out=None, stream=0):
    """fp32 softmax(q k^T * scale) [images, heads, tokens, tokens] (vit.py:70, return_attention path)"""
    _check_cuda(qkv, out)
    if out is None:
        out = torch.empty((images, heads, tokens, tokens), dtype=torch.float32, device=qkv.device)
    assert out.dtype == torch.float32 and out.is_contiguous()
    call("eqxv_attention_fwd_bf16", ptr(qkv), None, ptr(out), images, tokens, heads, head_dim, float(scale), stream)
    return out


def patchify(x_nchw, p, out=None, stream=0):
    _check_cuda(x_nchw, out)
    n, c, h, w = x_nchw.shape
    assert x_nchw.dtype == torch.float32 and x_nchw.is_contiguous()
    if out is None:
        out = torch.empty((n * (h // p) * (w // p), c * p * p), dtype=BF16, device=x_nchw.device)
    call("eqxv_patchify_nchw_f32_bf16", ptr(x_nchw), ptr(out), n, c, h, w, p, stream)
    return out


def vit_assemble_tokens(patches, cls, pos, n, np_, d, out=None, stream=0):
    _check_cuda(patches, cls, pos, out)
    if out is None:
        out = torch.empty((n * (np_ + 1), d), dtype=BF16, device=patches.device)
    call("eqxv_vit_assemble_tokens_bf16", ptr(patches), ptr(cls), ptr(pos), ptr(out), n, np_, d, stream)
    return out


def gather_rows(x, n, tokens, row, out=None, stream=0):
    _check_cuda(x, out)
    d = x.shape[1]
    if out is None:
        out = torch.empty((n, d), dtype=BF16, device=x.device)
    call("eqxv_gather_rows_bf16", ptr(x), x.stride(0), ptr(out), out.stride(0), n, tokens, row, d, stream)
    return out


def dwconv(x, wgt, bias, *, k, stride=1, pad=0, dil=1, act=0, out=None, tile=False, stream=0):
    """depthwise conv; x [N,H,W,C] bf16 view, wgt fp32 [k*k, w_pitch], bias fp32 [C];
    tile=True forces the shared-memory stencil kernel (eqxv_dwconv_tile_bf16)"""
    _check_cuda(x, wgt, bias, out)
    n, h, w, c = x.shape
    ho, wo = conv_out_size(h, k, stride, pad, dil), conv_out_size(w, k, stride, pad, dil)
    if out is None:
        out = torch.empty((n, ho, wo, c), dtype=BF16, device=x.device)
    call("eqxv_dwconv_tile_bf16" if tile else "eqxv_dwconv_bn_act_bf16", ptr(x), ptr(wgt), ptr(bias), ptr(out),
         n, h, w, c, k, stride, pad, dil,
         x.stride(2), out.stride(2), wgt.stride(0), act, stream)
    return out


def dwconv_pool_workspace_bytes(n, h, w, c, k, stride, pad) -> int:
    b = C.c_int64()
    call("eqxv_dwconv_pool_workspace_bytes", n, h, w, c, k, stride, pad, C.byref(b))
    return b.value


def dwconv_pool(x, wgt, bias, *, k, stride=1, pad=0, act=0, out=None, pooled=None, workspace=None, stream=0):
    """depthwise conv + BN + act AND the SE squeeze: pooled[n, c] = mean over the output map (eqxv_dwconv_bn_act_pool_bf16).
    `workspace`: zero-initialised uint8 tensor of dwconv_pool_workspace_bytes(...) bytes (left zeroed by the kernel)."""
    _check_cuda(x, wgt, bias, out, pooled, workspace)
    n, h, w, c = x.shape
    ho, wo = conv_out_size(h, k, stride, pad, 1), conv_out_size(w, k, stride, pad, 1)
    if out is None:
        out = torch.empty((n, ho, wo, c), dtype=BF16, device=x.device)
    if pooled is None:
        pooled = torch.empty((n, c), dtype=BF16, device=x.device)
    need = dwconv_pool_workspace_bytes(n, h, w, c, k, stride, pad)
    if workspace is None and need:
        workspace = torch.zeros(need, dtype=torch.uint8, device=x.device)
    call("eqxv_dwconv_bn_act_pool_bf16", ptr(x), ptr(wgt), ptr(bias), ptr(out), ptr(pooled), ptr(workspace),
         0 if workspace is None else workspace.numel(), n, h, w, c, k, stride, pad, x.stride(2), out.stride(2),
         wgt.stride(0), pooled.stride(0), act, stream)
    return out, pooled


def eltwise(x, *, scale=None, shift=None, other=None, gate=None, rows_per_image=1, act=0, out=None, stream=0):
    """rows x c: out = act(x*scale + shift + other) * gate[row // rows_per_image]"""
    _check_cuda(x, scale, shift, other, gate, out)
    rows, c = x.shape
    if out is None:
        out = torch.empty((rows, c), dtype=BF16, device=x.device)
    call("eqxv_eltwise_bf16", ptr(x), ptr(scale), ptr(shift), ptr(other), ptr(gate), ptr(out), rows, c,
         x.stride(0), other.stride(0) if other is not None else 0, gate.stride(0) if gate is not None else 0,
         out.stride(0), rows_per_image, act, stream)
    return out


def resize_bilinear_to_nchw(x, c, oh, ow, out=None, stream=0):
    _check_cuda(x, out)
    n, h, w, _ = x.shape
    if out is None:
        out = torch.empty((n, c, oh, ow), dtype=torch.float32, device=x.device)
    call("eqxv_resize_bilinear_nhwc_bf16_to_nchw_f32", ptr(x), ptr(out), n, c, h, w, oh, ow, x.stride(2), stream)
    return out


def resize_bilinear(x, oh, ow, out=None, stream=0):
    _check_cuda(x, out)
    n, h, w, c = x.shape
    if out is None:
        out = torch.empty((n, oh, ow, c), dtype=BF16, device=x.device)
    call("eqxv_resize_bilinear_nhwc_bf16", ptr(x), ptr(out), n, c, h, w, oh, ow, x.stride(2), out.stride(2), stream)
    return out


def copy2d(dst, src, stream=0):
    """dst[:, :c] = src[:, :c] for 2-D views with arbitrary row pitch"""
    _check_cuda(dst, src)
    rows, c = src.shape
    es = src.element_size()
    call("eqxv_copy2d_async", ptr(dst), dst.stride(0) * es, ptr(src), src.stride(0) * es, c * es, rows, stream)
    return dst


def window_attention(qkv, bias, *, n, h, w, heads, head_dim, window, shift, scale, out=None, stream=0):
    _check_cuda(qkv, bias, out)
    if out is None:
        out = torch.empty((n * h * w, heads * head_dim), dtype=BF16, device=qkv.device)
    call("eqxv_window_attention_bf16", ptr(qkv), ptr(bias), ptr(out), n, h, w, heads, head_dim, window,
         shift[0], shift[1], float(scale), stream)
    return out


def swin_v2_qk_normalize(qkv, scale_q, *, n, h, w, heads, head_dim, window, shift, stream=0):
    """in place: q, k /= L2 norm over the windows of each image; q *= scale_q[head] (swin.py:158-166)"""
    _check_cuda(qkv, scale_q)
    if qkv.stride(0) != 3 * heads * head_dim:
        raise _lib.EqxvError("swin_v2_qk_normalize expects a dense qkv matrix")
    call("eqxv_swin_v2_qk_normalize_bf16", ptr(qkv), ptr(scale_q), n, h, w, heads, head_dim, window, shift[0],
         shift[1], stream)
    return qkv


def patch_merge(x, out=None, stream=0):
    _check_cuda(x, out)
    n, h, w, c = x.shape
    if out is None:
        out = torch.empty((n, h // 2, w // 2, 4 * c), dtype=BF16, device=x.device)
    call("eqxv_patch_merge_bf16", ptr(x), ptr(out), n, h, w, c, x.stride(2), out.stride(2), stream)
    return out


# ---- input edge: uint8 HWC images, ToTensor + Normalize fused (tests/conftest.py:20-41 of the reference) ----------
def _check_u8(x, lut):
    _check_cuda(x, lut)
    if x.dtype != torch.uint8 or x.dim() != 4 or not x.is_contiguous() or x.shape[3] > 4:
        raise _lib.EqxvError("input edge: expected a contiguous uint8 [N, H, W, C<=4] image batch")
    if lut.dtype != torch.float32 or tuple(lut.shape) != (x.shape[3], 256) or not lut.is_contiguous():
        raise _lib.EqxvError("input edge: lut must be fp32 [C, 256]")


def u8_to_nchw_f32(x, lut, out=None, stream=0):
    _check_u8(x, lut)
    n, h, w, c = x.shape
    if out is None:
        out = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
    call("eqxv_u8hwc_to_nchw_f32", ptr(x), ptr(lut), ptr(out), n, h, w, c, stream)
    return out


def u8_pack_stem_input(x, lut, pad=3, out=None, stream=0):
    _check_u8(x, lut)
    n, h, w, c = x.shape
    if out is None:
        out = torch.empty((n, h + 2 * pad, w + 8, 8), dtype=BF16, device=x.device)
    call("eqxv_u8hwc_pack_stem_input", ptr(x), ptr(lut), ptr(out), n, h, w, c, pad, stream)
    return out


def u8_pack_stem_input_c4(x, lut, pad=3, out=None, stream=0):
    _check_u8(x, lut)
    n, h, w, c = x.shape
    if out is None:
        out = torch.empty((n, h + 2 * pad, (w + 8) // 2, 8), dtype=BF16, device=x.device)
    call("eqxv_u8hwc_pack_stem_input_c4", ptr(x), ptr(lut), ptr(out), n, h, w, c, pad, stream)
    return out


def u8_to_nhwc(x, lut, out=None, stream=0):
    _check_u8(x, lut)
    n, h, w, c = x.shape
    if out is None:
        out = torch.empty((n, h, w, 8), dtype=BF16, device=x.device)
    call("eqxv_u8hwc_to_nhwc_bf16", ptr(x), ptr(lut), ptr(out), n, h, w, c, stream)
    return out


def u8_patchify(x, lut, p, out=None, stream=0):
    _check_u8(x, lut)
    n, h, w, c = x.shape
    if out is None:
        out = torch.empty((n * (h // p) * (w // p), c * p * p), dtype=BF16, device=x.device)
    call("eqxv_u8hwc_patchify_bf16", ptr(x), ptr(lut), ptr(out), n, h, w, c, p, stream)
    return out


def u8_resize_bilinear(x, oh, ow, out=None, stream=0):
    _check_cuda(x, out)
    if x.dtype != torch.uint8 or x.dim() != 4 or not x.is_contiguous():
        raise _lib.EqxvError("u8_resize_bilinear: expected a contiguous uint8 [N, H, W, C] batch")
    n, h, w, c = x.shape
    if out is None:
        out = torch.empty((n, oh, ow, c), dtype=torch.uint8, device=x.device)
    call("eqxv_u8hwc_resize_bilinear", ptr(x), ptr(out), n, h, w, c, oh, ow, stream)
    return out
