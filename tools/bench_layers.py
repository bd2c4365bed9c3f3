"""Per-layer timing of the igemm kernel on the ResNet-50 (B=256) and ViT-B/16 (64 img) shapes of
SURVEY.md Appendix A. Prints ms, TFLOP/s and GB/s per layer plus the weighted totals.
Usage: python tools/bench_layers.py [resnet|vit|all] [batch]
"""
from __future__ import annotations

import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eqxvision_b200 import _lib, ops  # noqa: E402

DEV = "cuda:0"


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters


def conv_layer(n, h, w, cin, cout, k, stride, dil=1, res=False, count=1, name=""):
    pad = dil * (k - 1) // 2
    x = torch.randn(n, h, w, cin, device=DEV).to(torch.bfloat16)
    wt = (torch.randn(cout, k * k * cin, device=DEV) * (k * k * cin) ** -0.5).to(torch.bfloat16)
    b = torch.randn(cout, device=DEV)
    ho, wo = ops.conv_out_size(h, k, stride, pad, dil), ops.conv_out_size(w, k, stride, pad, dil)
    r = torch.randn(n, ho, wo, cout, device=DEV).to(torch.bfloat16) if res else None
    out = torch.empty(n, ho, wo, cout, device=DEV, dtype=torch.bfloat16)
    fn = lambda: ops.conv2d(x, wt, b, cin=cin, cout=cout, kh=k, kw=k, stride=stride, pad=pad, dil=dil,
                            act=1, residual=r, out=out)
    ms = timeit(fn)
    flops = 2.0 * n * ho * wo * cout * k * k * cin
    byts = 2.0 * (n * h * w * cin / (stride * stride if k == 1 else 1) + n * ho * wo * cout * (2 if res else 1))
    print(f"{name:34s} x{count}: {ms:7.3f} ms  {flops / ms / 1e9:8.1f} TFLOP/s  {byts / ms / 1e6:8.1f} GB/s",
          flush=True)
    return ms * count, flops * count, byts * count


def gemm_layer(m, n, k, act=0, res=False, count=1, name=""):
    a = torch.randn(m, k, device=DEV).to(torch.bfloat16)
    wt = (torch.randn(n, k, device=DEV) * k ** -0.5).to(torch.bfloat16)
    b = torch.randn(n, device=DEV)
    r = torch.randn(m, n, device=DEV).to(torch.bfloat16) if res else None
    out = torch.empty(m, n, device=DEV, dtype=torch.bfloat16)
    fn = lambda: ops.gemm(a, wt, b, act=act, residual=r, out=out)
    ms = timeit(fn)
    flops = 2.0 * m * n * k
    byts = 2.0 * (m * k + m * n * (2 if res else 1) + n * k)
    print(f"{name:34s} x{count}: {ms:7.3f} ms  {flops / ms / 1e9:8.1f} TFLOP/s  {byts / ms / 1e6:8.1f} GB/s",
          flush=True)
    return ms * count, flops * count, byts * count


def resnet50(B):
    L = []
    # (h, cin, cout, k, stride, res, count, name)
    spec = [
        (56, 64, 64, 1, 1, False, 1, "l1 1x1 64->64"),
        (56, 64, 64, 3, 1, False, 3, "l1 3x3 64"),
        (56, 64, 256, 1, 1, True, 4, "l1 1x1 64->256 (+res)"),
        (56, 256, 64, 1, 1, False, 2, "l1 1x1 256->64"),
        (56, 256, 128, 1, 1, False, 1, "l2 1x1 256->128"),
        (56, 128, 128, 3, 2, False, 1, "l2 3x3 s2 128"),
        (28, 128, 512, 1, 1, True, 4, "l2 1x1 128->512 (+res)"),
        (56, 256, 512, 1, 2, False, 1, "l2 ds 1x1 s2 256->512"),
        (28, 512, 128, 1, 1, False, 3, "l2 1x1 512->128"),
        (28, 128, 128, 3, 1, False, 3, "l2 3x3 128"),
        (28, 512, 256, 1, 1, False, 1, "l3 1x1 512->256"),
        (28, 256, 256, 3, 2, False, 1, "l3 3x3 s2 256"),
        (14, 256, 1024, 1, 1, True, 6, "l3 1x1 256->1024 (+res)"),
        (28, 512, 1024, 1, 2, False, 1, "l3 ds 1x1 s2 512->1024"),
        (14, 1024, 256, 1, 1, False, 5, "l3 1x1 1024->256"),
        (14, 256, 256, 3, 1, False, 5, "l3 3x3 256"),
        (14, 1024, 512, 1, 1, False, 1, "l4 1x1 1024->512"),
        (14, 512, 512, 3, 2, False, 1, "l4 3x3 s2 512"),
        (7, 512, 2048, 1, 1, True, 3, "l4 1x1 512->2048 (+res)"),
        (14, 1024, 2048, 1, 2, False, 1, "l4 ds 1x1 s2 1024->2048"),
        (7, 2048, 512, 1, 1, False, 2, "l4 1x1 2048->512"),
        (7, 512, 512, 3, 1, False, 2, "l4 3x3 512"),
    ]
    for (h, cin, cout, k, s, res, cnt, name) in spec:
        L.append(conv_layer(B, h, h, cin, cout, k, s, res=res, count=cnt, name=name))
    # stem
    x = torch.rand(B, 3, 224, 224, device=DEV)
    xp = ops.pack_stem_input(x)
    wt = torch.randn(64, 448, device=DEV).to(torch.bfloat16)
    b = torch.randn(64, device=DEV)
    out = torch.empty(B, 112, 112, 64, device=DEV, dtype=torch.bfloat16)
    ms = timeit(lambda: ops.pack_stem_input(x, out=xp))
    print(f"{'pack_stem_input':34s} x1: {ms:7.3f} ms  {(x.numel() * 4 + xp.numel() * 2) / ms / 1e6:8.1f} GB/s")
    L.append((ms, 0, x.numel() * 4 + xp.numel() * 2))
    ms = timeit(lambda: ops.conv_stem7x7(xp, wt, b, n=B, h=224, w=224, cout=64, out=out))
    fl = 2.0 * B * 112 * 112 * 64 * 147
    print(f"{'stem 7x7 s2':34s} x1: {ms:7.3f} ms  {fl / ms / 1e9:8.1f} TFLOP/s  {(xp.numel() + out.numel()) * 2 / ms / 1e6:8.1f} GB/s")
    L.append((ms, fl, (xp.numel() + out.numel()) * 2))
    mp = torch.empty(B, 56, 56, 64, device=DEV, dtype=torch.bfloat16)
    ms = timeit(lambda: ops.maxpool2d(out, 3, 2, 1, out=mp))
    print(f"{'maxpool 3x3 s2':34s} x1: {ms:7.3f} ms  {(out.numel() + mp.numel()) * 2 / ms / 1e6:8.1f} GB/s")
    L.append((ms, 0, (out.numel() + mp.numel()) * 2))
    x7 = torch.randn(B, 7, 7, 2048, device=DEV).to(torch.bfloat16)
    gp = torch.empty(B, 1, 1, 2048, device=DEV, dtype=torch.bfloat16)
    ms = timeit(lambda: ops.adaptive_avgpool(x7, 1, 1, out=gp))
    print(f"{'global avgpool':34s} x1: {ms:7.3f} ms  {x7.numel() * 2 / ms / 1e6:8.1f} GB/s")
    L.append((ms, 0, x7.numel() * 2))
    L.append(gemm_layer(B, 1000, 2048, name="fc 2048->1000"))
    tot_ms = sum(l[0] for l in L)
    tot_fl = sum(l[1] for l in L)
    tot_by = sum(l[2] for l in L)
    print(f"ResNet-50 B={B}: sum of kernels {tot_ms:.3f} ms -> {B / tot_ms * 1e3:.0f} img/s ; "
          f"{tot_fl / tot_ms / 1e9:.1f} TFLOP/s ; {tot_by / tot_ms / 1e6:.1f} GB/s")


def vit(B):
    M = B * 197
    L = []
    L.append(gemm_layer(B * 196, 768, 768, name="patch embed", count=1))
    L.append(gemm_layer(M, 2304, 768, name="qkv", count=12))
    L.append(gemm_layer(M, 768, 768, res=True, name="proj (+res)", count=12))
    L.append(gemm_layer(M, 3072, 768, act=3, name="fc1 (gelu)", count=12))
    L.append(gemm_layer(M, 768, 3072, res=True, name="fc2 (+res)", count=12))
    qkv = torch.randn(M, 2304, device=DEV).to(torch.bfloat16)
    o = torch.empty(M, 768, device=DEV, dtype=torch.bfloat16)
    ms = timeit(lambda: ops.attention(qkv, B, 197, 12, 64, 0.125, out=o))
    fl = 4.0 * B * 12 * 197 * 197 * 64
    print(f"{'attention':34s} x12: {ms:7.3f} ms  {fl / ms / 1e9:8.1f} TFLOP/s")
    L.append((ms * 12, fl * 12, 0))
    x = torch.randn(M, 768, device=DEV).to(torch.bfloat16)
    g = torch.randn(768, device=DEV)
    y = torch.empty_like(x)
    ms = timeit(lambda: ops.layernorm(x, g, g, 1e-5, out=y))
    print(f"{'layernorm':34s} x24: {ms:7.3f} ms  {x.numel() * 4 / ms / 1e6:8.1f} GB/s")
    L.append((ms * 24, 0, x.numel() * 4 * 24))
    tot_ms = sum(l[0] for l in L)
    tot_fl = sum(l[1] for l in L)
    print(f"ViT-B/16 B={B}: sum of kernels {tot_ms:.3f} ms -> {B / tot_ms * 1e3:.0f} img/s ; "
          f"{tot_fl / tot_ms / 1e9:.1f} TFLOP/s")


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    _lib.init(0)
    if which in ("resnet", "all"):
        resnet50(int(sys.argv[2]) if len(sys.argv) > 2 else 256)
    if which in ("vit", "all"):
        vit(int(sys.argv[2]) if len(sys.argv) > 2 else 64)
