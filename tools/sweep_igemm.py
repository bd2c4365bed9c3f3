"""Sweep of the igemm tiling choices (CTA pair on/off, block_n, epilogue warps per quadrant) over the
ResNet-50 (B=256) stage-2..4 layer shapes and the ViT-B/16 (64 img) GEMMs; prints the best configuration per
layer next to the library's default choice. Usage: python tools/sweep_igemm.py [resnet|vit|all]"""
import itertools
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eqxvision_b200 import _lib, ops  # noqa: E402

_lib.init(0)
DEV = "cuda:0"
KEYS = ("EQXV_NO_PAIR", "EQXV_FORCE_PAIR", "EQXV_BLOCK_N", "EQXV_EPI_SUB")


def timeit(fn, iters=12, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e3


def setenv(cfg):
    for k in KEYS:
        os.environ.pop(k, None)
    for k, v in cfg.items():
        if v:
            os.environ[k] = str(v)


def conv_case(n, hw, cin, cout, k, stride, res):
    pad = (k - 1) // 2
    x = torch.randn(n, hw, hw, cin, device=DEV).to(torch.bfloat16)
    wt = (torch.randn(cout, k * k * cin, device=DEV) * (k * k * cin) ** -0.5).to(torch.bfloat16)
    b = torch.randn(cout, device=DEV)
    ho = ops.conv_out_size(hw, k, stride, pad, 1)
    r = torch.randn(n, ho, ho, cout, device=DEV).to(torch.bfloat16) if res else None
    out = torch.empty(n, ho, ho, cout, device=DEV, dtype=torch.bfloat16)
    return lambda: ops.conv2d(x, wt, b, cin=cin, cout=cout, kh=k, kw=k, stride=stride, pad=pad, act=1, residual=r, out=out)


def gemm_case(m, n, k, act, res):
    a = torch.randn(m, k, device=DEV).to(torch.bfloat16)
    wt = (torch.randn(n, k, device=DEV) * k ** -0.5).to(torch.bfloat16)
    b = torch.randn(n, device=DEV)
    r = torch.randn(m, n, device=DEV).to(torch.bfloat16) if res else None
    out = torch.empty(m, n, device=DEV, dtype=torch.bfloat16)
    return lambda: ops.gemm(a, wt, b, act=act, residual=r, out=out)


B = 256
RESNET = [("l2.c1 512->128 @28", lambda: conv_case(B, 28, 512, 128, 1, 1, False), 3),
          ("l2.c2 3x3 128 @28", lambda: conv_case(B, 28, 128, 128, 3, 1, False), 3),
          ("l2.c3 128->512 @28 +res", lambda: conv_case(B, 28, 128, 512, 1, 1, True), 4),
          ("l3.down 512->1024 s2", lambda: conv_case(B, 28, 512, 1024, 1, 2, False), 1),
          ("l3.c1 1024->256 @14", lambda: conv_case(B, 14, 1024, 256, 1, 1, False), 5),
          ("l3.c2 3x3 256 @14", lambda: conv_case(B, 14, 256, 256, 3, 1, False), 5),
          ("l3.c3 256->1024 @14 +res", lambda: conv_case(B, 14, 256, 1024, 1, 1, True), 6),
          ("l4.down 1024->2048 s2", lambda: conv_case(B, 14, 1024, 2048, 1, 2, False), 1),
          ("l4.c1 2048->512 @7", lambda: conv_case(B, 7, 2048, 512, 1, 1, False), 2),
          ("l4.c2 3x3 512 @7", lambda: conv_case(B, 7, 512, 512, 3, 1, False), 2),
          ("l4.c3 512->2048 @7 +res", lambda: conv_case(B, 7, 512, 2048, 1, 1, True), 3)]
M = 64 * 197
VIT = [("qkv 768->2304", lambda: gemm_case(M, 2304, 768, 0, False), 12),
       ("proj 768->768 +res", lambda: gemm_case(M, 768, 768, 0, True), 12),
       ("fc1 768->3072 gelu", lambda: gemm_case(M, 3072, 768, 3, False), 12),
       ("fc2 3072->768 +res", lambda: gemm_case(M, 768, 3072, 0, True), 12)]

MB = 128 * 112 * 112
B4 = [("b4 s1 proj 48->24 @112", lambda: gemm_case(MB, 24, 48, 0, False), 1),
      ("b4 s1 proj 24->24 @112 +res", lambda: gemm_case(MB, 24, 24, 0, True), 1),
      ("b4 s2 expand 24->144 silu", lambda: gemm_case(MB, 144, 24, 2, False), 1),
      ("b4 s2 proj 144->32 @56", lambda: gemm_case(MB // 4, 32, 144, 0, False), 1),
      ("b4 s2 expand 32->192 silu", lambda: gemm_case(MB // 4, 192, 32, 2, False), 3),
      ("b4 s2 proj 192->32 @56 +res", lambda: gemm_case(MB // 4, 32, 192, 0, True), 3),
      ("b4 s3 expand 56->336 silu", lambda: gemm_case(MB // 16, 336, 56, 2, False), 3),
      ("b4 s4 expand 112->672 silu", lambda: gemm_case(MB // 64, 672, 112, 2, False), 5),
      ("b4 s5 expand 160->960 silu", lambda: gemm_case(MB // 64, 960, 160, 2, False), 5),
      ("b4 s6 expand 272->1632 silu", lambda: gemm_case(MB // 256, 1632, 272, 2, False), 7)]

which = sys.argv[1] if len(sys.argv) > 1 else "all"
cases = (RESNET if which in ("resnet", "all") else []) + (VIT if which in ("vit", "all") else []) + \
    (B4 if which in ("b4", "all") else [])
gain = 0.0
for name, make, count in cases:
    fn = make()
    setenv({})
    base = timeit(fn)
    best = (base, "default")
    for nopair, bn, sub in itertools.product((0, 1), (0, 64, 128, 192, 256), (0, 1, 2)):
        if not (nopair or bn or sub):
            continue
        setenv({"EQXV_NO_PAIR": nopair, "EQXV_BLOCK_N": bn, "EQXV_EPI_SUB": sub})
        try:
            t = timeit(fn, iters=8, warm=1)
        except Exception as ex:  # noqa: BLE001
            continue
        if t < best[0]:
            best = (t, f"nopair={nopair} bn={bn or 'auto'} sub={sub or 'auto'}")
    setenv({})
    gain += (base - best[0]) * count
    print(f"{name:28s} x{count}: default {base:7.1f} us   best {best[0]:7.1f} us  ({best[1]})", flush=True)
print(f"sum of (default - best) x count: {gain:.0f} us")
