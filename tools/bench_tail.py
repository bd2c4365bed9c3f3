"""(1) last-K-chunk tail shift on / off for EfficientNet-B4's projections; (2) stem + max-pool fused vs separate.
usage: python tools/bench_tail.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eqxvision_b200 import _lib, _pack, ops  # noqa: E402

_lib.init(0)


def timed(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


for m, n, k, hw in [(128 * 56 * 56, 32, 144, 3136), (128 * 28 * 28, 56, 336, 784), (128 * 14 * 14, 112, 672, 196),
                    (128 * 14 * 14, 160, 960, 196), (128 * 7 * 7, 272, 1632, 49), (128 * 56 * 56, 32, 192, 3136)]:
    a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
    gate = torch.rand(m // hw, k, device="cuda").to(torch.bfloat16)
    w = torch.randn(n, k, 1, 1) * k ** -0.5
    b = torch.randn(n, device="cuda")
    out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    wp = _pack.pack_conv_weight(w, k).cuda()
    line = f"{m:8d} x {n:4d} x {k:4d}: gemm {timed(lambda: ops.gemm(a, wp, b, out=out)):7.1f}  gated {timed(lambda: ops.gemm_gated(a, gate, wp, b, rows_per_image=hw, out=out)):7.1f}"
    if _pack.tail_shift_applies(k):
        ws = _pack.pack_conv_weight(w, k, tail_shift=True).cuda()
        a4 = a.view(1, 1, m, k)
        o4 = out.view(1, 1, m, n)
        line += f"  | tail shift: gemm {timed(lambda: ops.conv2d(a4, ws, b, cin=k, cout=n, kh=1, kw=1, out=o4, k_tail_shift=True)):7.1f}"
        line += f"  gated {timed(lambda: ops.gemm_gated(a, gate, ws, b, rows_per_image=hw, out=out, k_tail_shift=True)):7.1f}"
    print(line + f"   (A read alone at 6.4 TB/s: {m * k * 2 / 6.4e6:6.1f} us)", flush=True)

n, h = 256, 224
x = torch.rand(n, 3, h, h, device="cuda")
wt = _pack.pack_stem_weight(torch.randn(64, 3, 7, 7) * 0.1).cuda()
bias = torch.randn(64, device="cuda")
xpad = ops.pack_stem_input(x)
y = torch.empty(n, 112, 112, 64, device="cuda", dtype=torch.bfloat16)
yp = torch.empty(n, 56, 56, 64, device="cuda", dtype=torch.bfloat16)
print(f"stem {timed(lambda: ops.conv_stem(xpad, wt, bias, n=n, h=h, w=h, cout=64, out=y)):7.1f} us  "
      f"maxpool {timed(lambda: ops.maxpool2d(y, k=3, stride=2, pad=1, out=yp)):7.1f} us  "
      f"fused {timed(lambda: ops.conv_stem_maxpool(xpad, wt, bias, n=n, h=h, w=h, cout=64, out=yp)):7.1f} us")
