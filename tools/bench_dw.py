"""Per-shape timing of the depthwise kernels on EfficientNet-B4's layers (batch 128): the packed-FMA per-image kernel
(csrc/dwconv_img.cu, with and without the fused SE squeeze) against the register-strip kernel (EQXV_DWIMG=0 in a second
process), CUDA events, achieved GB/s against algorithmic bytes (in + out, bf16).

  python tools/bench_dw.py            # per-image kernel: plain, fused squeeze
  EQXV_DWIMG=0 python tools/bench_dw.py   # the round-1 strip kernel through the same entry"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eqxvision_b200 import _lib, _pack, ops  # noqa: E402

_lib.init(0)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
# (channels, input hw, k, stride) of the 32 MBConv blocks of efficientnet_b4 @224 (distinct shapes, count)
SHAPES = [(48, 112, 3, 1, 2), (144, 112, 3, 2, 1), (192, 56, 3, 1, 3), (192, 56, 5, 2, 1), (336, 28, 5, 1, 3),
          (336, 28, 3, 2, 1), (672, 14, 3, 1, 5), (672, 14, 5, 1, 1), (960, 14, 5, 1, 5), (960, 14, 5, 2, 1),
          (1632, 7, 5, 1, 7), (1632, 7, 3, 1, 1), (2688, 7, 3, 1, 1)]
legacy = os.environ.get("EQXV_DWIMG") == "0"
tot = [0.0, 0.0]
print(("strip kernel (EQXV_DWIMG=0)" if legacy else "per-image packed-FMA kernel") + f", batch {N}")
print(f"{'c':>5s} {'hw':>4s} k s  x{'n':<3s} {'plain us':>9s} {'GB/s':>6s} {'+squeeze us':>11s} {'GB/s':>6s}")


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 10 * 1e3


for c, hw, k, s, cnt in SHAPES:
    x = torch.randn(N, hw, hw, c, device="cuda").to(torch.bfloat16)
    w = _pack.pack_depthwise_weight(torch.randn(c, 1, k, k), c).cuda()
    b = torch.randn(c, device="cuda")
    pad = (k - 1) // 2
    ho = (hw + 2 * pad - k) // s + 1
    out = torch.empty(N, ho, ho, c, device="cuda", dtype=torch.bfloat16)
    pooled = torch.empty(N, c, device="cuda", dtype=torch.bfloat16)
    need = ops.dwconv_pool_workspace_bytes(N, hw, hw, c, k, s, pad)
    ws = torch.zeros(max(need, 16), dtype=torch.uint8, device="cuda")
    nbytes = (x.numel() + out.numel()) * 2
    t0 = timed(lambda: ops.dwconv(x, w, b, k=k, stride=s, pad=pad, act=2, out=out))
    t1 = timed(lambda: ops.dwconv_pool(x, w, b, k=k, stride=s, pad=pad, act=2, out=out, pooled=pooled, workspace=ws))
    tot[0] += t0 * cnt
    tot[1] += t1 * cnt
    print(f"{c:5d} {hw:4d} {k} {s}  x{cnt:<3d} {t0:9.1f} {nbytes / t0 / 1e3:6.0f} {t1:11.1f} {nbytes / t1 / 1e3:6.0f}")
print(f"sum over the 32 blocks: plain {tot[0]:.0f} us, with fused squeeze {tot[1]:.0f} us")
