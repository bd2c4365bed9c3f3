"""Per-shape timing of the depthwise kernels on EfficientNet-B4's layers (batch 128): register-strip kernel
vs shared-memory stencil kernel, CUDA events, achieved GB/s against algorithmic bytes (in + out, bf16)."""
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
tot = {False: 0.0, True: 0.0}
print(f"{'c':>5s} {'hw':>4s} k s  x{'n':<3s} {'strip us':>9s} {'GB/s':>6s} {'stencil us':>10s} {'GB/s':>6s}")
for c, hw, k, s, cnt in SHAPES:
    x = torch.randn(N, hw, hw, c, device="cuda").to(torch.bfloat16)
    w = _pack.pack_depthwise_weight(torch.randn(c, 1, k, k), c).cuda()
    b = torch.randn(c, device="cuda")
    pad = (k - 1) // 2
    ho = (hw + 2 * pad - k) // s + 1
    out = torch.empty(N, ho, ho, c, device="cuda", dtype=torch.bfloat16)
    nbytes = (x.numel() + out.numel()) * 2
    res = {}
    for tile in (False, True):
        for _ in range(3):
            ops.dwconv(x, w, b, k=k, stride=s, pad=pad, act=2, out=out, tile=tile)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.dwconv(x, w, b, k=k, stride=s, pad=pad, act=2, out=out, tile=tile)
        e1.record()
        torch.cuda.synchronize()
        res[tile] = e0.elapsed_time(e1) / 10 * 1e3
        tot[tile] += res[tile] * cnt
    print(f"{c:5d} {hw:4d} {k} {s}  x{cnt:<3d} {res[False]:9.1f} {nbytes / res[False] / 1e3:6.0f} {res[True]:10.1f} {nbytes / res[True] / 1e3:6.0f}")
print(f"sum over the 32 blocks: strip {tot[False]:.0f} us, stencil {tot[True]:.0f} us")
