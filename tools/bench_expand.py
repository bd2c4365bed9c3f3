"""EfficientNet-B4's 1x1 expansions at batch 128 (one K block, SiLU): eight- vs sixteen-warp epilogue (EQXV_EPI_SUB).
usage: python tools/bench_expand.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eqxvision_b200 import _lib, ops  # noqa: E402

_lib.init(0)
CASES = [(128 * 112 * 112, 144, 24, 2), (128 * 56 * 56, 192, 32, 2), (128 * 28 * 28, 336, 56, 2), (128 * 14 * 14, 672, 112, 2),
         (128 * 14 * 14, 960, 160, 2), (128 * 7 * 7, 1632, 272, 2), (256 * 56 * 56, 64, 64, 1), (256 * 56 * 56, 256, 64, 0)]
for m, n, k, act in CASES:
    a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
    wt = (torch.randn(n, k, device="cuda") * k ** -0.5).to(torch.bfloat16)
    b = torch.randn(n, device="cuda")
    out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    line = f"{m:8d} x {n:4d} x {k:4d} act {act}:"
    for sub in ("2", "4", ""):
        os.environ["EQXV_EPI_SUB"] = sub
        if not sub:
            os.environ.pop("EQXV_EPI_SUB")
        for _ in range(3):
            ops.gemm(a, wt, b, act=act, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.gemm(a, wt, b, act=act, out=out)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        line += f"  sub={sub or 'auto'} {us:7.1f} us ({(m * n * 2 + m * k * 2) / us / 1e3:5.0f} GB/s)"
    print(line, flush=True)
