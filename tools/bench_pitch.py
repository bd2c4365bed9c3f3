"""narrow-K 1x1 expansion: effect of the A operand's row pitch (TMA fetches rows of K*2 bytes) and of the output pitch.
usage: python tools/bench_pitch.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eqxvision_b200 import _lib, ops  # noqa: E402

_lib.init(0)


def run(m, n, k, act, apitch, opitch):
    a = torch.randn(m, apitch, device="cuda").to(torch.bfloat16)[:, :k]
    wt = (torch.randn(n, k, device="cuda") * k ** -0.5).to(torch.bfloat16)
    b = torch.randn(n, device="cuda")
    out = torch.empty(m, opitch, device="cuda", dtype=torch.bfloat16)[:, :n]
    for _ in range(3):
        ops.gemm(a, wt, b, act=act, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.gemm(a, wt, b, act=act, out=out)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 100


for m, n, k, act in [(128 * 112 * 112, 144, 24, 2), (128 * 112 * 112, 144, 24, 0), (128 * 56 * 56, 192, 32, 2), (128 * 28 * 28, 336, 56, 2),
                     (128 * 112 * 112, 24, 48, 0), (128 * 112 * 112, 24, 144, 0)]:
    line = f"{m:8d} x {n:4d} x {k:4d} act {act}:"
    for ap in sorted({k, 32 * ((k + 31) // 32), 64 * ((k + 63) // 64)}):
        for op in sorted({n, 64 * ((n + 63) // 64)}):
            line += f"  a{ap}/o{op} {run(m, n, k, act, ap, op):6.1f}"
    print(line, flush=True)
