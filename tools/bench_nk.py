import os, sys, torch
sys.path.insert(0, "/root/repo")
from eqxvision_b200 import _lib, ops
_lib.init(0)
m = 128 * 112 * 112
for n, k in [(144, 24), (192, 24), (144, 32), (128, 24), (128, 32), (128, 64), (192, 64), (160, 32), (144, 64), (256, 64), (64, 24), (64, 64)]:
    a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
    wt = (torch.randn(n, k, device="cuda") * k ** -0.5).to(torch.bfloat16)
    b = torch.randn(n, device="cuda")
    out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
    for _ in range(3): ops.gemm(a, wt, b, act=2, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.gemm(a, wt, b, act=2, out=out)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    print(f"N={n:4d} K={k:3d}: {us:7.1f} us  {(m*n*2+m*k*2)/us/1e3:6.0f} GB/s  {us*1.965e3/(m/128/148):7.0f} clk/tile", flush=True)
