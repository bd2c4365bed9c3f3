"""one GEMM shape, a few launches (ncu target): python tools/gemm_one.py M N K ACT RES"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eqxvision_b200 import _lib, ops  # noqa: E402

_lib.init(0)
m, n, k, act, res = (int(v) for v in sys.argv[1:6])
a = torch.randn(m, k, device="cuda").to(torch.bfloat16)
wt = (torch.randn(n, k, device="cuda") * k ** -0.5).to(torch.bfloat16)
b = torch.randn(n, device="cuda")
r = torch.randn(m, n, device="cuda").to(torch.bfloat16) if res else None
out = torch.empty(m, n, device="cuda", dtype=torch.bfloat16)
for _ in range(5):
    ops.gemm(a, wt, b, act=act, residual=r, out=out)
torch.cuda.synchronize()
