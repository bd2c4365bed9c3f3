"""one depthwise layer, a few launches (ncu target): python tools/dw_one.py C HW K S [pool]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eqxvision_b200 import _lib, _pack, ops  # noqa: E402

_lib.init(0)
c, hw, k, s = (int(v) for v in sys.argv[1:5])
pool = len(sys.argv) > 5
N = 128
x = torch.randn(N, hw, hw, c, device="cuda").to(torch.bfloat16)
w = _pack.pack_depthwise_weight(torch.randn(c, 1, k, k), c).cuda()
b = torch.randn(c, device="cuda")
pad = (k - 1) // 2
ho = (hw + 2 * pad - k) // s + 1
out = torch.empty(N, ho, ho, c, device="cuda", dtype=torch.bfloat16)
pooled = torch.empty(N, c, device="cuda", dtype=torch.bfloat16)
ws = torch.zeros(max(ops.dwconv_pool_workspace_bytes(N, hw, hw, c, k, s, pad), 16), dtype=torch.uint8, device="cuda")
for _ in range(6):
    if pool:
        ops.dwconv_pool(x, w, b, k=k, stride=s, pad=pad, act=2, out=out, pooled=pooled, workspace=ws)
    else:
        ops.dwconv(x, w, b, k=k, stride=s, pad=pad, act=2, out=out)
torch.cuda.synchronize()
