"""ncu helper: runs a few igemm layer shapes a handful of times (select with substrings)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eqxvision_b200 import _lib, ops
DEV = "cuda:0"
CASES = {
    "l1_1x1_64_64": (256, 56, 64, 64, 1, 1, False),
    "l1_1x1_64_256_res": (256, 56, 64, 256, 1, 1, True),
    "l1_3x3_64": (256, 56, 64, 64, 3, 1, False),
    "l3_3x3_256": (256, 14, 256, 256, 3, 1, False),
    "l3_1x1_256_1024_res": (256, 14, 256, 1024, 1, 1, True),
    "l2_1x1_512_128": (256, 28, 512, 128, 1, 1, False),
}
_lib.init(0)
sel = sys.argv[1:] or list(CASES)
for name in sel:
    n, h, cin, cout, k, s, res = CASES[name]
    pad = (k - 1) // 2
    x = torch.randn(n, h, h, cin, device=DEV).to(torch.bfloat16)
    wt = (torch.randn(cout, k * k * cin, device=DEV) * 0.05).to(torch.bfloat16)
    b = torch.randn(cout, device=DEV)
    ho = ops.conv_out_size(h, k, s, pad, 1)
    r = torch.randn(n, ho, ho, cout, device=DEV).to(torch.bfloat16) if res else None
    out = torch.empty(n, ho, ho, cout, device=DEV, dtype=torch.bfloat16)
    for _ in range(3):
        ops.conv2d(x, wt, b, cin=cin, cout=cout, kh=k, kw=k, stride=s, pad=pad, act=1, residual=r, out=out)
    torch.cuda.synchronize()
    print("ran", name)
