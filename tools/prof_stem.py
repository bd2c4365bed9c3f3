"""ncu helper: the ResNet stem (pack + 7x7 s2 conv) and the max-pool at B=256, a few launches."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eqxvision_b200 import _lib, _pack, ops
_lib.init(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
x = torch.rand(n, 3, 224, 224, device="cuda")
wt = torch.randn(64, 3, 7, 7) * 0.1
wp = _pack.pack_stem_weight(wt).cuda()
b = torch.randn(64, device="cuda")
for _ in range(3):
    xp = ops.pack_stem_input(x)
    y = ops.conv_stem(xp, wp, b, n=n, h=224, w=224, cout=64, kh=7, kw=7, stride=2, pad=3, act=1)
    z = ops.maxpool2d(y, 3, 2, 1)
torch.cuda.synchronize()
print("ok", y.shape, z.shape)
