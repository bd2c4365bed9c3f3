"""Phase timeline of the fused bottleneck kernel (CTA 0, first 14 tiles), from clock64 stamps.
Issuer thread: 0 a3full seen, 1 c3 issued, 2 c2(next) issued, 3 yfull seen, 4 c1' issued.
Epilogue warp 2: 5 d3full seen, 6 residual landed, 7 e3 done, 8 e2(next) done, 9 d1full seen, 10 e4 done.
usage: python tools/bneck_timeline.py [down] [next]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eqxvision_b200 import _lib, ops
_lib.init(0)
down, nxt = int(sys.argv[1]) if len(sys.argv) > 1 else 0, int(sys.argv[2]) if len(sys.argv) > 2 else 1
B, H, W, bf = 256, 56, 56, torch.bfloat16
rnd = lambda *s, scale=1.0: (torch.randn(*s, device="cuda") * scale).to(bf)
t1, x0, res = rnd(B, H, W, 64), rnd(B, H, W, 64), rnd(B, H, W, 256)
w2, w3, w1n = rnd(64, 576, scale=0.04), rnd(256, 128 if down else 64, scale=0.1), rnd(64, 256, scale=0.06)
b2, b3, b1n = (torch.randn(c, device="cuda") * 0.1 for c in (64, 256, 64))
y, nx = torch.empty(B, H, W, 256, device="cuda", dtype=bf), torch.empty(B, H, W, 64, device="cuda", dtype=bf)
kw = dict(residual=None if down else res, x0=x0 if down else None, out=y)
if nxt:
    kw.update(w1n=w1n, b1n=b1n, next_out=nx)
ts = torch.zeros(16 * 16, dtype=torch.int64, device="cuda")
for _ in range(2):
    ops.bottleneck64(t1, w2, b2, w3, b3, **kw)
torch.cuda.synchronize()
_lib.call("eqxv_debug_bottleneck_timeline", ts.data_ptr())
ops.bottleneck64(t1, w2, b2, w3, b3, **kw)
torch.cuda.synchronize()
_lib.call("eqxv_debug_bottleneck_timeline", None)
t = ts.cpu().reshape(16, 16)
t0 = int(t[t > 0].min())
ev = list(range(11))
print(f"down={down} next={nxt}\ntile |" + "".join(f"{e:>7d}" for e in ev))
for g in range(14):
    print(f"{g:4d} |" + "".join(f"{(int(t[g, e]) - t0 if t[g, e] > 0 else -1):7d}" for e in ev))
