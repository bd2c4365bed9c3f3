"""Phase timeline of the first-layer kernel (CTA 0, first 20 tiles), from clock64 stamps (ResNet stem, batch 256).
Issuer: 0 accumulator free, 1 operand tile landed, 2 MMAs issued + committed.
Gather warp 0: 3 stage free, 4 copies issued, 5 arrive for tile (column index = tile whose copies have LANDED).
Epilogue warp of the tile's group: 6 accumulator full seen, 7 store issued.
usage: python tools/stem_timeline.py [c4]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eqxvision_b200 import _lib, _pack, ops
_lib.init(0)
c4 = len(sys.argv) > 1 and sys.argv[1] == "c4"
n, h = 256, 224
x = torch.rand(n, 3, h, h, device="cuda")
w = torch.randn(64, 3, 7, 7) * 0.1
wt = (_pack.pack_stem_weight_c4(w) if c4 else _pack.pack_stem_weight(w)).cuda()
bias = torch.randn(64, device="cuda")
xpad = ops.pack_stem_input_c4(x) if c4 else ops.pack_stem_input(x)
y = torch.empty(n, 112, 112, 64, device="cuda", dtype=torch.bfloat16)
run = lambda: ops.conv_stem(xpad, wt, bias, n=n, h=h, w=h, cout=64, out=y, c4=c4)
ts = torch.zeros(24 * 16, dtype=torch.int64, device="cuda")
for _ in range(2):
    run()
torch.cuda.synchronize()
_lib.call("eqxv_debug_stem_timeline", ts.data_ptr())
run()
torch.cuda.synchronize()
_lib.call("eqxv_debug_stem_timeline", None)
t = ts.cpu().reshape(24, 16)
t0 = int(t[t > 0].min())
print(f"c4={c4}\ntile |" + "".join(f"{e:>7d}" for e in range(8)))
for g in range(20):
    print(f"{g:4d} |" + "".join(f"{(int(t[g, e]) - t0 if t[g, e] > 0 else -1):7d}" for e in range(8)))
