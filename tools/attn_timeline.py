"""Phase timeline of the tcgen05 attention kernels (CTA 0, first 12 tiles), from clock64 stamps (EQXV_ATTN_PP selects the kernel).
MMA thread: 0 loop top, 1 P ready, 2 PV issued+committed, 3 after prefetch, 4 QK(g+2) issued.
Softmax warp 1: 8 tile start, 9 S ready, 10 max done, 11 exp done, 12 sums exchanged, 13 O(g-1) drained, 14 P written."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eqxvision_b200 import _lib, ops
_lib.init(0)
imgs, tokens, heads = 64, 197, 12
qkv = torch.randn(imgs * tokens, 3 * heads * 64, device="cuda").to(torch.bfloat16)
ts = torch.zeros(16 * 16, dtype=torch.int64, device="cuda")
for _ in range(2):
    ops.attention(qkv, imgs, tokens, heads, 64, 0.125)
torch.cuda.synchronize()
_lib.call("eqxv_debug_attention_timeline", ts.data_ptr())
ops.attention(qkv, imgs, tokens, heads, 64, 0.125)
torch.cuda.synchronize()
_lib.call("eqxv_debug_attention_timeline", None)
t = ts.cpu().reshape(16, 16)
t0 = int(t[t > 0].min())
print("tile |" + "".join(f"{e:>7d}" for e in list(range(0, 5)) + list(range(8, 15))))
for g in range(12):
    row = [int(t[g, e]) - t0 if t[g, e] > 0 else -1 for e in list(range(0, 5)) + list(range(8, 15))]
    print(f"{g:4d} |" + "".join(f"{v:7d}" for v in row))
