"""ncu helper: ViT-B/16 attention (64 images x 12 heads x 197 tokens) and LayerNorm launches."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eqxvision_b200 import _lib, ops
_lib.init(0)
imgs, tokens, heads = 64, 197, 12
qkv = torch.randn(imgs * tokens, 3 * heads * 64, device="cuda").to(torch.bfloat16)
x = torch.randn(imgs * tokens, 768, device="cuda").to(torch.bfloat16)
g = torch.ones(768, device="cuda"); b = torch.zeros(768, device="cuda")
for _ in range(3):
    out = ops.attention(qkv, imgs, tokens, heads, 64, 0.125)
    y = ops.layernorm(x, g, b, 1e-5)
torch.cuda.synchronize()
print("ok", out.shape)
