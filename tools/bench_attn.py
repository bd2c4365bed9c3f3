"""ViT-B/16 attention launch (64 images x 12 heads x 197 tokens): lock-step kernel vs ping-pong kernel."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eqxvision_b200 import _lib, ops  # noqa: E402

_lib.init(0)
imgs, tokens, heads = 64, 197, 12
qkv = torch.randn(imgs * tokens, 3 * heads * 64, device="cuda").to(torch.bfloat16)
ref = None
for pp in ("0", "1"):
    os.environ["EQXV_ATTN_PP"] = pp
    for _ in range(3):
        out = ops.attention(qkv, imgs, tokens, heads, 64, 0.125)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        out = ops.attention(qkv, imgs, tokens, heads, 64, 0.125)
    e1.record()
    torch.cuda.synchronize()
    if ref is None:
        ref = out.float()
    print(f"EQXV_ATTN_PP={pp}: {e0.elapsed_time(e1) / 20 * 1e3:7.1f} us per launch, "
          f"rel-L2 vs lock-step {((out.float() - ref).norm() / ref.norm()).item():.2e}")
