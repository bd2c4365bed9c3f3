import os, sys, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eqxvision_b200 import _lib, ops
_lib.init(0)
dev = torch.device("cuda", 0)
def rb(*shape, scale=1.0, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16).to(dev)
for m, d, n, act in [(12608, 768, 2048, 0), (300, 192, 576, 0), (1000, 96, 288, 0), (256, 768, 256, 0)]:
    g = torch.Generator().manual_seed(5)
    a0 = rb(m, d, seed=1); w0 = rb(d, d, scale=d ** -0.5, seed=2); b0 = torch.randn(d, generator=g).to(dev)
    res = rb(m, d, seed=3) * 3 + 1.5
    stats = torch.zeros(m, (d + 63) // 64, 2, device=dev)
    x = ops.gemm_rowstats(a0, w0, b0, residual=res, stats=stats)
    torch.cuda.synchronize()
    xs = x.float()
    ref_x = a0.float() @ w0.float().t() + b0 + res.float()
    print(m, d, n, "x rel", ((xs - ref_x).norm() / ref_x.norm()).item())
    s0 = torch.stack([xs[:, c:c + 64].sum(1) for c in range(0, d, 64)], 1)
    s1 = torch.stack([(xs[:, c:c + 64] ** 2).sum(1) for c in range(0, d, 64)], 1)
    print("  stats err", (stats[..., 0] - s0).abs().max().item(), (stats[..., 1] - s1).abs().max().item(), "max", s1.max().item())
    gamma = (1 + 0.2 * torch.randn(d, generator=g)).to(dev); beta = (0.3 * torch.randn(d, generator=g)).to(dev)
    w1 = (torch.randn(n, d, generator=g) * d ** -0.5).to(dev); b1 = torch.randn(n, generator=g).to(dev)
    wf = (w1 * gamma[None, :]).to(torch.bfloat16)
    bias = (b1.double() + w1.double() @ beta.double()).float()
    y = ops.gemm_ln(x, wf, bias, wf.float().sum(1), stats, 1e-5, act=act)
    torch.cuda.synchronize()
    ref = F.layer_norm(xs, (d,), gamma, beta, 1e-5) @ w1.t() + b1
    e = (y.float() - ref)
    print("  y rel", (e.norm() / ref.norm()).item(), "rows err", e.norm(dim=1)[:4].tolist(), "by col block", [e[:, c:c+64].norm().item() for c in range(0, min(n, 512), 64)])
    mean = xs.mean(1); var = xs.var(1, unbiased=False); rstd = torch.rsqrt(var + 1e-5)
    emu = rstd[:, None] * (xs @ wf.float().t() - mean[:, None] * wf.float().sum(1)[None]) + bias
    print("  vs host formula", ((y.float() - emu).norm() / emu.norm()).item())
