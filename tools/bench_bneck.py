"""ResNet-50 layer1 at batch B: the fused bottleneck kernel against the layer-by-layer launches it replaces.
usage: python tools/bench_bneck.py [B] [iters]   (timing: CUDA events on the launching stream, L2 flushed by the working set)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eqxvision_b200 import _lib, ops  # noqa: E402

_lib.init(0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = "cuda"
bf = torch.bfloat16
H = W = 56


def rnd(*s, scale=1.0):
    return (torch.randn(*s, device=dev) * scale).to(bf)


t1 = rnd(B, H, W, 64)
x0 = rnd(B, H, W, 64)
res = rnd(B, H, W, 256)
w2 = rnd(64, 576, scale=576 ** -0.5)
w3 = rnd(256, 64, scale=0.125)
w3cat = rnd(256, 128, scale=0.09)
w1n = rnd(64, 256, scale=1 / 16)
b2, b3, b1n = (torch.randn(c, device=dev) * 0.1 for c in (64, 256, 64))
y = torch.empty(B, H, W, 256, device=dev, dtype=bf)
nx = torch.empty(B, H, W, 64, device=dev, dtype=bf)
t2 = torch.empty(B, H, W, 64, device=dev, dtype=bf)
dn = torch.empty(B, H, W, 256, device=dev, dtype=bf)


def timed(name, fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:54s} {e0.elapsed_time(e1) / iters * 1e3:8.1f} us", flush=True)


def layer_identity():
    ops.conv2d(t1, w2, b2, cin=64, cout=64, kh=3, kw=3, pad=1, act=1, out=t2)
    ops.conv2d(t2, w3, b3, cin=64, cout=256, kh=1, kw=1, act=1, residual=res, out=y)
    ops.conv2d(y, w1n, b1n, cin=256, cout=64, kh=1, kw=1, act=1, out=nx)


def layer_down():
    ops.conv2d(x0, w3, b3, cin=64, cout=256, kh=1, kw=1, act=0, out=dn)
    ops.conv2d(t1, w2, b2, cin=64, cout=64, kh=3, kw=3, pad=1, act=1, out=t2)
    ops.conv2d(t2, w3, b3, cin=64, cout=256, kh=1, kw=1, act=1, residual=dn, out=y)
    ops.conv2d(y, w1n, b1n, cin=256, cout=64, kh=1, kw=1, act=1, out=nx)


timed("layer by layer: c2 + c3(+res) + next c1", layer_identity)
timed("fused identity + next", lambda: ops.bottleneck64(t1, w2, b2, w3, b3, residual=res, out=y, w1n=w1n, b1n=b1n, next_out=nx))
timed("layer by layer: c2 + c3(+res)", lambda: (ops.conv2d(t1, w2, b2, cin=64, cout=64, kh=3, kw=3, pad=1, act=1, out=t2),
                                                 ops.conv2d(t2, w3, b3, cin=64, cout=256, kh=1, kw=1, act=1, residual=res, out=y)))
timed("fused identity", lambda: ops.bottleneck64(t1, w2, b2, w3, b3, residual=res, out=y))
timed("layer by layer: down + c2 + c3(+res) + next c1", layer_down)
timed("fused downsample + next", lambda: ops.bottleneck64(t1, w2, b2, w3cat, b3, x0=x0, out=y, w1n=w1n, b1n=b1n, next_out=nx))
timed("fused downsample", lambda: ops.bottleneck64(t1, w2, b2, w3cat, b3, x0=x0, out=y))
w1n128 = rnd(128, 256, scale=1 / 16)
b1n128 = torch.randn(128, device=dev) * 0.1
nx128 = torch.empty(B, H, W, 128, device=dev, dtype=bf)
timed("layer by layer: c2 + c3(+res) + next c1 (128)", lambda: (ops.conv2d(t1, w2, b2, cin=64, cout=64, kh=3, kw=3, pad=1, act=1, out=t2),
                                                                 ops.conv2d(t2, w3, b3, cin=64, cout=256, kh=1, kw=1, act=1, residual=res, out=y),
                                                                 ops.conv2d(y, w1n128, b1n128, cin=256, cout=128, kh=1, kw=1, act=1, out=nx128)))
timed("fused identity + next (128)", lambda: ops.bottleneck64(t1, w2, b2, w3, b3, residual=res, out=y, w1n=w1n128, b1n=b1n128, next_out=nx128))
