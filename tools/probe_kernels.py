"""GPU bring-up probe: runs each kernel family against a torch (GPU, fp32) reference and prints
detailed error diagnostics. Usage (on a B200 box):  python tools/probe_kernels.py <group>
Each group runs in its own process so that a trapped kernel does not poison the others.
"""
from __future__ import annotations

import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eqxvision_b200 import _lib, ops  # noqa: E402

DEV = "cuda:0"
ACTS = {0: lambda v: v, 1: F.relu, 2: F.silu, 3: lambda v: F.gelu(v, approximate="tanh"),
        4: F.hardswish, 5: torch.sigmoid, 6: F.hardsigmoid, 7: F.relu6}


def report(name, got, ref, tol=2e-2):
    got = got.float()
    ref = ref.float()
    err = (got - ref).abs()
    denom = ref.abs().max().item() + 1e-12
    rel_l2 = ((got - ref).norm() / (ref.norm() + 1e-12)).item()
    mx = err.max().item()
    bad = (err > tol * denom).float().mean().item()
    nan = torch.isnan(got).sum().item()
    ok = (rel_l2 < tol) and nan == 0
    print(f"[{'OK ' if ok else 'BAD'}] {name}: rel_l2={rel_l2:.3e} max_abs={mx:.3e} ref_max={denom:.3e} "
          f"bad_frac={bad:.4f} nan={nan} shape={tuple(got.shape)}", flush=True)
    if not ok:
        idx = torch.nonzero(err > tol * denom)[:8]
        for i in idx:
            t = tuple(i.tolist())
            print(f"      at {t}: got {got[t].item():.5f} ref {ref[t].item():.5f}")
        # coarse error map over the last two dims collapsed
        e2 = err.reshape(-1, err.shape[-1])
        rows = e2.shape[0]
        rb = min(rows, 16)
        cb = min(e2.shape[1], 16)
        em = e2[: rows // rb * rb].reshape(rb, -1, e2.shape[1]).amax(1)
        em = em[:, : e2.shape[1] // cb * cb].reshape(rb, cb, -1).amax(2)
        print("      error map (row-blocks x col-blocks, max abs):")
        for r in range(rb):
            print("      " + " ".join(f"{v:8.2e}" for v in em[r].tolist()))
    return ok


def rand_bf16(*shape, scale=1.0, seed=0):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16).to(DEV)


def conv_case(name, n, h, w, cin, cout, k, stride, pad, dil, act=0, res=False, res_after_act=False,
              seed=0):
    x = rand_bf16(n, h, w, cin, seed=seed)
    wt = rand_bf16(cout, k, k, cin, scale=(k * k * cin) ** -0.5, seed=seed + 1)
    bias = torch.randn(cout, generator=torch.Generator().manual_seed(seed + 2)).to(DEV)
    ho = ops.conv_out_size(h, k, stride, pad, dil)
    wo = ops.conv_out_size(w, k, stride, pad, dil)
    r = rand_bf16(n, ho, wo, cout, seed=seed + 3) if res else None
    y = ops.conv2d(x, wt.reshape(cout, -1), bias, cin=cin, cout=cout, kh=k, kw=k, stride=stride, pad=pad,
                   dil=dil, act=act, residual=r, res_after_act=res_after_act)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float().permute(0, 3, 1, 2), bias, stride=stride,
                   padding=pad, dilation=dil)
    if res and not res_after_act:
        ref = ref + r.float().permute(0, 3, 1, 2)
    ref = ACTS[act](ref)
    if res and res_after_act:
        ref = ref + r.float().permute(0, 3, 1, 2)
    return report(name, y.float(), ref.permute(0, 2, 3, 1))


def gemm_case(name, m, n, k, act=0, res=False, out_f32=False, seed=0):
    a = rand_bf16(m, k, seed=seed)
    wt = rand_bf16(n, k, scale=k ** -0.5, seed=seed + 1)
    bias = torch.randn(n, generator=torch.Generator().manual_seed(seed + 2)).to(DEV)
    r = rand_bf16(m, n, seed=seed + 3) if res else None
    y = ops.gemm(a, wt, bias, act=act, residual=r, out_f32=out_f32)
    torch.cuda.synchronize()
    ref = a.float() @ wt.float().t() + bias
    if res:
        ref = ref + r.float()
    ref = ACTS[act](ref)
    return report(name, y, ref, tol=2e-2 if not out_f32 else 1e-2)


def group_gemm():
    ok = True
    ok &= gemm_case("gemm 128x64x64", 128, 64, 64)
    ok &= gemm_case("gemm 256x64x64", 256, 64, 64)
    ok &= gemm_case("gemm 128x64x256 (4 k-steps)", 128, 64, 256)
    ok &= gemm_case("gemm 128x256x64", 128, 256, 64)
    ok &= gemm_case("gemm 1000x128x192 partial M", 1000, 128, 192)
    ok &= gemm_case("gemm 4096x256x1024 many tiles", 4096, 256, 1024)
    ok &= gemm_case("gemm 40000x64x64 persistent wrap", 40000, 64, 64)
    ok &= gemm_case("gemm 12608x2304x768 qkv", 12608, 2304, 768)
    ok &= gemm_case("gemm 12608x768x3072 +res", 12608, 768, 3072, res=True)
    ok &= gemm_case("gemm 12608x3072x768 gelu", 12608, 3072, 768, act=3)
    ok &= gemm_case("gemm 256x1000x2048 f32 out", 256, 1000, 2048, out_f32=True)
    ok &= gemm_case("gemm 300x272x1632 odd N", 300, 272, 1632)
    ok &= gemm_case("gemm 512x24x144 tiny N", 512, 24, 144, act=2)
    ok &= gemm_case("gemm 512x64x24 tiny K", 512, 64, 24, act=1)
    return ok


def group_conv():
    ok = True
    ok &= conv_case("conv1x1 56x56 64->256 +res relu", 4, 56, 56, 64, 256, 1, 1, 0, 1, act=1, res=True)
    ok &= conv_case("conv3x3 s1 56x56 64->64 relu", 2, 56, 56, 64, 64, 3, 1, 1, 1, act=1)
    ok &= conv_case("conv3x3 s1 28x28 128->128", 8, 28, 28, 128, 128, 3, 1, 1, 1)
    ok &= conv_case("conv3x3 s1 14x14 256->256", 32, 14, 14, 256, 256, 3, 1, 1, 1)
    ok &= conv_case("conv3x3 s1 7x7 512->512 n=5", 5, 7, 7, 512, 512, 3, 1, 1, 1)
    ok &= conv_case("conv3x3 s1 17x13 odd 72->40", 3, 17, 13, 72, 40, 3, 1, 1, 1, act=2)
    ok &= conv_case("conv3x3 d2 p2 32x32 64->64", 2, 32, 32, 64, 64, 3, 1, 2, 2)
    ok &= conv_case("conv3x3 d12 p12 16x16 (tap skipping)", 2, 16, 16, 64, 64, 3, 1, 12, 12)
    ok &= conv_case("conv3x3 res_after_act silu", 2, 28, 28, 64, 64, 3, 1, 1, 1, act=2, res=True,
                    res_after_act=True)
    return ok


def group_conv_s2():
    ok = True
    ok &= conv_case("conv1x1 s2 56x56 256->512", 2, 56, 56, 256, 512, 1, 2, 0, 1)
    ok &= conv_case("conv3x3 s2 p1 56x56 128->128", 2, 56, 56, 128, 128, 3, 2, 1, 1, act=1)
    ok &= conv_case("conv3x3 s2 p1 14x14 512->512", 8, 14, 14, 512, 512, 3, 2, 1, 1)
    ok &= conv_case("conv3x3 s2 p1 224x224 8->48 (effnet stem)", 2, 224, 224, 8, 48, 3, 2, 1, 1, act=2)
    return ok


def group_stem():
    n, h, w, cout = 3, 224, 224, 64
    g = torch.Generator().manual_seed(0)
    x = torch.rand(n, 3, h, w, generator=g).to(DEV)
    wt = (torch.randn(cout, 3, 7, 7, generator=g) * 0.1).to(DEV)
    bias = torch.randn(cout, generator=g).to(DEV)
    xpad = ops.pack_stem_input(x)
    torch.cuda.synchronize()
    ref_pad = torch.zeros(n, h + 6, w + 8, 8, device=DEV)
    ref_pad[:, 3:h + 3, 3:w + 3, :3] = x.permute(0, 2, 3, 1)
    ok = report("pack_stem_input", xpad, ref_pad.to(torch.bfloat16), tol=1e-6)
    # weights [cout, 7 (r), 8 (s), 8 (c)]
    wp = torch.zeros(cout, 7, 8, 8, device=DEV)
    wp[:, :, :7, :3] = wt.permute(0, 2, 3, 1)
    wp = wp.to(torch.bfloat16)
    y = ops.conv_stem7x7(xpad, wp.reshape(cout, -1), bias, n=n, h=h, w=w, cout=cout, act=1)
    torch.cuda.synchronize()
    ref = F.relu(F.conv2d(x.to(torch.bfloat16).float(), wt.to(torch.bfloat16).float(), bias, stride=2,
                          padding=3))
    ok &= report("conv_stem7x7 + relu", y, ref.permute(0, 2, 3, 1))
    return ok


def group_pointwise():
    ok = True
    x = rand_bf16(4, 112, 112, 64)
    y = ops.maxpool2d(x, 3, 2, 1)
    ref = F.max_pool2d(x.float().permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
    ok &= report("maxpool 3x3 s2 p1", y, ref, tol=1e-6)
    y = ops.maxpool2d(x, 2, 2, 0)
    ref = F.max_pool2d(x.float().permute(0, 3, 1, 2), 2, 2, 0).permute(0, 2, 3, 1)
    ok &= report("maxpool 2x2 s2", y, ref, tol=1e-6)
    y = ops.avgpool2d(x, 2, 2)
    ref = F.avg_pool2d(x.float().permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1)
    ok &= report("avgpool 2x2 s2", y, ref, tol=1e-2)
    x7 = rand_bf16(6, 7, 7, 2048)
    y = ops.adaptive_avgpool(x7, 1, 1)
    ok &= report("global avgpool", y, x7.float().mean((1, 2), keepdim=True), tol=1e-2)
    xi = torch.rand(3, 3, 32, 40, device=DEV)
    y = ops.nchw_to_nhwc(xi, 8)
    ref = torch.zeros(3, 32, 40, 8, device=DEV)
    ref[..., :3] = xi.permute(0, 2, 3, 1)
    ok &= report("nchw->nhwc pad8", y, ref.to(torch.bfloat16), tol=1e-6)
    xb = rand_bf16(2, 9, 11, 40)
    y = ops.nhwc_to_nchw(xb)
    ok &= report("nhwc->nchw f32", y, xb.float().permute(0, 3, 1, 2), tol=1e-6)
    # layernorm
    xl = rand_bf16(1000, 768, scale=2.0)
    gmm = torch.randn(768, device=DEV)
    bta = torch.randn(768, device=DEV)
    y = ops.layernorm(xl, gmm, bta, 1e-5)
    ok &= report("layernorm 768", y, F.layer_norm(xl.float(), (768,), gmm, bta, 1e-5))
    xl = rand_bf16(333, 96, scale=2.0)
    y = ops.layernorm(xl, gmm[:96].contiguous(), bta[:96].contiguous(), 1e-5)
    ok &= report("layernorm 96", y, F.layer_norm(xl.float(), (96,), gmm[:96], bta[:96], 1e-5))
    # patchify + tokens
    xi = torch.rand(2, 3, 224, 224, device=DEV)
    rows = ops.patchify(xi, 16)
    ref = xi.reshape(2, 3, 14, 16, 14, 16).permute(0, 2, 4, 1, 3, 5).reshape(2 * 196, 768)
    ok &= report("patchify", rows, ref.to(torch.bfloat16), tol=1e-6)
    cls = torch.randn(768, device=DEV)
    pos = torch.randn(197, 768, device=DEV)
    tok = ops.vit_assemble_tokens(rows, cls, pos, 2, 196, 768)
    ref = torch.cat([cls.expand(2, 1, 768), rows.float().reshape(2, 196, 768)], 1) + pos
    ok &= report("assemble tokens", tok, ref.reshape(-1, 768), tol=1e-2)
    gr = ops.gather_rows(tok, 2, 197, 0)
    ok &= report("gather rows", gr, tok.reshape(2, 197, 768)[:, 0], tol=1e-6)
    return ok


def group_attention():
    ok = True
    for (imgs, tokens, heads) in [(2, 197, 12), (1, 64, 3), (3, 50, 6), (1, 785, 6)]:
        qkv = rand_bf16(imgs * tokens, 3 * heads * 64, seed=tokens)
        out = ops.attention(qkv, imgs, tokens, heads, 64, 0.125)
        torch.cuda.synchronize()
        q, k, v = qkv.float().reshape(imgs, tokens, 3, heads, 64).permute(2, 0, 3, 1, 4)
        att = torch.softmax(q @ k.transpose(-1, -2) * 0.125, -1)
        ref = (att @ v).permute(0, 2, 1, 3).reshape(imgs * tokens, heads * 64)
        ok &= report(f"attention imgs={imgs} tokens={tokens} heads={heads}", out, ref)
    return ok


GROUPS = {"gemm": group_gemm, "conv": group_conv, "conv_s2": group_conv_s2, "stem": group_stem,
          "pointwise": group_pointwise, "attention": group_attention}

if __name__ == "__main__":
    names = sys.argv[1:] or list(GROUPS)
    _lib.init(0)
    print("lib:", _lib.load().eqxv_version().decode(), "SMs:", _lib.load().eqxv_sm_count(), flush=True)
    allok = True
    for nme in names:
        print(f"=== {nme} ===", flush=True)
        try:
            allok &= bool(GROUPS[nme]())
        except Exception as ex:  # noqa: BLE001
            allok = False
            print(f"[EXC] {nme}: {type(ex).__name__}: {ex}", flush=True)
    print("PROBE", "PASS" if allok else "FAIL")
    sys.exit(0 if allok else 1)
