"""Kernel-level parity on a B200: every C-ABI compute entry against a torch fp32 functional
reference on bf16-rounded inputs. bf16 outputs: rel-L2 <= 4e-3 (one rounding, 2^-9 relative),
fp32 outputs: rel-L2 <= 1e-5; pure data-movement kernels must be bit exact."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

ACTS = {0: lambda v: v, 1: F.relu, 2: F.silu, 3: lambda v: F.gelu(v, approximate="tanh"), 4: F.hardswish,
        5: torch.sigmoid, 6: F.hardsigmoid, 7: F.relu6}
TOL_BF16 = 4e-3


def rb(device, *shape, scale=1.0, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(torch.bfloat16).to(device)


def rel_l2(got, ref):
    got, ref = got.float(), ref.float()
    assert not torch.isnan(got).any()
    return ((got - ref).norm() / (ref.norm() + 1e-12)).item()


CONV_CASES = [
    # n, h, w, cin, cout, k, stride, pad, dil, act, res, res_after
    (4, 56, 56, 64, 256, 1, 1, 0, 1, 1, True, False),     # R50 layer1 expand + identity + relu
    (2, 56, 56, 64, 64, 3, 1, 1, 1, 1, False, False),
    (8, 28, 28, 128, 128, 3, 1, 1, 1, 0, False, False),
    (32, 14, 14, 256, 256, 3, 1, 1, 1, 1, False, False),
    (5, 7, 7, 512, 512, 3, 1, 1, 1, 0, False, False),      # n not a multiple of the image tile
    (3, 17, 13, 72, 40, 3, 1, 1, 1, 2, False, False),      # ragged everything
    (2, 32, 32, 64, 64, 3, 1, 2, 2, 0, False, False),      # dilation 2 (DeepLab backbone)
    (2, 16, 16, 64, 64, 3, 1, 12, 12, 0, False, False),    # ASPP d12: taps skipped when fully padded
    (2, 28, 28, 64, 64, 3, 1, 1, 1, 2, True, True),        # act then residual (FusedMBConv order)
    (2, 56, 56, 256, 512, 1, 2, 0, 1, 0, False, False),    # downsample 1x1 stride 2
    (2, 56, 56, 128, 128, 3, 2, 1, 1, 1, False, False),    # 3x3 stride 2
    (8, 14, 14, 512, 512, 3, 2, 1, 1, 0, False, False),
    (2, 224, 224, 8, 48, 3, 2, 1, 1, 2, False, False),     # EfficientNet stem on the padded image
    (1, 1, 1, 2048, 272, 1, 1, 0, 1, 4, False, False),     # SE-style 1x1 on a 1x1 map, odd N
    # halo-tile path (stride 1, maps that tile into 8x16 blocks): one staged tile serves all taps
    (2, 112, 112, 24, 24, 3, 1, 1, 1, 2, True, True),      # FusedMBConv single conv: act then residual, cin < 64
    (1, 56, 56, 128, 320, 3, 1, 1, 1, 1, True, False),     # two n-tiles + residual before relu
    (2, 64, 64, 64, 64, 3, 1, 2, 2, 1, False, False),      # dilation 2 halo
    (1, 224, 224, 64, 64, 3, 1, 1, 1, 1, False, False),    # VGG conv1_2
    (2, 48, 40, 72, 40, 5, 1, 2, 1, 0, False, False),      # 5x5, channels not a multiple of 64
    (3, 56, 56, 256, 256, 3, 1, 1, 1, 0, False, False),    # several K blocks through the A ring
    (2, 30, 23, 64, 96, 3, 1, 1, 1, 1, False, False),      # ragged edges inside the halo tiles
    # BASELINE configs[4] shapes (DeepLabV3-R50 @512^2: 64x64 maps, deeplabv3.py:38-55,77-135): the dilated taps
    # overlap the map only PARTIALLY, so the producer's and the issuer's tap-skip predicates must agree tile by tile
    (1, 64, 64, 2048, 256, 3, 1, 12, 12, 1, False, False),  # ASPP d12
    (1, 64, 64, 2048, 256, 3, 1, 24, 24, 1, False, False),  # ASPP d24
    (1, 64, 64, 2048, 256, 3, 1, 36, 36, 1, False, False),  # ASPP d36
    (2, 64, 64, 256, 256, 3, 1, 2, 2, 1, False, False),     # layer3 conv2, dilation 2 (resnet.py:286-333)
    (1, 64, 64, 512, 512, 3, 1, 4, 4, 1, False, False),     # layer4 conv2, dilation 4
    (1, 64, 64, 512, 512, 3, 1, 2, 2, 1, False, False),     # layer4 block 0 conv2: previous dilation
    (1, 64, 64, 1024, 256, 1, 1, 0, 1, 1, False, False),    # layer3 conv1 on the dilated (stride-8) map
    (1, 64, 64, 2048, 512, 3, 1, 1, 1, 1, False, False),    # FCNHead 3x3 on 2048 channels (fcn.py:19-34)
    (1, 64, 64, 1280, 256, 1, 1, 0, 1, 1, False, False),    # ASPP projection over the 5-branch concat
    (3, 33, 33, 320, 256, 3, 1, 12, 12, 0, False, False),   # odd map: ragged tiles AND partial tap overlap
    (2, 40, 28, 128, 64, 3, 1, 24, 24, 0, False, False),    # d24 on a map where some taps of some tiles vanish
]


@pytest.mark.parametrize("case", CONV_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv2d_igemm(device, case):
    from eqxvision_b200 import ops

    n, h, w, cin, cout, k, stride, pad, dil, act, res, res_after = case
    x = rb(device, n, h, w, cin, seed=1)
    wt = rb(device, cout, k, k, cin, scale=(k * k * cin) ** -0.5, seed=2)
    bias = torch.randn(cout, generator=torch.Generator().manual_seed(3)).to(device)
    ho, wo = ops.conv_out_size(h, k, stride, pad, dil), ops.conv_out_size(w, k, stride, pad, dil)
    r = rb(device, n, ho, wo, cout, seed=4) if res else None
    y = ops.conv2d(x, wt.reshape(cout, -1), bias, cin=cin, cout=cout, kh=k, kw=k, stride=stride, pad=pad,
                   dil=dil, act=act, residual=r, res_after_act=res_after)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.float().permute(0, 3, 1, 2), bias, stride=stride,
                   padding=pad, dilation=dil)
    rr = r.float().permute(0, 3, 1, 2) if res else 0
    ref = ACTS[act](ref) + rr if res_after else ACTS[act](ref + rr)
    assert rel_l2(y, ref.permute(0, 2, 3, 1)) < TOL_BF16


# n, h, w, channels, groups, k, stride, dil, act, res
GROUPED_CASES = [(2, 56, 56, 128, 32, 3, 1, 1, 1, False), (2, 28, 28, 256, 32, 3, 2, 1, 1, False),
                 (3, 14, 14, 512, 32, 3, 1, 1, 1, True), (2, 7, 7, 1024, 32, 3, 1, 1, 0, False),
                 (1, 16, 16, 256, 32, 3, 1, 2, 1, False), (2, 9, 11, 64, 1, 1, 1, 1, 2, False),
                 (2, 14, 14, 128, 2, 3, 1, 1, 0, False)]


@pytest.mark.parametrize("case", GROUPED_CASES, ids=lambda c: "x".join(map(str, c)))
def test_grouped_conv_block_diagonal(device, case):
    """equinox.nn.Conv2d(groups=G) as used by ResNeXt (resnet.py:19-23, 83) through the block-diagonal
    64-channel packing (EQXV_FLAG_GROUPED_BLOCK64) against torch's grouped convolution"""
    from eqxvision_b200 import _pack, ops

    n, h, w, c, groups, k, stride, dil, act, res = case
    pad = dil * (k - 1) // 2
    x = rb(device, n, h, w, c, seed=1)
    wt = torch.randn(c, c // groups, k, k, generator=torch.Generator().manual_seed(2)) * (k * k * c / groups) ** -0.5
    wt = wt.to(torch.bfloat16).float()
    bias = torch.randn(c, generator=torch.Generator().manual_seed(3)).to(device)
    ho, wo = ops.conv_out_size(h, k, stride, pad, dil), ops.conv_out_size(w, k, stride, pad, dil)
    r = rb(device, n, ho, wo, c, seed=4) if res else None
    wp = _pack.pack_grouped_weight(wt, groups).to(device)
    assert wp.shape == (c, k * k * 64)
    y = ops.conv2d(x, wp, bias, cin=c, cout=c, kh=k, kw=k, stride=stride, pad=pad, dil=dil, act=act, residual=r,
                   grouped_block64=True)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.to(device), bias, stride=stride, padding=pad, dilation=dil,
                   groups=groups)
    if res:
        ref = ref + r.float().permute(0, 3, 1, 2)
    assert rel_l2(y, ACTS[act](ref).permute(0, 2, 3, 1)) < TOL_BF16


GEMM_CASES = [
    (128, 64, 64, 0, False, False), (1000, 128, 192, 0, False, False), (40000, 64, 64, 1, False, False),
    (12608, 2304, 768, 0, False, False), (12608, 768, 3072, 0, True, False), (12608, 3072, 768, 3, False, False),
    (256, 1000, 2048, 0, False, True), (300, 272, 1632, 0, False, False), (512, 24, 144, 2, False, False),
    (512, 64, 24, 1, False, False), (7, 448, 8, 5, False, False), (130, 2688, 112, 6, False, False),
    # all-epilogue tiles (one K block, no residual, many chunks): the sixteen-warp epilogue (igemm.cu epilogue_warps16)
    (200704, 144, 24, 2, False, False), (100352, 192, 32, 2, False, False), (50176, 336, 56, 2, False, False),
    (160000, 64, 64, 1, False, False), (70000, 272, 128, 4, False, False),
    # one row: one warp per output column (csrc/gemv.cu; AlexNet's single-image classifier)
    (1, 4096, 9216, 1, False, False), (1, 1000, 4096, 0, False, True), (1, 272, 1632, 4, False, False), (1, 1000, 768, 0, False, True),
    (1, 4096, 4096, 1, False, False), (1, 72, 40, 2, False, False),
]


@pytest.mark.parametrize("case", GEMM_CASES, ids=lambda c: "x".join(map(str, c)))
def test_gemm(device, case):
    from eqxvision_b200 import ops

    m, n, k, act, res, f32 = case
    a = rb(device, m, k, seed=1)
    wt = rb(device, n, k, scale=k ** -0.5, seed=2)
    bias = torch.randn(n, generator=torch.Generator().manual_seed(3)).to(device)
    r = rb(device, m, n, seed=4) if res else None
    y = ops.gemm(a, wt, bias, act=act, residual=r, out_f32=f32)
    ref = ACTS[act](a.float() @ wt.float().t() + bias + (r.float() if res else 0))
    assert rel_l2(y, ref) < (1e-5 if f32 else TOL_BF16)


def test_gemm_strided_views_and_padding_columns_stay_untouched(device):
    from eqxvision_b200 import ops

    big_in = rb(device, 300, 96, seed=1)
    big_out = torch.full((300, 80), 7.0, dtype=torch.bfloat16, device=device)
    a = big_in[:, 16:80]                      # pitch 96, 64 columns, 32-byte offset
    out = big_out[:, 8:48]                    # pitch 80, 40 columns
    wt = rb(device, 40, 64, scale=0.125, seed=2)
    ops.gemm(a, wt, None, out=out)
    ref = a.float() @ wt.float().t()
    assert rel_l2(out, ref) < TOL_BF16
    assert (big_out[:, :8] == 7).all() and (big_out[:, 48:] == 7).all()


def test_stem(device):
    from eqxvision_b200 import _pack, ops

    n, h, w, cout = 3, 224, 224, 64
    g = torch.Generator().manual_seed(0)
    x = torch.rand(n, 3, h, w, generator=g).to(device)
    wt = torch.randn(cout, 3, 7, 7, generator=g) * 0.1
    bias = torch.randn(cout, generator=g).to(device)
    xpad = ops.pack_stem_input(x)
    ref_pad = torch.zeros(n, h + 6, w + 8, 8, device=device)
    ref_pad[:, 3:h + 3, 3:w + 3, :3] = x.permute(0, 2, 3, 1)
    assert torch.equal(xpad, ref_pad.to(torch.bfloat16))
    y = ops.conv_stem7x7(xpad, _pack.pack_stem_weight(wt).to(device), bias, n=n, h=h, w=w, cout=cout, act=1)
    ref = F.relu(F.conv2d(x.to(torch.bfloat16).float(), wt.to(torch.bfloat16).float().to(device), bias,
                          stride=2, padding=3))
    assert rel_l2(y, ref.permute(0, 2, 3, 1)) < TOL_BF16


def test_pooling_and_layout(device):
    from eqxvision_b200 import ops

    x = rb(device, 4, 112, 112, 64)
    xn = x.float().permute(0, 3, 1, 2)
    assert torch.equal(ops.maxpool2d(x, 3, 2, 1).float(), F.max_pool2d(xn, 3, 2, 1).permute(0, 2, 3, 1))
    assert torch.equal(ops.maxpool2d(x, 2, 2, 0).float(), F.max_pool2d(xn, 2, 2, 0).permute(0, 2, 3, 1))
    assert rel_l2(ops.avgpool2d(x, 2, 2), F.avg_pool2d(xn, 2, 2).permute(0, 2, 3, 1)) < TOL_BF16
    x7 = rb(device, 6, 7, 7, 2048)
    assert rel_l2(ops.adaptive_avgpool(x7, 1, 1), x7.float().mean((1, 2), keepdim=True)) < TOL_BF16
    x14 = rb(device, 2, 14, 14, 512)
    ref = F.adaptive_avg_pool2d(x14.float().permute(0, 3, 1, 2), 7).permute(0, 2, 3, 1)
    assert rel_l2(ops.adaptive_avgpool(x14, 7, 7), ref) < TOL_BF16
    xi = torch.rand(3, 3, 32, 40, device=device)
    ref = torch.zeros(3, 32, 40, 8, device=device)
    ref[..., :3] = xi.permute(0, 2, 3, 1)
    assert torch.equal(ops.nchw_to_nhwc(xi, 8), ref.to(torch.bfloat16))
    xb = rb(device, 2, 9, 11, 40)
    assert torch.equal(ops.nhwc_to_nchw(xb), xb.float().permute(0, 3, 1, 2))


@pytest.mark.parametrize("rows,d", [(1000, 768), (333, 96), (64, 2048), (5, 192), (12608, 768), (700, 1024), (3, 256),
                                     (50, 512)])
def test_layernorm(device, rows, d):
    from eqxvision_b200 import ops

    x = rb(device, rows, d, scale=2.0)
    g = torch.randn(d, device=device)
    b = torch.randn(d, device=device)
    assert rel_l2(ops.layernorm(x, g, b, 1e-5), F.layer_norm(x.float(), (d,), g, b, 1e-5)) < TOL_BF16


@pytest.mark.parametrize("imgs,tokens,heads", [(2, 197, 12), (1, 64, 3), (3, 50, 6), (1, 785, 6), (2, 1, 2), (64, 197, 12),
                                               (5, 256, 4), (3, 128, 2), (2, 129, 3), (7, 210, 5), (1, 16, 1)])
def test_attention(device, imgs, tokens, heads):
    from eqxvision_b200 import ops

    qkv = rb(device, imgs * tokens, 3 * heads * 64, seed=tokens)
    out = ops.attention(qkv, imgs, tokens, heads, 64, 0.125)
    q, k, v = qkv.float().reshape(imgs, tokens, 3, heads, 64).permute(2, 0, 3, 1, 4)
    att = torch.softmax(q @ k.transpose(-1, -2) * 0.125, -1)
    ref = (att @ v).permute(0, 2, 1, 3).reshape(imgs * tokens, heads * 64)
    assert rel_l2(out, ref) < 6e-3  # P is rounded to bf16 before P.V


@pytest.mark.parametrize("imgs,tokens,heads", [(2, 197, 12), (1, 64, 3), (3, 50, 6), (1, 785, 2), (2, 1, 2), (1, 33, 1)])
def test_attention_probabilities_output(device, imgs, tokens, heads):
    """the materialised softmax matrix (vit.py:70, returned by _VitBlock(return_attention=True)): fp32,
    rows sum to one, equal to the fp32 softmax of the bf16 q/k the device holds"""
    from eqxvision_b200 import ops

    qkv = rb(device, imgs * tokens, 3 * heads * 64, seed=tokens + 1)
    probs = ops.attention_probs(qkv, imgs, tokens, heads, 64, 0.125)
    assert probs.shape == (imgs, heads, tokens, tokens) and probs.dtype == torch.float32
    q, k, _ = qkv.float().reshape(imgs, tokens, 3, heads, 64).permute(2, 0, 3, 1, 4)
    ref = torch.softmax(q @ k.transpose(-1, -2) * 0.125, -1)
    assert (probs - ref).abs().max().item() < 2e-5          # fp32 accumulation order only
    assert (probs.sum(-1) - 1).abs().max().item() < 1e-5


def test_vit_glue(device):
    from eqxvision_b200 import ops

    xi = torch.rand(2, 3, 224, 224, device=device)
    rows = ops.patchify(xi, 16)
    ref = xi.reshape(2, 3, 14, 16, 14, 16).permute(0, 2, 4, 1, 3, 5).reshape(2 * 196, 768)
    assert torch.equal(rows, ref.to(torch.bfloat16))
    cls = torch.randn(768, device=device)
    pos = torch.randn(197, 768, device=device)
    tok = ops.vit_assemble_tokens(rows, cls, pos, 2, 196, 768)
    ref = torch.cat([cls.expand(2, 1, 768), rows.float().reshape(2, 196, 768)], 1) + pos
    assert rel_l2(tok, ref.reshape(-1, 768)) < TOL_BF16
    assert torch.equal(ops.gather_rows(tok, 2, 197, 0), tok.reshape(2, 197, 768)[:, 0])


def test_argument_errors_are_reported_not_crashed(device):
    from eqxvision_b200 import _lib, ops

    x = rb(device, 1, 8, 8, 12)  # 12 channels: not a multiple of 8
    with pytest.raises(_lib.EqxvError, match="multiple of 8"):
        ops.conv2d(x, rb(device, 16, 12), None, cin=12, cout=16, kh=1, kw=1)
    with pytest.raises(_lib.EqxvError, match="unsupported"):
        ops.attention(rb(device, 8, 3 * 2 * 32), 1, 8, 2, 32, 0.1)
    with pytest.raises(_lib.EqxvError):
        ops.conv2d(torch.zeros(1, 8, 8, 8, dtype=torch.bfloat16), rb(device, 16, 8), None, cin=8, cout=16,
                   kh=1, kw=1)  # CPU tensor: no fallback


@pytest.mark.parametrize("kh,kw,stride,pad,cin,cout,hw", [(7, 7, 2, 3, 3, 64, 224), (3, 3, 1, 1, 3, 64, 64),
                                                        (3, 3, 2, 1, 3, 48, 224), (3, 3, 2, 1, 3, 16, 96),
                                                        (5, 5, 1, 2, 4, 40, 33), (4, 4, 4, 0, 3, 96, 224),
                                                        (7, 7, 2, 3, 3, 64, 50), (3, 3, 2, 1, 3, 24, 31)])
def test_stem_conv_variants(device, kh, kw, stride, pad, cin, cout, hw):
    """first-layer convs on the raw fp32 image (ResNet/DenseNet 7x7 s2, VGG 3x3 s1, EfficientNet/MobileNet 3x3 s2)"""
    from eqxvision_b200 import _pack, ops

    g = torch.Generator().manual_seed(1)
    x = torch.randn(2, cin, hw, hw, generator=g).to(device)
    wt = torch.randn(cout, cin, kh, kw, generator=g) * (cin * kh * kw) ** -0.5
    bias = torch.randn(cout, generator=g).to(device)
    xpad = ops.pack_stem_input(x, pad=pad)
    y = ops.conv_stem(xpad, _pack.pack_stem_weight(wt).to(device), bias, n=2, h=hw, w=hw, cout=cout, kh=kh, kw=kw,
                      stride=stride, pad=pad, act=2)
    ref = F.silu(F.conv2d(x.to(torch.bfloat16).float(), wt.to(torch.bfloat16).float().to(device), bias,
                          stride=stride, padding=pad))
    assert rel_l2(y, ref.permute(0, 2, 3, 1)) < TOL_BF16


@pytest.mark.parametrize("c,hw,k,stride,dil,act", [(48, 112, 3, 1, 1, 2), (144, 112, 3, 2, 1, 2), (336, 28, 5, 1, 1, 2),
                                                   (192, 56, 5, 2, 1, 4), (960, 14, 5, 1, 1, 4), (96, 7, 5, 1, 1, 1),
                                                   (72, 17, 3, 1, 2, 1), (40, 3, 5, 1, 1, 0),
                                                   # shared-memory stencil path: ragged tiles, partial channel blocks,
                                                   # every (k, stride) instantiation, all three tile widths
                                                   (64, 56, 3, 1, 1, 2), (24, 33, 3, 2, 1, 2), (200, 15, 3, 1, 1, 1),
                                                   (136, 29, 5, 2, 1, 4), (8, 9, 5, 1, 1, 0), (1632, 7, 5, 1, 1, 2),
                                                   (672, 14, 5, 2, 1, 2), (32, 112, 3, 1, 1, 2), (88, 25, 3, 2, 1, 6)])
@pytest.mark.parametrize("tile", [False, True], ids=["strip", "smem_stencil"])
def test_depthwise(device, c, hw, k, stride, dil, act, tile):
    from eqxvision_b200 import _pack, ops

    if tile and dil != 1:
        pytest.skip("the shared-memory stencil kernel is dilation-1 only")

    n = 3
    pad = (k - 1) // 2 * dil
    x = rb(device, n, hw, hw, c, seed=1)
    g = torch.Generator().manual_seed(2)
    wt = torch.randn(c, 1, k, k, generator=g) * (k * k) ** -0.5
    bias = torch.randn(c, generator=g)
    y = ops.dwconv(x, _pack.pack_depthwise_weight(wt, c).to(device), bias.to(device), k=k, stride=stride, pad=pad,
                   dil=dil, act=act, tile=tile)
    ref = ACTS[act](F.conv2d(x.float().permute(0, 3, 1, 2), wt.to(device), bias.to(device), stride=stride,
                             padding=pad, dilation=dil, groups=c))
    assert rel_l2(y, ref.permute(0, 2, 3, 1)) < TOL_BF16


def test_eltwise_variants(device):
    from eqxvision_b200 import ops

    rows, c, hw = 4 * 49, 272, 49
    x = rb(device, rows, c, seed=1)
    o = rb(device, rows, c, seed=2)
    gate = rb(device, 4, c, seed=3)
    sc = torch.rand(c, device=device) + 0.5
    sh = torch.randn(c, device=device)
    y = ops.eltwise(x, scale=sc, shift=sh, act=1)                         # BatchNorm + ReLU
    assert rel_l2(y, F.relu(x.float() * sc + sh)) < TOL_BF16
    y = ops.eltwise(x, other=o, act=1)                                    # residual add + ReLU
    assert rel_l2(y, F.relu(x.float() + o.float())) < TOL_BF16
    y = ops.eltwise(x, gate=gate, rows_per_image=hw)                      # SE gate
    ref = x.float().reshape(4, hw, c) * gate.float().reshape(4, 1, c)
    assert rel_l2(y, ref.reshape(rows, c)) < TOL_BF16
    big = torch.zeros(rows, 512, dtype=torch.bfloat16, device=device)
    ops.eltwise(x, act=4, out=big[:, 64:64 + c])                          # strided destination slice
    assert rel_l2(big[:, 64:64 + c], F.hardswish(x.float())) < TOL_BF16
    assert (big[:, :64] == 0).all() and (big[:, 64 + c:] == 0).all()


def test_global_avgpool_shapes(device):
    from eqxvision_b200 import ops

    # maps of >= 1024 pixels take the 8-CTA cluster kernel (partial sums meet through distributed shared memory),
    # smaller ones the single-CTA kernel; (3, 33, 72): pixel count not divisible by the cluster size
    for (n, hw, c) in [(5, 112, 48), (2, 64, 2048), (3, 14, 960), (7, 7, 24), (2, 5, 8), (3, 33, 72), (2, 56, 144),
                       (1, 32, 8)]:
        x = rb(device, n, hw, hw, c, seed=c)
        got = ops.adaptive_avgpool(x, 1, 1)
        assert rel_l2(got, x.float().mean((1, 2), keepdim=True)) < TOL_BF16
        assert torch.equal(got, ops.adaptive_avgpool(x, 1, 1))      # deterministic (fixed reduction order)


def test_bilinear_resize(device):
    from eqxvision_b200 import ops

    x = rb(device, 2, 16, 12, 24, seed=1)
    xn = x.float().permute(0, 3, 1, 2)
    y = ops.resize_bilinear_to_nchw(x, 21, 128, 96)
    ref = F.interpolate(xn[:, :21], size=(128, 96), mode="bilinear", align_corners=False)
    assert torch.allclose(y, ref, atol=1e-5)
    for (oh, ow) in ((64, 50), (33, 517), (16, 12)):      # widths that are not multiples of 4 / wider than one block
        y2 = ops.resize_bilinear_to_nchw(x, 21, oh, ow)
        ref2 = F.interpolate(xn[:, :21], size=(oh, ow), mode="bilinear", align_corners=False)
        assert torch.allclose(y2, ref2, atol=1e-5), (oh, ow)
    yb = ops.resize_bilinear(x, 40, 30)
    refb = F.interpolate(xn, size=(40, 30), mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
    assert rel_l2(yb, refb) < TOL_BF16
    one = rb(device, 2, 1, 1, 256, seed=2)                               # ASPP pooling branch: pure broadcast
    assert torch.equal(ops.resize_bilinear(one, 8, 8), one.expand(2, 8, 8, 256))


def test_copy2d(device):
    from eqxvision_b200 import ops

    src = rb(device, 100, 64, seed=1)
    dst = torch.zeros(100, 160, dtype=torch.bfloat16, device=device)
    ops.copy2d(dst[:, 32:96], src)
    assert torch.equal(dst[:, 32:96], src) and (dst[:, :32] == 0).all() and (dst[:, 96:] == 0).all()


@pytest.mark.parametrize("n,h,w,heads,shift", [(2, 14, 14, 3, (0, 0)), (2, 14, 14, 3, (3, 3)), (1, 56, 56, 3, (3, 3)),
                                               (3, 7, 7, 24, (0, 0)), (1, 28, 14, 6, (3, 3))])
def test_window_attention(device, n, h, w, heads, shift):
    """eqxv_window_attention_bf16 vs the oracle's swin.py:117-253 restatement (identity qkv/proj)"""
    from eqxvision_b200 import ops
    from oracle import models as om

    BF16 = torch.bfloat16
    ws, d = 7, 32
    c = heads * d
    g = torch.Generator().manual_seed(h * 100 + heads + shift[0])
    qkv = torch.randn((n, h, w, 3 * c), generator=g).to(BF16)
    table = 0.5 * torch.randn(((2 * ws - 1) ** 2, heads), generator=g)
    idx = torch.randint(0, (2 * ws - 1) ** 2, (ws ** 4,), generator=g)
    bias = table[idx].reshape(ws * ws, ws * ws, heads).permute(2, 0, 1).contiguous()
    got = ops.window_attention(qkv.reshape(-1, 3 * c).to(device), bias.to(device), n=n, h=h, w=w, heads=heads,
                               head_dim=d, window=ws, shift=shift, scale=d ** -0.5)
    torch.cuda.synchronize()
    # oracle: feed qkv through identity "linears" by calling the attention core on pre-computed qkv
    x = qkv.float()
    sh = [0 if ws >= h else shift[0], 0 if ws >= w else shift[1]]
    xr = torch.roll(x, shifts=(-sh[0], -sh[1]), dims=(1, 2)) if sum(sh) else x
    nw = (h // ws) * (w // ws)
    xw = xr.reshape(n, h // ws, ws, w // ws, ws, 3 * c).permute(0, 1, 3, 2, 4, 5).reshape(n * nw, ws * ws, 3 * c)
    q, k, v = xw.reshape(n * nw, ws * ws, 3, heads, d).permute(2, 0, 3, 1, 4)
    logits = (q * d ** -0.5) @ k.transpose(-1, -2) + bias
    if sum(sh):
        mask = om.swin_shift_mask(h, w, ws, sh)
        logits = (logits.reshape(n, nw, heads, ws * ws, ws * ws) + mask[None, :, None]).reshape(logits.shape)
    out = (torch.softmax(logits, -1) @ v).permute(0, 2, 1, 3).reshape(n * nw, ws * ws, c)
    out = out.reshape(n, h // ws, w // ws, ws, ws, c).permute(0, 1, 3, 2, 4, 5).reshape(n, h, w, c)
    if sum(sh):
        out = torch.roll(out, shifts=(sh[0], sh[1]), dims=(1, 2))
    assert rel_l2(got.float().cpu().reshape(n, h, w, c), out) < 4e-3


def test_patch_merge_is_bit_exact(device):
    from eqxvision_b200 import ops

    BF16 = torch.bfloat16
    g = torch.Generator().manual_seed(5)
    x = torch.randn((3, 8, 6, 24), generator=g).to(BF16)
    got = ops.patch_merge(x.to(device)).cpu()
    ref = torch.cat([x[:, 0::2, 0::2], x[:, 1::2, 0::2], x[:, 0::2, 1::2], x[:, 1::2, 1::2]], -1)  # swin.py:26-31
    assert torch.equal(got, ref)


@pytest.mark.parametrize("c,hw,k,stride,act,n", [(48, 112, 3, 1, 2, 3), (144, 112, 3, 2, 2, 2), (336, 28, 5, 1, 2, 3),
                                                 (192, 56, 5, 2, 4, 2), (960, 14, 5, 1, 4, 5), (1632, 7, 5, 1, 2, 4),
                                                 (2688, 7, 3, 1, 2, 3), (72, 17, 3, 1, 1, 2), (40, 9, 5, 2, 0, 3),
                                                 (8, 5, 3, 1, 2, 1), (672, 14, 5, 2, 2, 130)])
def test_depthwise_with_fused_squeeze(device, c, hw, k, stride, act, n):
    """eqxv_dwconv_bn_act_pool_bf16: the depthwise output AND the SqueezeExcitation squeeze (squeeze.py:52) from one kernel.
    y against torch's grouped conv; pooled against the mean of the kernel's own stored output (that is what the next layer
    reads); bitwise reproducible and independent of where an image sits in the batch."""
    from eqxvision_b200 import _pack, ops

    pad = (k - 1) // 2
    x = rb(device, n, hw, hw, c, seed=1)
    g = torch.Generator().manual_seed(2)
    wt = torch.randn(c, 1, k, k, generator=g) * (k * k) ** -0.5
    bias = torch.randn(c, generator=g)
    wp, bd = _pack.pack_depthwise_weight(wt, c).to(device), bias.to(device)
    y, pooled = ops.dwconv_pool(x, wp, bd, k=k, stride=stride, pad=pad, act=act)
    ref = ACTS[act](F.conv2d(x.float().permute(0, 3, 1, 2), wt.to(device), bd, stride=stride, padding=pad, groups=c))
    assert rel_l2(y, ref.permute(0, 2, 3, 1)) < TOL_BF16
    want = y.float().mean((1, 2))
    assert pooled.shape == (n, c)
    assert (pooled.float() - want).abs().max().item() <= 2 ** -7 * want.abs().max().item() + 1e-6
    assert rel_l2(pooled, want) < 3e-3
    y2, pooled2 = ops.dwconv_pool(x, wp, bd, k=k, stride=stride, pad=pad, act=act)
    assert torch.equal(y, y2) and torch.equal(pooled, pooled2)
    if n > 1:
        perm = torch.randperm(n, generator=torch.Generator().manual_seed(3)).to(device)
        y3, pooled3 = ops.dwconv_pool(x[perm].contiguous(), wp, bd, k=k, stride=stride, pad=pad, act=act)
        assert torch.equal(y3, y[perm]) and torch.equal(pooled3, pooled[perm])
    # and the plain entry (no squeeze) takes the same kernel: identical y
    assert torch.equal(ops.dwconv(x, wp, bd, k=k, stride=stride, pad=pad, act=act), y)


@pytest.mark.parametrize("m,d,n,act", [(12608, 768, 2048, 0), (12608, 768, 1536, 3), (300, 192, 576, 0), (197, 384, 1536, 3),
                                      (1000, 96, 288, 0)])
def test_layernorm_folded_into_its_neighbour_gemms(device, m, d, n, act):
    """eqxv_gemm_res_rowstats_bf16 -> eqxv_gemm_ln_act_bf16 (vit.py:149,154, mlps.py:61-65): the first GEMM adds the
    residual and emits row statistics, the second applies LayerNorm to x on the fly. Reference: torch's layer_norm on
    the stored x followed by the linear layer, fp32."""
    from eqxvision_b200 import ops

    g = torch.Generator().manual_seed(5)
    a0 = rb(device, m, d, seed=1)
    w0 = rb(device, d, d, scale=d ** -0.5, seed=2)
    b0 = torch.randn(d, generator=g).to(device)
    res = rb(device, m, d, seed=3) * 3 + 1.5                     # rows with a mean well away from zero
    stats = torch.zeros(m, (d + 63) // 64, 2, device=device)
    x = ops.gemm_rowstats(a0, w0, b0, residual=res, stats=stats)
    ref_x = a0.float() @ w0.float().t() + b0 + res.float()
    assert rel_l2(x, ref_x) < TOL_BF16
    xs = x.float()
    want = torch.stack([xs.reshape(m, -1, 64).sum(2) if d % 64 == 0 else
                        torch.stack([xs[:, c:c + 64].sum(1) for c in range(0, d, 64)], 1),
                        (xs * xs).reshape(m, -1, 64).sum(2) if d % 64 == 0 else
                        torch.stack([(xs[:, c:c + 64] ** 2).sum(1) for c in range(0, d, 64)], 1)], 2)
    assert torch.allclose(stats, want, rtol=1e-5, atol=1e-3)
    gamma = (1 + 0.2 * torch.randn(d, generator=g)).to(device)
    beta = (0.3 * torch.randn(d, generator=g)).to(device)
    w1 = (torch.randn(n, d, generator=g) * d ** -0.5).to(device)
    b1 = torch.randn(n, generator=g).to(device)
    wf = (w1 * gamma[None, :]).to(torch.bfloat16)
    y = ops.gemm_ln(x, wf, (b1.double() + w1.double() @ beta.double()).float(), wf.float().sum(1), stats, 1e-5, act=act)
    ref = ACTS[act](F.layer_norm(xs, (d,), gamma, beta, 1e-5) @ w1.t() + b1)
    assert rel_l2(y, ref) < 6e-3     # two bf16 roundings (filter w*gamma, output) against an fp32 reference


@pytest.mark.parametrize("n_img,hw,k,n,res", [(3, 49, 2688, 448, False), (5, 196, 672, 112, True), (2, 3136, 192, 32, True),
                                              (4, 12544, 48, 24, False), (7, 49, 1632, 272, True), (3, 100, 40, 24, False),
                                              (5, 12544, 24, 24, True), (3, 1000, 56, 200, False), (40, 3136, 48, 24, True)])
def test_gemm_with_se_gate_on_the_a_operand(device, n_img, hw, k, n, res):
    """eqxv_gemm_gated_bf16: out = (a * gate[image]) @ w^T + bias (+ residual) (squeeze.py:61 + efficientnet.py:161-170).
    Must equal the two-kernel path (eltwise gate, then GEMM) BITWISE: the in-smem product is rounded to bf16 exactly as
    the gate pass stored it."""
    from eqxvision_b200 import ops

    m = n_img * hw
    a = rb(device, m, k, seed=1)
    gate = torch.sigmoid(rb(device, n_img, k, seed=2).float()).to(torch.bfloat16)
    wt = rb(device, n, k, scale=k ** -0.5, seed=3)
    bias = torch.randn(n, generator=torch.Generator().manual_seed(4)).to(device)
    r = rb(device, m, n, seed=5) if res else None
    got = ops.gemm_gated(a, gate, wt, bias, rows_per_image=hw, residual=r)
    gated = ops.eltwise(a, gate=gate, rows_per_image=hw)
    want = ops.gemm(gated, wt, bias, residual=r)
    torch.cuda.synchronize()
    ref = (a.float() * gate.float().repeat_interleave(hw, 0)).to(torch.bfloat16).float() @ wt.float().t() + bias
    if res:
        ref = ref + r.float()
    assert rel_l2(got, ref) < TOL_BF16
    assert rel_l2(got, want) < 1e-3     # (tile shapes may differ between the pair and single-CTA kernels: not bitwise)


# n, h, w, downsample shortcut, output channels of the fused next conv1 (0: none)
BOTTLENECK_CASES = [(2, 56, 56, False, 0), (3, 56, 56, False, 64), (2, 56, 56, True, 64), (1, 56, 56, True, 0),
                    (3, 24, 40, False, 64),     # odd tile count: the pair's phantom tile
                    (2, 30, 23, True, 64),      # ragged edges in both directions
                    (37, 56, 56, False, 64),    # more pairs than clusters: several tiles per CTA, slot / phase wrap
                    (35, 56, 56, False, 128),   # last block of the stage + the next stage's 256 -> 128 (resnet.py:290-297)
                    (40, 56, 56, True, 64), (36, 56, 56, False, 0)]


@pytest.mark.parametrize("case", BOTTLENECK_CASES, ids=lambda c: "x".join(map(str, c)))
def test_bottleneck64_fused(device, case):
    """eqxv_bottleneck64_fused_bf16 (resnet.py:144-162 on the 64-channel trunk of layer1): against the layer-by-layer
    kernels of the same library (same rounding points: must agree to accumulation order) and against torch fp32."""
    from eqxvision_b200 import ops

    n, h, w, down, nxt = case
    t1 = rb(device, n, h, w, 64, seed=1)
    w2 = rb(device, 64, 3, 3, 64, scale=576 ** -0.5, seed=2)
    w3 = rb(device, 256, 64, scale=64 ** -0.5, seed=3)
    g = torch.Generator().manual_seed(4)
    b2, b3, b1n, bd = (torch.randn(c, generator=g).to(device) * 0.5 for c in (64, 256, max(nxt, 8), 256))
    res = rb(device, n, h, w, 256, seed=5) if not down else None
    x0 = rb(device, n, h, w, 64, seed=6) if down else None
    wd = rb(device, 256, 64, scale=64 ** -0.5, seed=7)
    w1n = rb(device, nxt, 256, scale=256 ** -0.5, seed=8) if nxt else None
    w3cat = torch.cat([w3, wd], dim=1).contiguous() if down else w3
    b3cat = b3 + bd if down else b3
    got = ops.bottleneck64(t1, w2.reshape(64, -1), b2, w3cat, b3cat, residual=res, x0=x0, w1n=w1n,
                           b1n=b1n if nxt else None)
    y, nx = got if nxt else (got, None)
    # layer by layer, same library
    t2 = ops.conv2d(t1, w2.reshape(64, -1), b2, cin=64, cout=64, kh=3, kw=3, pad=1, act=1)
    # torch fp32 on the bf16-rounded intermediates
    t2_ref = F.relu(F.conv2d(t1.float().permute(0, 3, 1, 2), w2.float().permute(0, 3, 1, 2), b2, padding=1))
    t2r = t2_ref.permute(0, 2, 3, 1).to(torch.bfloat16).float()
    assert rel_l2(t2, t2r) < TOL_BF16
    if down:
        yr = F.relu(t2r @ w3.float().t() + x0.float() @ wd.float().t() + b3cat)
    else:
        yr = F.relu(t2r @ w3.float().t() + b3 + res.float())
    assert rel_l2(y, yr) < TOL_BF16
    if not down:
        y_layer = ops.conv2d(t2, w3, b3, cin=64, cout=256, kh=1, kw=1, act=1, residual=res)
        assert rel_l2(y, y_layer) < 1e-3
    if nxt:
        nr = F.relu(y.float() @ w1n.float().t() + b1n)      # on the kernel's own (bf16) y: isolates the last GEMM
        assert rel_l2(nx, nr) < TOL_BF16
        n_layer = ops.conv2d(y, w1n, b1n, cin=256, cout=nxt, kh=1, kw=1, act=1)
        assert rel_l2(nx, n_layer) < 1e-3


def test_bottleneck64_refuses_what_does_not_fit(device):
    """downsample shortcut + a 128-wide next convolution needs 235 KB of shared memory (never occurs in ResNet: block 0
    of a stage is not its last block): the entry fails loudly instead of launching"""
    from eqxvision_b200 import _lib, ops

    t1, x0 = rb(device, 1, 16, 16, 64), rb(device, 1, 16, 16, 64)
    b = torch.zeros(256, device=device)
    with pytest.raises(_lib.EqxvError, match="shared memory"):
        ops.bottleneck64(t1, rb(device, 64, 576), b[:64], rb(device, 256, 128), b, x0=x0, w1n=rb(device, 128, 256), b1n=b[:128])
    with pytest.raises(_lib.EqxvError, match="exactly one"):
        ops.bottleneck64(t1, rb(device, 64, 576), b[:64], rb(device, 256, 64), b)


# n, h, w, cin, cout, k, pad, dil, act, res   (cin > 64, cin % 64 != 0)
TAIL_CASES = [(2, 56, 56, 144, 32, 1, 0, 1, 0, True),      # EfficientNet projection: three K chunks, single-CTA kernel
              (2, 28, 28, 336, 56, 1, 0, 1, 0, False),     # six K chunks: CTA-pair kernel
              (1, 14, 14, 1632, 272, 1, 0, 1, 2, False),   # 26 K chunks
              (2, 30, 23, 72, 40, 3, 1, 1, 1, False),      # 3x3 through the halo kernel, two chunks per tap
              (3, 17, 13, 200, 96, 3, 1, 1, 2, True),      # 3x3 through the generic kernel, ragged map
              (2, 16, 16, 136, 64, 3, 2, 2, 0, False),     # dilated
              (1, 1, 1, 144, 72, 1, 0, 1, 2, False), (1, 1, 1, 1632, 272, 1, 0, 1, 2, False)]   # one row: gemv.cu reads the shifted filter


@pytest.mark.parametrize("case", TAIL_CASES, ids=lambda c: "x".join(map(str, c)))
def test_conv2d_k_tail_shift(device, case):
    """EQXV_FLAG_K_TAIL_SHIFT: the last K chunk fetched from channels [cin - 64, cin) (every TMA box in bounds) with the
    filter packed to match: same result as the plain layout (to fp32 summation order) and as torch"""
    from eqxvision_b200 import _pack, ops

    n, h, w, cin, cout, k, pad, dil, act, res = case
    x = rb(device, n, h, w, cin, seed=1)
    wt = rb(device, cout, cin, k, k, scale=(k * k * cin) ** -0.5, seed=2).float().cpu()      # OIHW
    bias = torch.randn(cout, generator=torch.Generator().manual_seed(3)).to(device)
    ho, wo = ops.conv_out_size(h, k, 1, pad, dil), ops.conv_out_size(w, k, 1, pad, dil)
    r = rb(device, n, ho, wo, cout, seed=4) if res else None
    kw = dict(cin=cin, cout=cout, kh=k, kw=k, stride=1, pad=pad, dil=dil, act=act, residual=r)
    plain = ops.conv2d(x, _pack.pack_conv_weight(wt, cin).to(device), bias, **kw)
    ws = _pack.pack_conv_weight(wt, cin, tail_shift=True)
    assert ws.shape == (cout, k * k * 64 * -(-cin // 64))
    assert torch.equal(_pack.unshift_tail(ws, k * k, cin), _pack.pack_conv_weight(wt, cin))
    y = ops.conv2d(x, ws.to(device), bias, k_tail_shift=True, **kw)
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wt.to(device), bias, padding=pad, dilation=dil)
    ref = ACTS[act](ref + (r.float().permute(0, 3, 1, 2) if res else 0)).permute(0, 2, 3, 1)
    assert rel_l2(y, ref) < TOL_BF16 and rel_l2(y, plain) < 1e-3


@pytest.mark.parametrize("k", [144, 336, 1632])
def test_gemm_gated_k_tail_shift(device, k):
    from eqxvision_b200 import _pack, ops

    n_img, hw, n = 3, 196, 48
    a = rb(device, n_img * hw, k, seed=1)
    gate = torch.rand(n_img, k, generator=torch.Generator().manual_seed(2)).to(torch.bfloat16).to(device)
    wt = rb(device, n, k, scale=k ** -0.5, seed=3).float().cpu()
    bias = torch.randn(n, generator=torch.Generator().manual_seed(4)).to(device)
    plain = ops.gemm_gated(a, gate, wt.to(torch.bfloat16).to(device), bias, rows_per_image=hw)
    ws = _pack.pack_conv_weight(wt.reshape(n, k, 1, 1), k, tail_shift=True).to(device)
    y = ops.gemm_gated(a, gate, ws, bias, rows_per_image=hw, k_tail_shift=True)
    ag = (a.float() * gate.float().repeat_interleave(hw, 0)).to(torch.bfloat16).float()
    assert rel_l2(y, ag @ wt.to(device).t() + bias) < TOL_BF16 and rel_l2(y, plain) < 1e-3
    with pytest.raises(Exception):
        ops.conv2d(rb(device, 1, 8, 8, 64), rb(device, 32, 64), bias[:32], cin=64, cout=32, kh=1, kw=1, k_tail_shift=True)


@pytest.mark.parametrize("n,h,w", [(3, 224, 224), (2, 64, 64), (5, 96, 128), (37, 224, 224)])
def test_stem_with_fused_maxpool(device, n, h, w):
    """eqxv_conv_stem_maxpool_bf16 (resnet.py:243-253): must equal the two separate entries BITWISE - the convolution is
    the same MMA sequence and max is exact, whatever the order the border contributions arrive in (red.global.max)."""
    from eqxvision_b200 import _pack, ops

    g = torch.Generator().manual_seed(0)
    x = torch.rand(n, 3, h, w, generator=g).to(device)
    wt = torch.randn(64, 3, 7, 7, generator=g) * 0.1
    bias = torch.randn(64, generator=g).to(device)
    xpad = ops.pack_stem_input(x)
    wp = _pack.pack_stem_weight(wt).to(device)
    y = ops.conv_stem(xpad, wp, bias, n=n, h=h, w=w, cout=64)
    ref = ops.maxpool2d(y, k=3, stride=2, pad=1)
    for _ in range(2):   # twice: the entry re-zeroes the border pixels itself
        got = ops.conv_stem_maxpool(xpad, wp, bias, n=n, h=h, w=w, cout=64)
        assert got.shape == ref.shape and torch.equal(got, ref)
    with pytest.raises(Exception):   # 56x56 conv output does not tile into 16-row blocks
        ops.conv_stem_maxpool(ops.pack_stem_input(x[:, :, :112, :112].contiguous()), wp, bias, n=n, h=112, w=112, cout=64)


# n, cin, h, w, cout, k, pad, act
STEM_C4_CASES = [(3, 3, 224, 224, 64, 7, 3, 1), (2, 3, 380, 380, 48, 3, 1, 2), (5, 3, 97, 130, 64, 7, 3, 1), (2, 1, 64, 64, 32, 5, 2, 0),
                 (2, 4, 96, 64, 24, 3, 1, 4), (37, 3, 224, 224, 64, 7, 3, 1)]


@pytest.mark.parametrize("case", STEM_C4_CASES, ids=lambda c: "x".join(map(str, c)))
def test_stem_pixel_pair_layout(device, case):
    """eqxv_conv_stem_c4_bf16 (stride-2 first layers, <= 4 channels; resnet.py:243-251, efficientnet.py:327-337): pairs of
    4-channel pixels per 16-byte unit, half the K steps of the 8-channel layout; against torch and against that layout."""
    from eqxvision_b200 import _pack, ops

    n, cin, h, w, cout, k, pad, act = case
    g = torch.Generator().manual_seed(0)
    x = torch.rand(n, cin, h, w, generator=g).to(device)
    wt = torch.randn(cout, cin, k, k, generator=g) * (cin * k * k) ** -0.5
    bias = torch.randn(cout, generator=g).to(device)
    y8 = ops.conv_stem(ops.pack_stem_input(x, pad=pad), _pack.pack_stem_weight(wt).to(device), bias, n=n, h=h, w=w, cout=cout,
                       kh=k, kw=k, stride=2, pad=pad, act=act)
    x4 = ops.pack_stem_input_c4(x, pad=pad)
    assert x4.shape == (n, h + 2 * pad, (w + 8) // 2, 8)
    y4 = ops.conv_stem(x4, _pack.pack_stem_weight_c4(wt).to(device), bias, n=n, h=h, w=w, cout=cout, kh=k, kw=k, stride=2, pad=pad,
                       act=act, c4=True)
    ref = F.conv2d(x.to(torch.bfloat16).float(), wt.to(torch.bfloat16).float().to(device), bias, stride=2, padding=pad)
    ref = ACTS[act](ref).permute(0, 2, 3, 1)
    assert rel_l2(y4, ref) < TOL_BF16 and rel_l2(y4, y8) < 1e-3
    with pytest.raises(Exception):   # stride 1 has no pixel-pair variant
        ops.conv_stem(x4, _pack.pack_stem_weight_c4(wt).to(device), bias, n=n, h=h, w=w, cout=cout, kh=k, kw=k, stride=1, pad=pad,
                      act=act, c4=True)
