"""Host-side surface: constructor names/signatures, error conventions and the
tracer's fusion decisions — all without a GPU."""
import inspect

import pytest
import torch

import eqxvision_b200 as eb
from eqxvision_b200 import _trace as T
from eqxvision_b200 import functional as F
from eqxvision_b200 import layers, models, nn

KEY = eb.random.PRNGKey(0)


def trace(module, shape, kind="chw", method="__call__", **kw):
    fn = getattr(type(module), method)
    fn = getattr(fn, "__wrapped__", fn)
    kw.setdefault("key", KEY)
    return fn(module, T.Sym(kind, shape, T.Input()), **kw)


def test_exported_names():
    for name in ["resnet18", "resnet50", "resnet101", "wide_resnet50_2", "resnext50_32x4d", "ResNet",
                 "vit_tiny", "vit_small", "vit_base", "VisionTransformer", "_VitAttention", "_VitBlock"]:
        assert hasattr(models, name), name
    for name in ["ConvNormActivation", "SqueezeExcitation", "PatchEmbed", "MlpProjection", "DropPath",
                 "LayerNorm2d", "Linear2d"]:
        assert hasattr(layers, name), name
    assert callable(eb.vmap) and callable(eb.filter_jit) and callable(eb.tree_inference)


def test_constructor_signatures_match_reference():
    p = inspect.signature(models.ResNet.__init__).parameters
    assert list(p)[1:] == ["block", "layers", "num_classes", "groups", "width_per_group",
                           "replace_stride_with_dilation", "norm_layer", "key"]
    assert p["num_classes"].default == 1000 and p["key"].kind is inspect.Parameter.KEYWORD_ONLY
    p = inspect.signature(models.VisionTransformer.__init__).parameters
    assert list(p)[1:] == ["img_size", "patch_size", "in_chans", "num_classes", "embed_dim", "depth", "num_heads",
                           "mlp_ratio", "qkv_bias", "qk_scale", "drop_rate", "attn_drop_rate", "drop_path_rate",
                           "norm_layer", "key"]
    assert p["num_classes"].default == 0 and p["qkv_bias"].default is True
    p = inspect.signature(models.vit_base).parameters
    assert [p[k].default for k in ("patch_size", "embed_dim", "depth", "num_heads", "mlp_ratio")] == [16, 768, 12, 12, 4]
    p = inspect.signature(layers.ConvNormActivation.__init__).parameters
    assert list(p)[1:] == ["in_channels", "out_channels", "kernel_size", "stride", "padding", "groups",
                           "norm_layer", "activation_layer", "dilation", "use_bias", "key"]


def test_resnet_requires_key_and_traces_to_fused_graph():
    net = eb.tree_inference(models.resnet50(), True)
    with pytest.raises(RuntimeError, match="PRNGKey"):
        trace(net, (3, 224, 224), key=None)
    out = trace(net, (3, 224, 224))
    assert out.kind == "vec" and out.shape == (1000,)
    e = out.expr
    assert isinstance(e, T.Linear) and isinstance(e.x.expr, T.Ravel)
    last = e.x.expr.x.expr.x.expr           # ravel <- avgpool <- last bottleneck output
    assert isinstance(last, T.Conv) and last.bn is not None and last.res is not None and last.act2 == "relu"
    assert isinstance(last.x.expr, T.Conv) and last.x.expr.act1 == "relu" and last.x.expr.weight.shape[-1] == 3

    def count(expr, seen):
        if id(expr) in seen or not isinstance(expr, T.Expr):
            return 0
        seen.add(id(expr))
        n = 1 if isinstance(expr, T.Conv) else 0
        for s in getattr(expr, "__slots__", ()):
            v = getattr(expr, s)
            for u in (v if isinstance(v, (tuple, list)) else (v,)):
                if isinstance(u, T.Sym):
                    n += count(u.expr, seen)
        return n

    assert count(e, set()) == 53  # every BatchNorm / ReLU / residual add folded into its conv


def test_resnet_errors():
    with pytest.raises(ValueError):
        models.ResNet(models.classification.resnet._ResNetBottleneck, [1, 1, 1, 1],
                      replace_stride_with_dilation=[True])
    with pytest.raises(NotImplementedError):
        models.ResNet(models.classification.resnet._ResNetBottleneck, [1, 1, 1, 1], norm_layer=nn.LayerNorm)
    with pytest.raises(NotImplementedError, match="tree_inference"):
        trace(models.resnet18(), (3, 64, 64))  # training-mode BatchNorm is out of scope: loud, not silent


def test_dilated_resnet_keeps_resolution():
    net = eb.tree_inference(models.resnet50(replace_stride_with_dilation=[False, True, True]), True)
    x = T.Sym("chw", (3, 128, 128), T.Input())
    y = net.maxpool(net.relu(net.bn1(net.conv1(x))))
    y = net.layer3(net.layer2(net.layer1(y, key=KEY), key=KEY), key=KEY)
    assert y.shape == (1024, 16, 16)
    assert net.layer3.layers[0].conv2.dilation == (1, 1) and net.layer3.layers[1].conv2.dilation == (2, 2)
    assert net.layer4.layers[0].conv2.dilation == (2, 2) and net.layer4.layers[2].conv2.dilation == (4, 4)


def test_patch_embed_shape_and_size_check():
    pe = layers.PatchEmbed(224, 16, 3, 768, flatten=True, key=KEY)
    out = trace(pe, (3, 224, 224))
    assert out.kind == "tokens" and out.shape == (196, 768)       # reference test_layers.py:17
    with pytest.raises(ValueError):
        trace(pe, (3, 225, 224))


def test_vit_shapes_and_modes():
    for fn, dim in [(models.vit_tiny, 192), (models.vit_small, 384), (models.vit_base, 768)]:
        net = fn(num_classes=1000)
        assert trace(net, (3, 224, 224)).shape == (1000,)          # reference test_vit.py:101
    assert trace(models.vit_small(num_classes=0), (3, 224, 224)).shape == (384,)  # test_vit.py:113
    net = models.VisionTransformer(img_size=224, patch_size=16)
    with pytest.raises(ValueError):
        trace(net, (3, 224, 224), method="get_last_self_attention")  # test_vit.py:75-76
    net = eb.tree_inference(net, True)
    attn = trace(net, (3, 224, 224), method="get_last_self_attention")
    assert attn.shape == (1, 12, 197, 197)                          # test_vit.py:66
    out = trace(models.vit_base(num_classes=10), (3, 224, 224))
    head = out.expr
    assert isinstance(head, T.Linear) and isinstance(head.x.expr, T.LayerNormE)
    assert isinstance(head.x.expr.x.expr, T.SelectRow)              # final LayerNorm only on the CLS row


def test_vit_attention_and_block_submodules():
    att = models._VitAttention(32, num_heads=4, qkv_bias=True, key=KEY)
    out, probs = trace(att, (8, 32), kind="tokens")
    assert out.shape == (8, 32) and probs.shape == (1, 4, 8, 8)    # reference test_vit.py:19-29
    blk = models._VitBlock(32, num_heads=4, key=KEY)
    assert trace(blk, (8, 32), kind="tokens").shape == (8, 32)
    assert trace(blk, (8, 32), kind="tokens", return_attention=True).shape == (1, 4, 8, 8)


def test_mlp_and_droppath_and_dropout():
    mlp = layers.MlpProjection(16, 32, 8, act_layer=F.gelu, key=KEY)
    out = trace(mlp, (5, 16), kind="tokens")
    assert out.shape == (5, 8) and out.expr.x.expr.act1 == "gelu"
    dp = layers.DropPath(p=0.5)
    x = T.Sym("tokens", (4, 8), T.Input())
    with pytest.raises(RuntimeError):
        dp(x, key=None)
    assert eb.tree_inference(dp, True)(x, key=None) is x           # drop_path.py:44-45
    assert layers.DropPath(p=0.0)(x, key=None) is x
    with pytest.raises(NotImplementedError):
        nn.Dropout(0.5)(x, key=KEY)
    assert nn.Dropout(0.5, inference=True)(x) is x


def test_conv_norm_activation_and_se_trace():
    cna = eb.tree_inference(layers.ConvNormActivation(8, 16, 3, key=KEY), True)
    y = trace(cna, (8, 10, 10))
    assert y.shape == (16, 10, 10) and y.expr.bn is not None and y.expr.act1 == "relu"
    assert cna.layers[0].bias is None                               # bias only without a norm layer
    assert layers.ConvNormActivation(8, 16, 3, norm_layer=None, key=KEY).layers[0].bias is not None
    se = layers.SqueezeExcitation(16, 4, key=KEY)
    z = trace(se, (16, 6, 6))
    assert z.shape == (16, 6, 6) and isinstance(z.expr, T.ChannelScale)
    assert z.expr.s.expr.act1 == "sigmoid" and z.expr.s.expr.x.expr.act1 == "relu"


def test_tree_inference_is_functional():
    net = models.resnet18()
    inf = eb.tree_inference(net, True)
    assert inf is not net and inf.bn1.inference and not net.bn1.inference
    assert inf.conv1.weight is net.conv1.weight                     # arrays are shared, flags are not


def test_vmap_rejects_unsupported_axes_and_needs_cuda():
    net = eb.tree_inference(models.resnet18(), True)
    with pytest.raises(NotImplementedError):
        eb.vmap(net, in_axes=1)
    with pytest.raises(TypeError):
        eb.vmap(lambda x: x)
    if not torch.cuda.is_available():
        with pytest.raises(Exception, match="CUDA|fallback"):
            eb.vmap(net, axis_name="batch")(torch.zeros(1, 3, 64, 64), key=[KEY])


def test_swin_traces_to_fused_blocks_and_loads_positionally(tmp_path):
    """swin.py: field order == torchvision state_dict order (incl. the integer relative_position_index
    buffer); the (C,H,W)<->(HW,C) transposes of Linear2d/LayerNorm2d vanish; residual adds and the GELU
    fold into the GEMM epilogues"""
    from oracle import checkpoints as ck

    for name in ["SwinTransformer", "swin_t", "swin_s", "swin_b", "swin_v2_t", "swin_v2_s", "swin_v2_b"]:
        assert hasattr(models, name), name
    p = inspect.signature(models.SwinTransformer.__init__).parameters
    assert list(p)[1:] == ["patch_size", "embed_dim", "depths", "num_heads", "window_size", "mlp_ratio", "dropout",
                           "attention_dropout", "stochastic_depth_prob", "num_classes", "norm_layer", "block",
                           "downsample_layer", "key"]
    tv = ck.swin_model("swin_t", seed=3)
    path = tmp_path / "swin_t.pth"
    torch.save(tv.state_dict(), path)
    with pytest.warns(UserWarning, match="dynamic padding"):
        net = models.swin_t(torch_weights=str(path))
    blk = net.features[1][1]
    assert blk.attn.shift_size == [3, 3] and blk.attn.relative_position_index.dtype == torch.int64
    assert torch.equal(blk.attn.relative_position_index, tv.features[1][1].attn.relative_position_index)
    assert torch.equal(blk.attn.qkv.weight, tv.features[1][1].attn.qkv.weight.detach())
    assert torch.equal(net.features[2].reduction.weight, tv.features[2].reduction.weight.detach())
    assert torch.equal(net.head.bias, tv.head.bias.detach())
    # un-loaded model: the reference's index quirk (sum of relative coordinates, partly negative)
    with pytest.warns(UserWarning):
        raw = models.swin_t()
    idx = raw.features[1][0].attn.relative_position_index
    assert idx.shape == (49 * 49,) and idx.min() == -12 and idx.max() == 12
    assert raw.features[1][0].attn.get_relative_position_bias().shape == (3, 49, 49)

    out = trace(eb.tree_inference(net, True), (3, 224, 224))
    assert out.kind == "vec" and out.shape == (1000,)
    # walk back: head <- ravel <- avgpool <- ToMap(LayerNorm(tokens))
    pooled = out.expr.x.expr.x.expr
    assert isinstance(pooled, T.AdaptiveAvgPool)
    ln = pooled.x.expr.x.expr
    assert isinstance(ln, T.LayerNormE)
    fc2 = ln.x.expr                      # last block: x + mlp(norm2(x)) folded into fc2's epilogue
    assert isinstance(fc2, T.Linear) and fc2.res is not None
    fc1 = fc2.x.expr
    assert isinstance(fc1, T.Linear) and fc1.act1 == "gelu"
    proj = fc2.res.expr                  # x = x + attn(...) folded into proj's epilogue
    assert isinstance(proj, T.Linear) and proj.res is not None
    wa = proj.x.expr
    assert isinstance(wa, T.WindowAttention) and wa.window == (7, 7) and wa.shift == (0, 0) and wa.heads == 24
    with pytest.raises(ValueError, match="multiple of the window"):
        trace(eb.tree_inference(net, True), (3, 200, 200))


# ---- the rest of the zoo: constructors, error conventions, and the launch plan each model lowers to
def lower(net, shape=(3, 224, 224), batch=2, **kw):
    """host-side dry run of the engine: trace + lower into C-ABI steps on a CPU-device plan, nothing is executed"""
    from eqxvision_b200 import _engine as E

    plan = E.Plan(torch.device("cpu"), batch, shape)
    out = trace(net, shape, **kw)
    syms = []
    plan.out_struct = E._flatten_out(out, syms)
    for s in syms:
        plan.add_output(s)
    names = [fn.__name__ for fn, _ in plan.steps]
    return out, plan, {n: names.count(n) for n in set(names)}


def test_zoo_constructors_exported_with_reference_signatures():
    for name in ["alexnet", "AlexNet", "mobilenet_v2", "MobileNetV2", "squeezenet1_0", "squeezenet1_1", "SqueezeNet",
                 "googlenet", "GoogLeNet", "RegNet", "regnet_y_400mf", "regnet_y_128gf", "regnet_x_400mf",
                 "regnet_x_32gf", "ShuffleNetV2", "shufflenet_v2_x0_5", "shufflenet_v2_x1_0", "shufflenet_v2_x1_5",
                 "shufflenet_v2_x2_0", "ConvNeXt", "convnext_tiny", "convnext_small", "convnext_base", "convnext_large"]:
        assert hasattr(models, name), name
    p = inspect.signature(models.AlexNet.__init__).parameters
    assert list(p)[1:] == ["num_classes", "dropout", "key"] and p["dropout"].default == 0.5
    p = inspect.signature(models.MobileNetV2.__init__).parameters
    assert list(p)[1:] == ["num_classes", "width_mult", "inverted_residual_setting", "round_nearest", "block",
                           "norm_layer", "dropout", "key"]
    p = inspect.signature(models.SqueezeNet.__init__).parameters
    assert list(p)[1:] == ["version", "num_classes", "dropout", "key"] and p["version"].default == "1_0"
    p = inspect.signature(models.GoogLeNet.__init__).parameters
    assert list(p)[1:] == ["num_classes", "aux_logits", "blocks", "dropout", "dropout_aux", "key"]
    p = inspect.signature(models.RegNet.__init__).parameters
    assert list(p)[1:] == ["block_params", "num_classes", "stem_width", "stem_type", "block_type", "norm_layer",
                           "activation", "key"]


def test_every_name_the_reference_exports_from_models_exists():
    """eqxvision/models/__init__.py of the reference, name by name (the Python surface is the boundary)"""
    names = """AlexNet alexnet ConvNeXt convnext_base convnext_large convnext_small convnext_tiny DenseNet densenet121
    densenet161 densenet169 densenet201 EfficientNet efficientnet_b0 efficientnet_b1 efficientnet_b2 efficientnet_b3
    efficientnet_b4 efficientnet_b5 efficientnet_b6 efficientnet_b7 efficientnet_v2_l efficientnet_v2_m
    efficientnet_v2_s GoogLeNet googlenet mobilenet_v2 MobileNetV2 mobilenet_v3_large mobilenet_v3_small MobileNetV3
    RegNet regnet_x_1_6gf regnet_x_3_2gf regnet_x_8gf regnet_x_16gf regnet_x_32gf regnet_x_400mf regnet_x_800mf
    regnet_y_1_6gf regnet_y_3_2gf regnet_y_8gf regnet_y_16gf regnet_y_32gf regnet_y_128gf regnet_y_400mf
    regnet_y_800mf ResNet resnet18 resnet34 resnet50 resnet101 resnet152 resnext50_32x4d resnext101_32x8d
    wide_resnet50_2 wide_resnet101_2 shufflenet_v2_x0_5 shufflenet_v2_x1_0 shufflenet_v2_x1_5 shufflenet_v2_x2_0
    ShuffleNetV2 SqueezeNet squeezenet1_0 squeezenet1_1 swin_b swin_s swin_t swin_v2_b swin_v2_s swin_v2_t
    SwinTransformer VGG vgg11 vgg11_bn vgg13 vgg13_bn vgg16 vgg16_bn vgg19 vgg19_bn _VitAttention _VitBlock
    VisionTransformer vit_base vit_small vit_tiny DeepLabV3 deeplabv3 FCN fcn LRASPP lraspp_mobilenet_v3_large""".split()
    missing = [n for n in names if not hasattr(models, n)]
    assert not missing, missing
    from eqxvision_b200.models import classification, segmentation

    for sub in ("alexnet convnext densenet efficientnet googlenet mobilenetv2 mobilenetv3 regnet resnet shufflenetv2 "
                "squeezenet swin vgg vit").split():
        assert hasattr(classification, sub), sub
    for sub in ("deeplabv3", "fcn", "lraspp"):
        assert hasattr(segmentation, sub), sub
    for n in ("experimental", "layers", "models", "utils"):
        assert hasattr(eb, n), n


def test_alexnet_lowers_to_twelve_launches_and_requires_key():
    net = eb.tree_inference(models.alexnet(), True)
    with pytest.raises(RuntimeError, match="PRNGKey"):          # alexnet.py:78-79
        trace(net, (3, 224, 224), key=None)
    out, plan, steps = lower(net)
    assert out.shape == (1000,)
    # 11x11/4 first layer on the generic implicit GEMM; ReLU and bias fused; 6x6 adaptive pool is the identity at 224
    assert steps == {"nchw_to_nhwc": 1, "conv2d": 5, "maxpool2d": 3, "gemm": 3}
    assert [f.shape for f in [trace(net.features, (3, 224, 224))]] == [(256, 6, 6)]   # test_alexnet.py:23 reaches in


def test_mobilenet_v2_residual_folds_without_recomputing_the_producer():
    net = eb.tree_inference(models.mobilenet_v2(), True)
    out, plan, steps = lower(net)
    assert out.shape == (1000,)
    # 17 depthwise + 34 1x1 convs + stem + pool + head: `x + conv(x)` (mobilenetv2.py:86) must fold into the block's
    # own projection conv, not into a second copy of the conv that produced x
    assert steps["dwconv"] == 17 and steps["conv2d"] == 34 and steps["conv_stem"] == 1
    with pytest.raises(ValueError):
        models.MobileNetV2(inverted_residual_setting=[])


def test_regnet_block_params_and_lowering():
    from eqxvision_b200.models.classification.regnet import BlockParams

    bp = BlockParams.from_init_params(depth=16, w_0=48, w_a=27.89, w_m=2.09, group_width=8, se_ratio=0.25)
    assert (bp.widths, bp.depths, bp.group_widths) == ([48, 104, 208, 440], [1, 3, 6, 6], [8, 8, 8, 8])
    bp = BlockParams.from_init_params(depth=23, w_0=80, w_a=49.56, w_m=2.88, group_width=120)
    assert (bp.widths, bp.depths, bp.group_widths) == ([80, 240, 720, 1920], [2, 5, 15, 1], [80, 120, 120, 120])
    with pytest.raises(ValueError):
        BlockParams.from_init_params(depth=4, w_0=50, w_a=1.0, w_m=2.0, group_width=8)   # w_0 % 8 != 0
    out, plan, steps = lower(eb.tree_inference(models.regnet_y_400mf(), True))
    assert out.shape == (1000,)
    # 3 convs/block, 4 proj, SE fc1/fc2; in the first block of each stage the SE gate rides in the conv that follows
    # it (no activation on that conv: the shortcut add sits in the projection's epilogue) -> eqxv_gemm_gated_bf16
    assert steps["conv2d"] == 16 * 3 + 4 + 16 * 2 - 4 and steps["gemm_gated"] == 4 and steps["eltwise"] == 12


def test_grouped_weight_dense_expansion():
    from eqxvision_b200 import _pack

    g = torch.Generator().manual_seed(0)
    w = torch.randn(24, 4, 3, 3, generator=g)                  # 6 groups of width 4 (RegNet-style geometry)
    x = torch.randn(1, 24, 9, 9, generator=g)
    dense = _pack.expand_grouped_weight(w, 6)
    assert dense.shape == (24, 24, 3, 3)
    ref = torch.nn.functional.conv2d(x, w, padding=1, groups=6)
    assert torch.allclose(torch.nn.functional.conv2d(x, dense, padding=1), ref, atol=1e-5)


def test_squeezenet_and_googlenet_lowering():
    out, plan, steps = lower(eb.tree_inference(models.squeezenet1_0(), True))
    assert out.shape == (1000,) and steps["maxpool2d"] == 3 and steps["conv2d"] == 8 * 3 + 1
    assert sum(1 for fn, kw in plan.steps if fn.__name__ == "maxpool2d" and kw["ceil_mode"]) == 3
    with pytest.raises(ValueError):
        models.SqueezeNet("2_0")
    net = eb.tree_inference(models.googlenet(), True)
    with pytest.raises(RuntimeError, match="PRNGKey"):          # googlenet.py:112-113
        trace(net, (3, 224, 224), key=None)
    out, plan, steps = lower(net)
    assert out.shape == (1000,)
    assert steps["conv2d"] == 2 + 9 * 6 and steps["maxpool2d"] == 4 + 9     # no concat pass: branches store in place
    assert "copy2d" not in steps and "eltwise" not in steps


def test_convnext_layer_scale_folds_into_the_second_gemm():
    net = eb.tree_inference(models.convnext_tiny(), True)
    out, plan, steps = lower(net)
    assert out.shape == (1000,)
    # 18 blocks x (depthwise 7x7, LayerNorm, 2 GEMMs) + stem + 3 x (LayerNorm, conv2x2/2) + pool + LayerNorm + head:
    # layer_scale and the residual add are not separate passes
    assert steps == {"pack_stem_input": 1, "conv_stem": 1, "dwconv": 18, "layernorm": 18 + 1 + 3 + 1,
                     "gemm": 36 + 1, "conv2d": 3, "adaptive_avgpool": 1}
    blk = net.features.layers[1].layers[0]
    assert blk.block.layers[1].eps == 1e-5 and net.features.layers[0].layers[1].eps == 1e-6   # convnext.py:24 vs :120
    y = trace(blk, (96, 8, 8))
    lin = y.expr.x.expr                                                    # ToMap(Linear)
    assert isinstance(lin, T.Linear) and lin.res is not None
    assert torch.allclose(lin.weight, blk.block.layers[4].weight * blk.layer_scale.reshape(-1, 1))
    with pytest.raises(ValueError):
        models.ConvNeXt([])
    with pytest.raises(TypeError):
        models.ConvNeXt([(96, 192, 3)])


def test_unsupported_shapes_fail_at_trace_time_not_at_launch():
    out = trace(eb.tree_inference(models.googlenet(aux_logits=True), True), (3, 224, 224))   # aux heads: 14x14 -> 4x4
    assert isinstance(out, tuple) and len(out) == 3 and all(o.shape == (1000,) for o in out)
    with pytest.raises(NotImplementedError, match="larger than the map"):   # AlexNet at 127 px: 3x3 -> 6x6
        trace(eb.tree_inference(models.alexnet(), True), (3, 127, 127))
    with pytest.raises(ValueError, match="multiple of the window"):     # Swin-V2: window 8 does not tile 224/4 = 56... 7
        trace(eb.tree_inference(models.swin_v2_t(), True), (3, 224, 224))
    out = trace(eb.tree_inference(models.swin_v2_t(), True), (3, 256, 256))
    assert out.shape == (1000,)
    with pytest.raises(NotImplementedError):
        nn.AvgPool2d(3, 2, use_ceil=True)


def test_googlenet_loads_torchvision_checkpoint_with_aux_heads(tmp_path):
    """googlenet.py:320-335: the checkpoint carries aux1/aux2, the model is built with them, then aux_logits is cleared"""
    import torchvision

    torch.manual_seed(0)
    tv = torchvision.models.googlenet(weights=None, aux_logits=True, transform_input=False, init_weights=False)
    path = str(tmp_path / "g.pth")
    torch.save(tv.state_dict(), path)
    net = models.googlenet(torch_weights=path)
    assert net.aux_logits is False and net.aux1 is not None
    assert torch.equal(net.fc.weight, tv.fc.weight) and torch.equal(net.aux2.fc2.weight, tv.aux2.fc2.weight)
    assert torch.equal(net.inception5b.branch4.layers[1].conv.weight, tv.inception5b.branch4[1].conv.weight)


def test_wide_layers_split_into_column_chunks_at_lowering():
    """the GEMM stages the shift of all its output channels in 20 KB of shared memory (csrc/igemm.cu), so
    one launch covers at most ~5000 of them. ConvNeXt-Large's MLP (4 x 1536 = 6144, convnext.py:40-48) and the widest
    RegNet stage must therefore lower to several launches over column slices - and give the same numbers."""
    import plan_interpreter as PI
    from eqxvision_b200 import _engine as E
    from eqxvision_b200.models.classification.convnext import _CNBlockConfig

    assert E._n_chunks(4096) == [(0, 4096)] and E._n_chunks(6144) == [(0, 3072), (3072, 6144)]
    assert E._n_chunks(7392) == [(0, 3696), (3696, 7392)] and all(a % 8 == 0 for a, _ in E._n_chunks(7392))
    # one ConvNeXt-Large stage-4 block (dim 1536 -> MLP width 6144) on a 2x2 map
    net = eb.tree_inference(models.ConvNeXt([_CNBlockConfig(1536, None, 2)], num_classes=16), True)
    x = torch.randn(1, 3, 8, 8, generator=torch.Generator().manual_seed(0))
    got, plan = PI.run(net, x, fp32_activations=True)
    wide = [kw for fn, kw in plan.steps if fn.__name__ == "gemm" and kw["wgt"].shape[0] == 3072]
    assert len(wide) == 4 and all(kw["wgt"].shape[0] <= E.MAX_COUT_PER_LAUNCH
                                  for fn, kw in plan.steps if fn.__name__ in ("gemm", "conv2d"))
    saved = E.MAX_COUT_PER_LAUNCH
    try:
        E.MAX_COUT_PER_LAUNCH = 1 << 20
        ref, plan1 = PI.run(net, x, fp32_activations=True)
    finally:
        E.MAX_COUT_PER_LAUNCH = saved
    assert len(plan1.steps) == len(plan.steps) - 2
    assert torch.equal(got, ref)
