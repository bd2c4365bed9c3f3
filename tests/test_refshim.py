"""The reference's OWN code against the oracle: /root/reference/eqxvision is imported unmodified on top of the
stand-ins of oracle/refshim (jax / equinox are not installable here), its constructors build the models, its
`load_torch_weights` loads the seeded checkpoints and its `__call__` under `jax.vmap(net, axis_name="batch")` produces
the outputs the hand restatement in oracle/models.py must reproduce. Skipped where /root/reference does not exist
(the GPU box); tests/golden/golden_ref_v1.pt carries outputs generated the same way to every box."""
import os

import numpy as np
import pytest
import torch

from oracle import checkpoints as ck
from oracle import models as om
from oracle import refshim

needs_reference = pytest.mark.skipif(not refshim.available(), reason="reference sources not present on this box")
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_ref_v1.pt")


def run_reference(build, x, path, method=None):
    """build(eqxvision, path) -> model; returns the batched output(s) of the reference code as torch tensors"""
    with refshim.install() as ev:
        import equinox as eqx
        import jax

        net = eqx.tree_inference(build(ev, path), True)
        fn = net if method is None else getattr(net, method)
        keys = jax.random.split(jax.random.PRNGKey(0), x.shape[0])
        out = jax.vmap(fn, axis_name="batch")(jax.numpy.asarray(x.numpy()), key=keys)

    def conv(o):
        if o is None:
            return None
        if isinstance(o, (tuple, list)):
            return type(o)(conv(i) for i in o)
        return torch.from_numpy(np.ascontiguousarray(np.asarray(o, dtype=np.float32)))

    return conv(out)


CASES = [  # (reference constructor, oracle function, input hw, torchvision kwargs)
    ("alexnet", "alexnet", 224, {}),
    ("resnet18", "resnet", 64, {}),
    ("resnet50", "resnet", 64, {}),
    ("resnext50_32x4d", "resnet", 64, {}),
    ("vgg11", "vgg", 224, {}),
    ("vgg11_bn", "vgg", 224, {}),
    ("densenet121", "densenet", 64, {}),
    ("mobilenet_v2", "mobilenet_v2", 64, {}),
    ("mobilenet_v3_small", "mobilenet_v3", 64, {}),
    ("mobilenet_v3_large", "mobilenet_v3", 64, {}),
    ("efficientnet_b0", "efficientnet", 64, {}),
    ("efficientnet_b4", "efficientnet", 64, {}),            # BASELINE.json configs[3]
    ("efficientnet_v2_s", "efficientnet", 64, {}),
    ("wide_resnet50_2", "resnet", 64, {}),
    ("regnet_y_400mf", "regnet", 64, {}),
    ("regnet_x_400mf", "regnet", 64, {}),
    ("squeezenet1_0", "squeezenet", 96, {}),
    ("googlenet", "googlenet", 96, {"aux_logits": True, "transform_input": False, "init_weights": True}),
    ("shufflenet_v2_x1_0", "shufflenet_v2", 64, {}),
]


@needs_reference
@pytest.mark.parametrize("arch,fn,hw,kw", CASES, ids=[c[0] for c in CASES])
def test_oracle_reproduces_the_reference_code(tmp_path, arch, fn, hw, kw):
    sd = ck.torchvision_state_dict(arch, seed=1, calib_hw=min(hw, 96), **kw)
    path = str(tmp_path / "w.pth")
    torch.save(sd, path)
    x = ck.synthetic_images(2, h=hw, w=hw, seed=2)
    got = run_reference(lambda ev, p: getattr(ev.models, arch)(torch_weights=p), x, path)
    ref = getattr(om, fn)(sd, x, arch)
    assert got.shape == ref.shape
    assert torch.allclose(got, ref, atol=1e-4, rtol=1e-4), (got - ref).abs().max()


@needs_reference
def test_oracle_reproduces_the_reference_vit(tmp_path):
    """ViT (vit.py): shape-only tests in the reference itself; pinned here by executing vit.py. Also the default
    `num_classes=0` (CLS feature out) and `get_last_self_attention` (vit.py:275-292)."""
    cfg = dict(embed_dim=192, depth=3, heads=3, num_classes=10)
    sd = ck.vit_state_dict(seed=3, **cfg)
    path = str(tmp_path / "v.pth")
    torch.save(sd, path)
    x = ck.synthetic_images(2, seed=2)
    build = lambda ev, p: ev.models.vit_tiny(depth=3, num_classes=10, torch_weights=p)  # noqa: E731
    got = run_reference(build, x, path)
    ref = om.vit(sd, x, heads=3)
    assert torch.allclose(got, ref, atol=1e-4, rtol=1e-4), (got - ref).abs().max()
    attn = run_reference(build, x, path, method="get_last_self_attention")
    ref_attn = om.vit(sd, x, heads=3, return_last_attention=True)
    assert attn.numel() == ref_attn.numel()
    assert (attn.reshape(ref_attn.shape) - ref_attn).abs().max() < 1e-5


@needs_reference
def test_oracle_reproduces_the_reference_vit_base(tmp_path):
    """BASELINE.json configs[2]: the full ViT-B/16 (12 blocks, 12 heads, 768 wide, 1000 classes), one image"""
    sd = ck.vit_state_dict(embed_dim=768, depth=12, heads=12, num_classes=1000, seed=3)
    path = str(tmp_path / "vb.pth")
    torch.save(sd, path)
    x = ck.synthetic_images(1, seed=2)
    got = run_reference(lambda ev, p: ev.models.vit_base(num_classes=1000, torch_weights=p), x, path)
    ref = om.vit(sd, x, heads=12)
    assert got.shape == ref.shape == (1, 1000)
    assert torch.allclose(got, ref, atol=1e-4, rtol=1e-4), (got - ref).abs().max()


@needs_reference
def test_oracle_reproduces_the_reference_convnext(tmp_path):
    sd = ck.torchvision_state_dict("convnext_tiny", seed=1)
    path = str(tmp_path / "c.pth")
    torch.save(sd, path)
    x = ck.synthetic_images(1, h=64, w=64, seed=2)
    got = run_reference(lambda ev, p: _convnext_from_path(ev, p), x, path)
    ref = om.convnext(sd, x, "convnext_tiny")
    assert torch.allclose(got, ref, atol=1e-4, rtol=1e-4), (got - ref).abs().max()


def _convnext_from_path(ev, path):
    # convnext.py:218-221 ignores `torch_weights` and always fetches CLASSIFICATION_URLS[arch]: build, then load
    net = ev.models.convnext_tiny()
    return ev.utils.load_torch_weights(net, torch_weights=path)


@needs_reference
def test_oracle_reproduces_the_reference_swin(tmp_path):
    sd = ck.swin_model("swin_t", seed=1).state_dict()
    path = str(tmp_path / "s.pth")
    torch.save(sd, path)
    x = ck.synthetic_images(1, seed=2)
    got = run_reference(lambda ev, p: ev.models.swin_t(torch_weights=p), x, path)
    ref = om.swin(sd, x, "swin_t")
    assert torch.allclose(got, ref, atol=1e-4, rtol=1e-4), (got - ref).abs().max()


@needs_reference
def test_oracle_reproduces_the_reference_swin_v2(tmp_path):
    """Swin-V2 through the reference's own swin.py (swin.py:369-522, 583-636, 874-896): cosine attention normalised
    along axis 0, the cpb_mlp reshape, post-norm blocks, _PatchMergingV2; torchvision swin_v2_t checkpoint loaded
    positionally (logit_scale, relative_coords_table, relative_position_index, qkv, proj, cpb_mlp)."""
    sd = ck.swin_model("swin_v2_t", seed=1).state_dict()
    path = str(tmp_path / "s2.pth")
    torch.save(sd, path)
    x = ck.synthetic_images(1, h=256, w=256, seed=2)
    got = run_reference(lambda ev, p: ev.models.swin_v2_t(torch_weights=p), x, path)
    ref = om.swin_v2(sd, x, "swin_v2_t")
    assert got.shape == (1, 1000)
    # In the last stage the 8x8 map is ONE window, so the reference's axis-0 norm runs over a single element and
    # q / ||q|| is sign(q): discontinuous, which turns fp32 summation-order noise (numpy vs torch matmuls) into
    # differences of a few 1e-4. The two-stage model below, where every norm runs over >= 4 windows, holds 1e-4.
    assert torch.allclose(got, ref, atol=2e-3, rtol=2e-3), (got - ref).abs().max()
    assert ((got - ref).norm() / ref.norm()).item() < 1e-3
    # and it is NOT torchvision's Swin-V2 (the quirks are real): the stock model disagrees on the same checkpoint
    tv = ck.swin_model("swin_v2_t", seed=1)
    with torch.no_grad():
        assert (tv(x) - ref).abs().max() > 1e-2


@needs_reference
def test_oracle_reproduces_the_reference_swin_v2_two_stages_at_1e_4(tmp_path):
    """the same code path with >= 4 windows per image in every stage (no sign() degeneracy): the reference's own
    tolerance atol=1e-4"""
    from torchvision.models.swin_transformer import PatchMergingV2, SwinTransformerBlockV2

    cfg = dict(patch_size=[4, 4], embed_dim=96, depths=[2, 2], num_heads=[3, 6], window_size=[8, 8],
               stochastic_depth_prob=0.0, num_classes=10, block=SwinTransformerBlockV2, downsample_layer=PatchMergingV2)
    sd = ck.swin_model(cfg, seed=2).state_dict()
    path = str(tmp_path / "s2s.pth")
    torch.save(sd, path)
    x = ck.synthetic_images(2, h=128, w=128, seed=3)

    def build(ev, p):
        sw = ev.models.classification.swin
        net = ev.models.SwinTransformer(patch_size=[4, 4], embed_dim=96, depths=[2, 2], num_heads=[3, 6],
                                        window_size=[8, 8], stochastic_depth_prob=0.0, num_classes=10,
                                        block=sw._SwinTransformerBlockV2, downsample_layer=sw._PatchMergingV2)
        return ev.utils.load_torch_weights(net, torch_weights=p)

    got = run_reference(build, x, path)
    ref = om.swin_v2(sd, x, (96, [2, 2], [3, 6], 8))
    assert got.shape == (2, 10)
    assert torch.allclose(got, ref, atol=1e-4, rtol=1e-4), (got - ref).abs().max()


@needs_reference
def test_oracle_reproduces_the_reference_segmentation(tmp_path):
    x = ck.synthetic_images(1, h=64, w=64, seed=2)
    sd = ck.torchvision_model("deeplabv3_resnet50", seed=1, calib_hw=64, aux_loss=True).state_dict()
    path = str(tmp_path / "d.pth")
    torch.save(sd, path)
    aux, out = run_reference(lambda ev, p: ev.models.deeplabv3(
        intermediate_layers=lambda m: [m.layer3, m.layer4], aux_in_channels=1024, torch_weights=p), x, path)
    aux_r, out_r = om.deeplabv3_resnet50(sd, x)                       # (aux, out) order: _utils.py:58
    assert torch.allclose(out, out_r, atol=1e-4, rtol=1e-4) and torch.allclose(aux, aux_r, atol=1e-4, rtol=1e-4)
    sd = ck.torchvision_model("lraspp_mobilenet_v3_large", seed=1, calib_hw=64).state_dict()
    torch.save(sd, path)
    none, out = run_reference(lambda ev, p: ev.models.lraspp_mobilenet_v3_large(torch_weights=p), x, path)
    assert none is None
    assert torch.allclose(out, om.lraspp_mobilenet_v3_large(sd, x), atol=1e-4, rtol=1e-4)


@needs_reference
@pytest.mark.parametrize("arch,kw", [("resnet50", {}), ("efficientnet_b0", {}), ("densenet121", {}),
                                     ("mobilenet_v3_large", {}), ("regnet_y_400mf", {}), ("shufflenet_v2_x1_0", {}),
                                     ("googlenet", {"aux_logits": True, "transform_input": False, "init_weights": False})])
def test_positional_loader_places_every_tensor_where_the_reference_does(tmp_path, arch, kw):
    """utils.py:171-218: array leaves are replaced in pytree order, BatchNorm statistics in StateIndex order. The
    reference's loader (run through the shim) and eqxvision_b200.utils.load_torch_weights must put the same tensor on
    the same leaf, leaf by leaf."""
    import eqxvision_b200 as eb
    from eqxvision_b200 import nn as bnn

    sd = ck.torchvision_state_dict(arch, seed=1, **kw)
    path = str(tmp_path / "w.pth")
    torch.save(sd, path)
    with refshim.install() as ev:
        import equinox as eqx
        import jax.tree_util as jtu

        ref_net = getattr(ev.models, arch)(torch_weights=path)
        ref_arrays = [np.asarray(leaf) for leaf in jtu.tree_leaves(ref_net) if isinstance(leaf, np.ndarray)]
        ref_stats = [leaf._state for leaf in jtu.tree_leaves(ref_net)
                     if isinstance(leaf, eqx.experimental.StateIndex)]
    ours = getattr(eb.models, arch)(torch_weights=path)
    our_arrays = [leaf for leaf in bnn.tree_leaves(ours) if isinstance(leaf, torch.Tensor)]
    our_stats = [leaf.value for leaf in bnn.tree_leaves(ours) if isinstance(leaf, bnn.StateIndex)]
    assert len(ref_arrays) == len(our_arrays) > 0
    for i, (r, o) in enumerate(zip(ref_arrays, our_arrays)):
        assert tuple(r.shape) == tuple(o.shape), (i, r.shape, o.shape)
        assert np.array_equal(r, o.numpy()), i
    assert len(ref_stats) == len(our_stats)
    for r, o in zip(ref_stats, our_stats):
        if isinstance(o, tuple):
            assert np.array_equal(np.asarray(r[0]), o[0].numpy()) and np.array_equal(np.asarray(r[1]), o[1].numpy())
        else:
            assert bool(np.asarray(r)) is False and o is False


def _raises(fn):
    try:
        fn()
    except Exception as exc:  # noqa: BLE001 - the point is to learn which exception type comes out
        return type(exc)
    return None


@needs_reference
def test_error_conventions_match_the_reference_code():
    """SURVEY.md 8(b) error conventions, derived by provoking the reference's own code (through the shim) and this
    repo's host code with the same misuse: both must raise, and raise the same exception type."""
    import eqxvision_b200 as eb
    from eqxvision_b200 import _trace as T

    sym = T.Sym("chw", (3, 64, 64), T.Input())
    ours = {
        "resnet_without_key": lambda: type(eb.models.resnet18()).__call__.__wrapped__(
            eb.tree_inference(eb.models.resnet18(), True), sym, key=None),
        "alexnet_without_key": lambda: type(eb.models.alexnet()).__call__.__wrapped__(
            eb.tree_inference(eb.models.alexnet(), True), T.Sym("chw", (3, 224, 224), T.Input()), key=None),
        "bad_replace_stride_with_dilation": lambda: eb.models.resnet50(replace_stride_with_dilation=[True]),
        "empty_efficientnet_setting": lambda: eb.models.EfficientNet([], 0.2),
        "empty_mobilenet_v2_setting": lambda: eb.models.MobileNetV2(inverted_residual_setting=[]),
        "empty_convnext_setting": lambda: eb.models.ConvNeXt([]),
        "shufflenet_wrong_stage_count": lambda: eb.models.ShuffleNetV2([4, 8], [24, 48, 96, 192, 1024]),
        "invalid_regnet_width": lambda: eb.models.regnet.BlockParams.from_init_params(4, 50, 1.0, 2.0, 8),
        "load_without_path": lambda: eb.utils.load_torch_weights(eb.models.resnet18(), None),
        "patch_embed_size_mismatch": lambda: type(eb.layers.PatchEmbed(224, 16, 3, 32)).__call__.__wrapped__(
            eb.layers.PatchEmbed(224, 16, 3, 32), T.Sym("chw", (3, 200, 224), T.Input())),
    }
    with refshim.install() as ev:
        import equinox as eqx
        import jax

        img = jax.numpy.zeros((3, 64, 64))
        theirs = {
            "resnet_without_key": lambda: eqx.tree_inference(ev.models.resnet18(), True)(img, key=None),
            "alexnet_without_key": lambda: eqx.tree_inference(ev.models.alexnet(), True)(
                jax.numpy.zeros((3, 224, 224)), key=None),
            "bad_replace_stride_with_dilation": lambda: ev.models.resnet50(replace_stride_with_dilation=[True]),
            "empty_efficientnet_setting": lambda: ev.models.EfficientNet([], 0.2),
            "empty_mobilenet_v2_setting": lambda: ev.models.MobileNetV2(inverted_residual_setting=[]),
            "empty_convnext_setting": lambda: ev.models.ConvNeXt([]),
            "shufflenet_wrong_stage_count": lambda: ev.models.ShuffleNetV2([4, 8], [24, 48, 96, 192, 1024]),
            "invalid_regnet_width": lambda: ev.models.regnet.BlockParams.from_init_params(4, 50, 1.0, 2.0, 8),
            "load_without_path": lambda: ev.utils.load_torch_weights(ev.models.resnet18(), None),
            "patch_embed_size_mismatch": lambda: ev.layers.PatchEmbed(224, 16, 3, 32)(jax.numpy.zeros((3, 200, 224))),
        }
        expected = {k: _raises(f) for k, f in theirs.items()}
    for k, f in ours.items():
        assert expected[k] is not None, f"the reference does not raise on {k}"
        assert _raises(f) is expected[k], (k, _raises(f), expected[k])


@needs_reference
def test_oracle_reproduces_the_reference_fcn(tmp_path):
    x = ck.synthetic_images(1, h=64, w=64, seed=2)
    sd = ck.torchvision_model("fcn_resnet50", seed=1, calib_hw=64, aux_loss=True).state_dict()
    path = str(tmp_path / "f.pth")
    torch.save(sd, path)
    aux, out = run_reference(lambda ev, p: ev.models.fcn(
        intermediate_layers=lambda m: [m.layer3, m.layer4], aux_in_channels=1024, torch_weights=p), x, path)
    aux_r, out_r = om.fcn_resnet50(sd, x)
    assert torch.allclose(out, out_r, atol=1e-4, rtol=1e-4) and torch.allclose(aux, aux_r, atol=1e-4, rtol=1e-4)


@needs_reference
def test_oracle_reproduces_the_reference_googlenet_aux_heads(tmp_path):
    """aux_logits=True: (logits, aux2, aux1) (googlenet.py:174-175); the heads pool 14x14 -> 4x4 with EQUINOX's uneven
    adaptive rule (2 blocks of 4, then 2 of 3), which differs from torch's overlapping windows. CPU oracle only."""
    import torchvision

    torch.manual_seed(1)
    tv = torchvision.models.googlenet(weights=None, aux_logits=True, transform_input=False, init_weights=True)
    ck._perturb_and_calibrate(tv, 1, (4, 3, 96, 96))
    sd = tv.state_dict()
    path = str(tmp_path / "g.pth")
    torch.save(sd, path)
    x = ck.synthetic_images(1, seed=2)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        # googlenet(torch_weights=..., aux_logits=True) is a TypeError in the reference (googlenet.py:323-325 passes
        # aux_logits twice): build the class and load explicitly
        got = run_reference(lambda ev, p: ev.utils.load_torch_weights(ev.models.GoogLeNet(aux_logits=True), p), x, path)
    ref = om.googlenet(sd, x, aux_logits=True)
    assert len(got) == len(ref) == 3
    for g, r in zip(got, ref):
        assert torch.allclose(g, r, atol=1e-4, rtol=1e-4), (g - r).abs().max()
    with torch.no_grad():
        tv.train(False)
        tv_aux = tv.aux1(tv.inception4a(tv.maxpool3(tv.inception3b(tv.inception3a(tv.maxpool2(tv.conv3(tv.conv2(
            tv.maxpool1(tv.conv1(x))))))))))
    assert not torch.allclose(ref[2], tv_aux, atol=1e-3)       # torch's adaptive rule gives different numbers
