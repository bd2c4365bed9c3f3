"""The uint8 input edge: Resize -> ToTensor -> Normalize of the reference's own fixture
(/root/reference/tests/conftest.py:20-41) on the device, fused into the first layer's layout kernels.

Bar: the uint8 -> float step is BIT-exact against torchvision's transforms; a model fed with pixels returns the same
bits as the same model fed with the fp32 tensor the host pipeline produces from those pixels."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _pixels(n, h, w, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (n, h, w, 3), generator=g, dtype=torch.uint8)


# ------------------------------------------------------------------------------------------------ CPU (host logic)
def test_lut_is_bit_identical_to_torchvision_totensor_normalize():
    from torchvision.transforms import transforms

    from eqxvision_b200 import transforms as tr

    mean, std = tr.IMAGENET_MEAN, tr.IMAGENET_STD
    lut = tr.normalize_lut(mean, std)
    assert lut.shape == (3, 256) and lut.dtype == torch.float32
    # every uint8 value in every channel through the reference fixture's transforms (conftest.py:26-27)
    img = np.stack([np.arange(256, dtype=np.uint8)] * 3, -1).reshape(16, 16, 3)
    ref = transforms.Compose([transforms.ToTensor(), transforms.Normalize(mean, std)])(img)   # [3,16,16] fp32
    assert torch.equal(lut, ref.reshape(3, 256))
    other = tr.normalize_lut((0.5, 0.25, 0.1), (0.5, 0.3, 0.7))
    ref2 = transforms.Compose([transforms.ToTensor(), transforms.Normalize((0.5, 0.25, 0.1), (0.5, 0.3, 0.7))])(img)
    assert torch.equal(other, ref2.reshape(3, 256))
    with pytest.raises(ValueError):
        tr.normalize_lut((0.5,), (0.0,))


def test_images_u8_surface():
    from eqxvision_b200 import transforms as tr

    px = _pixels(2, 40, 48)
    b = tr.images_u8(px)
    assert b.shape == (2, 3, 40, 48)
    assert tr.images_u8(px.numpy(), size=32).shape == (2, 3, 32, 32)
    ref = b.reference_pipeline()
    from torchvision.transforms import transforms

    t = transforms.Compose([transforms.ToTensor(), transforms.Normalize(tr.IMAGENET_MEAN, tr.IMAGENET_STD)])
    assert torch.equal(ref[1], t(px[1].numpy()))
    with pytest.raises(TypeError):
        tr.images_u8(px.float())
    with pytest.raises(ValueError):
        tr.images_u8(px, mean=(0.5,), std=(0.5,))


@pytest.mark.parametrize("arch,hw", [("resnet18", 64), ("alexnet", 224), ("vit_tiny", 64)])
def test_u8_plan_lowers_to_the_same_result_as_the_fp32_plan(arch, hw, tmp_path):
    """CPU replay of the launch plan (tests/plan_interpreter.py): the uint8 plan differs from the fp32 plan only in its
    first layout step and must give the same bits (stem pack, generic NHWC and patch-row variants)."""
    import eqxvision_b200 as eb
    from eqxvision_b200 import transforms as tr
    import plan_interpreter as PI
    from tools import synthetic as syn

    if arch == "vit_tiny":
        sd = syn.vit_state_dict(embed_dim=192, depth=2, heads=3, num_classes=10, seed=3, img=hw)
        p = tmp_path / "v.pth"
        torch.save(sd, p)
        net = eb.models.vit_tiny(img_size=hw, depth=2, num_classes=10, torch_weights=str(p))
    else:
        sd = syn.torchvision_state_dict(arch, seed=1)
        p = tmp_path / "m.pth"
        torch.save(sd, p)
        net = getattr(eb.models, arch)(torch_weights=str(p))
    net = eb.tree_inference(net, True)
    b = tr.images_u8(_pixels(2, hw, hw, seed=5))
    got_u8, plan = PI.run(net, b)
    got_f32, _ = PI.run(net, b.reference_pipeline())
    assert torch.equal(got_u8, got_f32)
    names = [fn.__name__ for fn, _ in plan.steps]
    assert names[0] in ("u8_pack_stem_input", "u8_pack_stem_input_c4", "u8_to_nhwc", "u8_patchify")
    assert not any(n in ("pack_stem_input", "nchw_to_nhwc", "patchify", "u8_to_nchw_f32") for n in names)


def test_u8_plan_with_resize_step():
    from eqxvision_b200 import _engine as E

    plan = E.Plan(torch.device("cpu"), 2, (3, 32, 32), u8={"mean": (0.5,) * 3, "std": (0.5,) * 3, "raw_hw": (48, 40)})
    assert plan.steps[0][0].__name__ == "u8_resize_bilinear"
    assert tuple(plan.x_host_target.shape) == (2, 48, 40, 3) and tuple(plan.x_in.shape) == (2, 32, 32, 3)


# ------------------------------------------------------------------------------------------------ B200
@pytest.mark.gpu
@pytest.mark.parametrize("n,h,w", [(3, 224, 224), (2, 37, 53), (1, 512, 512)])
def test_u8_kernels_bit_exact(device, n, h, w):
    from eqxvision_b200 import ops
    from eqxvision_b200 import transforms as tr

    b = tr.images_u8(_pixels(n, h, w, seed=n))
    ref = b.reference_pipeline()                       # torchvision's ToTensor + Normalize on the host
    px, lut = b.pixels.to(device), tr.normalize_lut().to(device)
    torch.cuda.synchronize()
    got = ops.u8_to_nchw_f32(px, lut)
    torch.cuda.synchronize()
    assert torch.equal(got.cpu(), ref)                 # bit for bit
    ref_d = ref.to(device)
    for pad in (1, 3):
        a, bb = ops.u8_pack_stem_input(px, lut, pad=pad), ops.pack_stem_input(ref_d, pad=pad)
        torch.cuda.synchronize()
        assert torch.equal(a, bb)
        if w % 2 == 0:   # pixel-pair layout: the same pixels, 4 channels each, two per 16-byte unit
            a4, b4 = ops.u8_pack_stem_input_c4(px, lut, pad=pad), ops.pack_stem_input_c4(ref_d, pad=pad)
            torch.cuda.synchronize()
            assert torch.equal(a4, b4)
            assert torch.equal(a4.reshape(n, h + 2 * pad, w + 8, 4), bb[..., :4]) and (bb[..., 4:] == 0).all()
    a, bb = ops.u8_to_nhwc(px, lut), ops.nchw_to_nhwc(ref_d, c_pad=8)
    torch.cuda.synchronize()
    assert torch.equal(a, bb)
    if h % 16 == 0 and w % 16 == 0:
        a, bb = ops.u8_patchify(px, lut, 16), ops.patchify(ref_d, 16)
        torch.cuda.synchronize()
        assert torch.equal(a, bb)


@pytest.mark.gpu
@pytest.mark.parametrize("h,w,oh,ow", [(256, 320, 224, 224), (100, 75, 224, 224), (224, 224, 112, 96)])
def test_u8_resize_vs_torch_bilinear(device, h, w, oh, ow):
    import torch.nn.functional as F

    from eqxvision_b200 import ops

    px = _pixels(2, h, w, seed=7)
    got = ops.u8_resize_bilinear(px.to(device), oh, ow)
    torch.cuda.synchronize()
    ref = F.interpolate(px.permute(0, 3, 1, 2).float(), size=(oh, ow), mode="bilinear", align_corners=False)
    ref = ref.round().clamp_(0, 255).permute(0, 2, 3, 1)
    d = (got.cpu().float() - ref).abs()
    # same formula in fp32; an FMA-contraction difference on the host can move a value across a rounding tie
    assert d.max().item() <= 1 and (d == 0).float().mean().item() > 0.999


@pytest.mark.gpu
@pytest.mark.parametrize("arch,hw", [("resnet18", 224), ("alexnet", 224), ("vit_tiny", 224), ("efficientnet_b0", 224)])
def test_model_on_pixels_equals_model_on_host_pipeline_bitwise(device, save_checkpoint, arch, hw):
    import eqxvision_b200 as eb
    from eqxvision_b200 import transforms as tr
    from tools import synthetic as syn

    if arch == "vit_tiny":
        sd = syn.vit_state_dict(embed_dim=192, depth=2, heads=3, num_classes=10, seed=3)
        net = eb.models.vit_tiny(depth=2, num_classes=10, torch_weights=save_checkpoint(sd))
    else:
        net = getattr(eb.models, arch)(torch_weights=save_checkpoint(syn.torchvision_state_dict(arch, seed=1)))
    net = eb.tree_inference(net, True)
    fwd = eb.filter_jit(eb.vmap(net, axis_name="batch"))
    b = tr.images_u8(_pixels(5, hw, hw, seed=11))
    keys = eb.random.split(eb.random.PRNGKey(0), 5)
    got = fwd(b, key=keys)
    ref = fwd(b.reference_pipeline(), key=keys)
    assert torch.equal(got, ref)
    # with the Resize step of the fixture in front
    big = tr.images_u8(_pixels(5, 256, 288, seed=12), size=hw)
    assert torch.equal(fwd(big, key=keys), fwd(big.reference_pipeline(), key=keys)) or \
        ((fwd(big, key=keys) - fwd(big.reference_pipeline(), key=keys)).abs().max() < 0.05)   # resize ties, see above
    assert torch.equal(tr.to_model_input(b).cpu(), b.reference_pipeline())


@pytest.mark.gpu
def test_async_pipelined_calls_return_their_own_results(device, save_checkpoint):
    """the public call is asynchronous and rotates over EQXV_LANES plan instances for host-resident inputs: results of
    back-to-back calls must not alias or overwrite each other (pinned, pageable, device-resident and uint8 inputs)"""
    import eqxvision_b200 as eb
    from eqxvision_b200 import transforms as tr
    from tools import synthetic as syn

    net = eb.models.resnet18(torch_weights=save_checkpoint(syn.torchvision_state_dict("resnet18", seed=1)))
    net = eb.tree_inference(net, True)
    fwd = eb.filter_jit(eb.vmap(net, axis_name="batch"))
    keys = eb.random.split(eb.random.PRNGKey(0), 8)
    xs = [syn.synthetic_images(8, h=96, w=96, seed=s) for s in range(7)]
    want = []
    for x in xs:
        want.append(fwd(x, key=keys).cpu())                               # one at a time
    pinned = [x.pin_memory() for x in xs]
    outs = [fwd(x, key=keys) for x in pinned]                             # seven calls in flight
    eb.block_until_ready(outs)
    for o, w_ in zip(outs, want):
        assert torch.equal(o.cpu(), w_)
    outs = [fwd(x.to(device), key=keys) for x in xs]                      # device-resident inputs
    for o, w_ in zip(outs, want):
        assert torch.equal(o.cpu(), w_)
    px = [tr.images_u8(_pixels(8, 96, 96, seed=s)).pin_memory() for s in range(5)]
    want_u8 = [fwd(p.reference_pipeline(), key=keys).cpu() for p in px]
    outs = [fwd(p, key=keys) for p in px]
    for o, w_ in zip(outs, want_u8):
        assert torch.equal(o.cpu(), w_)
    assert len(net.__dict__["_eqxv_plans"]) <= 8
    from eqxvision_b200 import _engine

    _engine.clear_plans(net)
    assert len(net.__dict__["_eqxv_plans"]) == 0
