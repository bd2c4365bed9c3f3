"""Generates tests/golden/golden_ref_v1.pt: outputs of the REFERENCE'S OWN CODE (/root/reference/eqxvision, executed
through oracle/refshim because jax / equinox cannot be installed here) on seeded checkpoints and images.

Run in the build container (needs /root/reference):  python tests/golden/make_golden_ref.py
Only seeds and outputs are stored; checkpoints and images are regenerated from the seeds by tools/synthetic.py (CPU
generators, identical on every box), so the fixture travels to the GPU box where /root/reference does not exist.
Configurations marked gpu=True are the ones tests/test_gpu_zoo.py::test_reference_golden_vectors replays on the B200.
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import checkpoints as ck  # noqa: E402
from test_refshim import run_reference  # noqa: E402

# (key, constructor, torchvision kwargs, checkpoint seed, image seed, batch, hw, replayed on the GPU, tol vs fp32)
CASES = [
    ("alexnet_224", "alexnet", {}, 1, 2, 3, 224, True, 3e-2),
    ("resnet18_224", "resnet18", {}, 1, 2, 4, 224, True, 3e-2),
    ("resnet50_64", "resnet50", {}, 21, 22, 2, 64, False, None),
    ("mobilenet_v2_224", "mobilenet_v2", {}, 1, 2, 2, 224, True, 3e-2),
    ("efficientnet_b0_224", "efficientnet_b0", {}, 1, 2, 4, 224, True, 2e-2),
    ("densenet121_64", "densenet121", {}, 21, 22, 2, 64, False, None),
    ("mobilenet_v3_small_64", "mobilenet_v3_small", {}, 21, 22, 2, 64, False, None),
    ("regnet_y_400mf_224", "regnet_y_400mf", {}, 1, 2, 2, 224, True, 4e-2),
    ("squeezenet1_1_224", "squeezenet1_1", {}, 1, 2, 2, 224, True, 4e-2),
    ("googlenet_224", "googlenet", {"aux_logits": True, "transform_input": False, "init_weights": True}, 1, 2, 2, 224,
     True, 4e-2),
    ("shufflenet_v2_x0_5_64", "shufflenet_v2_x0_5", {}, 21, 22, 2, 64, False, None),
]


def main():
    out = {}
    tmp = os.path.join(os.environ.get("TMPDIR", "/tmp"), "golden_ref_ckpt.pth")
    for key, arch, kw, seed, img_seed, n, hw, gpu, tol in CASES:
        sd = ck.torchvision_state_dict(arch, seed=seed, **kw)
        torch.save(sd, tmp)
        x = ck.synthetic_images(n, h=hw, w=hw, seed=img_seed)
        y = run_reference(lambda ev, p: getattr(ev.models, arch)(torch_weights=p), x, tmp)
        out[key] = dict(arch=arch, tv_kwargs=kw, seed=seed, img_seed=img_seed, n=n, hw=hw, gpu=gpu, tol=tol,
                        expected=y.clone())
        print(key, tuple(y.shape), float(y.abs().max()))
    cfg = dict(embed_dim=192, depth=3, heads=3, num_classes=10)
    sd = ck.vit_state_dict(seed=23, **cfg)
    torch.save(sd, tmp)
    x = ck.synthetic_images(2, seed=24)
    y = run_reference(lambda ev, p: ev.models.vit_tiny(depth=3, num_classes=10, torch_weights=p), x, tmp)
    out["vit_tiny_3blk"] = dict(arch="vit_tiny", ctor_kw=dict(depth=3, num_classes=10), cfg=cfg, seed=23, img_seed=24,
                                n=2, hw=224, gpu=True, tol=3e-2, expected=y.clone())
    sd = ck.torchvision_state_dict("convnext_tiny", seed=1)
    torch.save(sd, tmp)
    x = ck.synthetic_images(2, seed=2)

    def build(ev, p):   # convnext.py:218-221 ignores the path it is given
        return ev.utils.load_torch_weights(ev.models.convnext_tiny(), torch_weights=p)

    y = run_reference(build, x, tmp)
    out["convnext_tiny_224"] = dict(arch="convnext_tiny", tv_kwargs={}, seed=1, img_seed=2, n=2, hw=224, gpu=True,
                                    tol=4e-2, expected=y.clone())
    os.unlink(tmp)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_ref_v1.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
