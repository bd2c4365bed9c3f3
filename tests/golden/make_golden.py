"""Generates tests/golden/golden_v1.pt: expected fp32 outputs for seeded checkpoints / images.

Run in the build container:  python tests/golden/make_golden.py
Source of the expected values: the CPU oracle (oracle/models.py), which is itself pinned against
torchvision (tests/test_oracle.py) and against the reference's own model files executed through
oracle/refshim (tests/test_refshim.py; golden_ref_v1.pt holds outputs generated that way). Only seeds and outputs are stored; weights and images are
regenerated from the seeds by oracle/checkpoints.py (CPU generator => identical on every box).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import checkpoints as ck  # noqa: E402
from oracle import models as om  # noqa: E402


def main():
    out = {}
    for arch, hw, n, tol in [("resnet18", 64, 4, 3e-2), ("resnet50", 224, 2, 6e-2)]:
        sd = ck.torchvision_state_dict(arch, seed=11)
        x = ck.synthetic_images(n, h=hw, w=hw, seed=12)
        out[f"{arch}_{hw}"] = dict(family="resnet", arch=arch, seed=11, img_seed=12, n=n, hw=hw, tol=tol,
                                   expected=om.resnet(sd, x, arch))
    cfg = dict(embed_dim=192, depth=4, heads=3, num_classes=10)
    sd = ck.vit_state_dict(seed=13, **cfg)
    x = ck.synthetic_images(2, seed=14)
    out["vit_tiny_4blk"] = dict(family="vit", ctor="vit_tiny", ctor_kw=dict(depth=4, num_classes=10), cfg=cfg,
                                seed=13, img_seed=14, n=2, tol=3e-2, expected=om.vit(sd, x, heads=3))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_v1.pt")
    torch.save(out, path)
    print("wrote", path, {k: tuple(v["expected"].shape) for k, v in out.items()})


if __name__ == "__main__":
    main()
