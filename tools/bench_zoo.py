#!/usr/bin/env python
"""Device-resident forward throughput of every model family of the zoo on one B200 (secondary evidence next to
bench.py's headline configs): seeded synthetic checkpoint -> load_torch_weights -> one CUDA-graph plan -> CUDA events
around K replays on the launching stream. Prints one JSON line per model and a table.

  python tools/bench_zoo.py [--batch 64] [--steps 20] [--models a,b,c]
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

DEFAULT = ["alexnet", "vgg11_bn", "resnet18", "resnet50", "resnext50_32x4d", "densenet121", "googlenet", "squeezenet1_1",
           "mobilenet_v2", "mobilenet_v3_large", "shufflenet_v2_x1_0", "efficientnet_b0", "efficientnet_v2_s",
           "regnet_y_400mf", "regnet_x_400mf", "convnext_tiny", "swin_t", "vit_small"]


def build(name):
    import torch

    import eqxvision_b200 as eb
    from tools import synthetic as syn

    kw = {}
    if name.startswith("vit_"):
        dim, heads = {"vit_tiny": (192, 3), "vit_small": (384, 6), "vit_base": (768, 12)}[name]
        sd = syn.vit_state_dict(embed_dim=dim, depth=12, heads=heads, num_classes=1000, seed=3)
        kw = {"num_classes": 1000}
    elif name.startswith("swin"):
        sd = syn.swin_model(name, seed=1).state_dict()
    elif name == "googlenet":
        sd = syn.torchvision_state_dict(name, seed=1, aux_logits=True, transform_input=False, init_weights=False)
    else:
        sd = syn.torchvision_state_dict(name, seed=1)
    f = tempfile.NamedTemporaryFile(suffix=".pth", delete=False)
    torch.save(sd, f.name)
    model = getattr(eb.models, name)(torch_weights=f.name, **kw)
    os.unlink(f.name)
    return eb.tree_inference(model, True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--models", default=",".join(DEFAULT))
    args = ap.parse_args()

    import torch

    import eqxvision_b200 as eb
    from eqxvision_b200 import _engine, _lib

    rows = []
    for name in args.models.split(","):
        try:
            model = build(name)
            plan = _engine.get_plan(model, "__call__", args.batch, (3, 224, 224), (), {"key": eb.random.PRNGKey(0)})
            st = _engine.stream_handle()
            plan.x_in.copy_(torch.rand(plan.x_in.shape, generator=torch.Generator().manual_seed(0)))
            torch.cuda.synchronize()
            for _ in range(3):
                plan.launch(st)
            _lib.call("eqxv_stream_sync", st)
            e0, e1 = C.c_void_p(), C.c_void_p()
            _lib.call("eqxv_event_create", C.byref(e0))
            _lib.call("eqxv_event_create", C.byref(e1))
            _lib.call("eqxv_event_record", e0, st)
            for _ in range(args.steps):
                plan.launch(st)
            _lib.call("eqxv_event_record", e1, st)
            _lib.call("eqxv_event_sync", e1)
            ms = C.c_float()
            _lib.call("eqxv_event_elapsed_ms", e0, e1, C.byref(ms))
            per = ms.value / args.steps
            row = {"model": name, "batch": args.batch, "ms_per_step": round(per, 4),
                   "img_per_s": round(args.batch / per * 1e3, 1), "launches": plan.num_launches,
                   "activation_gib": round(plan.act_bytes / 2 ** 30, 2)}
        except Exception as exc:  # noqa: BLE001 - keep going, report the failure in the table
            row = {"model": name, "error": f"{type(exc).__name__}: {exc}"[:200]}
        rows.append(row)
        print(json.dumps(row), flush=True)
        del model
        torch.cuda.empty_cache()
    print(f"\n{'model':24s} {'img/s':>10s} {'ms/step':>9s} {'launches':>9s}")
    for r in rows:
        if "error" in r:
            print(f"{r['model']:24s} ERROR {r['error']}")
        else:
            print(f"{r['model']:24s} {r['img_per_s']:10.1f} {r['ms_per_step']:9.3f} {r['launches']:9d}")


if __name__ == "__main__":
    main()
