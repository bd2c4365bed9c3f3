"""End-to-end bring-up check on a B200: model constructors + positional loader + plan builder + CUDA
graph vs the CPU oracle, then a quick throughput read-out.
Usage: python tools/e2e_check.py [resnet18|resnet50|vit_base|vit_tiny ...]
"""
from __future__ import annotations

import os
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import eqxvision_b200 as eb  # noqa: E402
from eqxvision_b200 import _engine, _lib  # noqa: E402
from oracle import checkpoints as ck  # noqa: E402
from oracle import models as om  # noqa: E402


def metrics(name, got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    rel = ((got - ref).norm() / ref.norm()).item()
    mx = (got - ref).abs().max().item()
    cos = torch.nn.functional.cosine_similarity(got.flatten(1), ref.flatten(1)).min().item()
    top1 = (got.flatten(1).argmax(1) == ref.flatten(1).argmax(1)).float().mean().item()
    print(f"[{name}] rel_l2={rel:.3e} max_abs={mx:.3e} ref_absmax={ref.abs().max():.3e} min_cos={cos:.6f} "
          f"top1_agree={top1:.3f}", flush=True)
    return rel


def save_sd(sd):
    f = tempfile.NamedTemporaryFile(suffix=".pth", delete=False)
    torch.save(sd, f.name)
    return f.name


def bench(model, batch, shape=(3, 224, 224), iters=20):
    plan = _engine.get_plan(model, "__call__", batch, shape, (), {"key": eb.random.PRNGKey(0)})
    st = _engine.stream_handle()
    x = torch.rand((batch,) + shape)
    plan.x_in.copy_(x)
    torch.cuda.synchronize()
    for _ in range(3):
        plan.launch(st)
    _lib.call("eqxv_stream_sync", st)
    t0 = time.perf_counter()
    for _ in range(iters):
        plan.launch(st)
    _lib.call("eqxv_stream_sync", st)
    dt = (time.perf_counter() - t0) / iters
    print(f"   batch {batch}: {dt * 1e3:.3f} ms/step -> {batch / dt:.0f} img/s  ({plan.num_launches} launches, "
          f"{plan.act_bytes / 2**20:.0f} MiB activations)", flush=True)


def run(name):
    key = eb.random.PRNGKey(0)
    if name.startswith("resnet"):
        sd = ck.torchvision_state_dict(name, seed=1)
        model = getattr(eb.models, name)(torch_weights=save_sd(sd))
        model = eb.tree_inference(model, True)
        x = ck.synthetic_images(8, seed=2)
        ref = om.resnet(sd, x, name)
        got = eb.vmap(model, axis_name="batch")(x, key=eb.random.split(key, 8))
        metrics(name, got, ref)
        bench(model, 256)
    elif name.startswith("vit"):
        cfg = {"vit_tiny": (192, 3), "vit_small": (384, 6), "vit_base": (768, 12)}[name]
        sd = ck.vit_state_dict(embed_dim=cfg[0], heads=cfg[1], num_classes=1000, seed=3)
        model = getattr(eb.models, name)(num_classes=1000, torch_weights=save_sd(sd))
        model = eb.tree_inference(model, True)
        x = ck.synthetic_images(4, seed=4)
        ref = om.vit(sd, x, heads=cfg[1])
        got = eb.vmap(model)(x, key=eb.random.split(key, 4))
        metrics(name, got, ref)
        bench(model, 64)
    elif name == "deeplabv3":
        tv = ck.torchvision_model("deeplabv3_resnet50", seed=1, calib_hw=64, aux_loss=True)
        sd = tv.state_dict()
        model = eb.models.deeplabv3(intermediate_layers=lambda m: [m.layer3, m.layer4], aux_in_channels=1024,
                                    torch_weights=save_sd(sd))
        model = eb.tree_inference(model, True)
        x = ck.synthetic_images(2, h=128, w=128, seed=2)
        aux_r, out_r = om.deeplabv3_resnet50(sd, x)
        aux, out = eb.vmap(model, axis_name="batch")(x, key=eb.random.split(key, 2))
        metrics(name + " out", out, out_r)
        metrics(name + " aux", aux, aux_r)
        bench(model, 4, (3, 512, 512), iters=5)
    else:
        fam = {"efficientnet": om.efficientnet, "mobilenet": om.mobilenet_v3, "vgg": om.vgg, "densenet": om.densenet}
        fn = next(v for k, v in fam.items() if name.startswith(k))
        sd = ck.torchvision_state_dict(name, seed=1)
        model = getattr(eb.models, name)(torch_weights=save_sd(sd))
        model = eb.tree_inference(model, True)
        x = ck.synthetic_images(4, seed=2)
        ref = fn(sd, x, name)
        from oracle import ops as O
        with O.emulate_bf16():
            emu = fn(sd, x, name)
        got = eb.vmap(model, axis_name="batch")(x, key=eb.random.split(key, 4))
        metrics(name + " vs fp32", got, ref)
        metrics(name + " vs emu ", got, emu)
        bench(model, 128)


if __name__ == "__main__":
    for n in (sys.argv[1:] or ["resnet18", "resnet50", "vit_base"]):
        t0 = time.time()
        try:
            run(n)
        except Exception as ex:  # noqa: BLE001
            import traceback
            traceback.print_exc()
            print(f"[{n}] FAILED: {type(ex).__name__}: {ex}", flush=True)
        print(f"   ({n}: {time.time() - t0:.1f}s)", flush=True)
