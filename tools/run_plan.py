"""Replay one model's launch plan eagerly a few times (ncu target: per-launch metrics of ONE step).
usage: python tools/run_plan.py <resnet50|vit_base|efficientnet_b4|deeplabv3_resnet50|alexnet> [passes]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (model table + synthetic checkpoints; nothing is timed here)

name = sys.argv[1]
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 3
import torch  # noqa: E402

import eqxvision_b200 as eb  # noqa: E402
from eqxvision_b200 import _engine, _lib  # noqa: E402

torch.cuda.set_device(0)
model, _ = bench.build_model(name)
batch, hw = bench.PER_GPU_BATCH[name], bench.MODELS[name]["hw"]
plan = _engine.get_plan(model, "__call__", batch, (3, hw, hw), (), {"key": eb.random.split(eb.random.PRNGKey(0), batch)})
plan.x_in.copy_(torch.rand(plan.x_in.shape))
torch.cuda.synchronize()
for _ in range(passes):
    plan.run_steps(plan.stream)
    _lib.call("eqxv_stream_sync", plan.stream)
print(name, "launches per step:", plan.num_launches, "first step:", plan.steps[0][0].__name__)
