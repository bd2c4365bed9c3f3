"""Per-launch device time of one model plan (eager replay with CUDA events).
Usage: python tools/model_breakdown.py <ctor> <batch> [hw] [topN]"""
import ctypes as C, os, sys, tempfile
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import eqxvision_b200 as eb
from eqxvision_b200 import _engine, _lib

name, batch = sys.argv[1], int(sys.argv[2])
hw = int(sys.argv[3]) if len(sys.argv) > 3 else 224
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 25
if name == "deeplabv3":
    model = eb.models.deeplabv3(intermediate_layers=lambda m: [m.layer3, m.layer4], aux_in_channels=1024)
elif name.startswith("vit"):
    model = getattr(eb.models, name)(num_classes=1000)
else:
    model = getattr(eb.models, name)()
model = eb.tree_inference(model, True)
plan = _engine.get_plan(model, "__call__", batch, (3, hw, hw), (), {"key": eb.random.PRNGKey(0)})
st = _engine.stream_handle()
def ev():
    e = C.c_void_p(); _lib.call("eqxv_event_create", C.byref(e)); return e
for _ in range(2):
    plan.run_steps(st)
_lib.call("eqxv_stream_sync", st)
evs = []
for fn, kw in plan.steps:
    a, b = ev(), ev()
    _lib.call("eqxv_event_record", a, st); fn(stream=st, **kw); _lib.call("eqxv_event_record", b, st)
    evs.append((fn, kw, a, b))
_lib.call("eqxv_stream_sync", st)
rows, agg = [], {}
for fn, kw, a, b in evs:
    t = C.c_float(); _lib.call("eqxv_event_elapsed_ms", a, b, C.byref(t))
    desc = fn.__name__
    for k in ("x", "a", "qkv", "patches"):
        if k in kw and hasattr(kw[k], "shape"):
            desc += f" in{tuple(kw[k].shape)}"; break
    if "out" in kw and hasattr(kw["out"], "shape"): desc += f" out{tuple(kw['out'].shape)}"
    for k in ("kh", "k", "stride", "dil"):
        if k in kw: desc += f" {k}={kw[k]}"
    rows.append((t.value, desc)); agg[fn.__name__] = agg.get(fn.__name__, [0, 0.0]); agg[fn.__name__][0] += 1; agg[fn.__name__][1] += t.value
tot = sum(r[0] for r in rows)
print(f"{name} batch {batch} @{hw}: {len(rows)} launches, sum {tot:.3f} ms -> {batch / tot * 1e3:.0f} img/s (eager, per-launch events)")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]): print(f"  {k:28s} x{n:4d} {t:8.3f} ms {100 * t / tot:5.1f}%")
for t, d in sorted(rows, key=lambda r: -r[0])[:topn]: print(f"    {t:7.3f} ms  {d}")
