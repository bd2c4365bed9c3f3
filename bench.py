#!/usr/bin/env python
"""bench.py — images/s of the B200 forward pass (BASELINE.json metric), one JSON line on rank 0.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...   # CPU arm: the oracle port of the reference
  torchrun --nproc-per-node N bench.py --gpus N ...         # one rank per GPU, weak scaling, no collective

Workload at N=1 (BASELINE.json configs[1]): ResNet-50 inference, bf16, batch 256 per GPU, synthetic
3x224x224 images, seeded synthetic checkpoint loaded through load_torch_weights. ViT-B/16 (64 images
per GPU = 512 over 8, configs[2]) is measured in the same run and reported under "secondary";
`--all-configs` adds EfficientNet-B4 (B=128, configs[3]), DeepLabV3-ResNet50 (4x3x512x512, configs[4]) and the
README's single-image AlexNet forward (configs[0]).
A "step" is one forward pass over one batch: one CUDA-graph replay (57 kernel launches for R50).

  value   : inputs resident in HBM, CUDA events on the launching stream, max over ranks
  e2e     : pinned host fp32 NCHW batch -> H2D -> graph -> D2H of every model output, every step, same
            events; EQXV_E2E_LANES (default 3) plans are used round-robin so that the copy of step i+1
            overlaps the kernels of step i, graph launches are FIFO-chained across lanes
  roofline: tensor-bound models: the tcgen05 implicit-GEMM family (all conv/linear launches of a step),
            algorithmic FLOPs (8.178 GFLOP/img for R50) over the summed per-launch device
            time of those launches, against the measured sustained cuBLAS bf16 peak; HBM-bound models
            (EfficientNet-B4): whole-step algorithmic bytes against the measured copy bandwidth;
            `traffic` = ncu DRAM bytes of those launches (profiles/traffic.json)
  cpu_baseline: the CPU oracle (torch fp32 restatement of the reference, kind "port") on a bounded sample
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# algorithmic FLOPs / bytes per image (bf16 activations, every conv/linear reads its
# input once and writes its output once, norm + activation + residual fused), the roofline that binds,
# per-GPU batch (BASELINE.json configs) and the bounded CPU sample.
MODELS = {
    "resnet50": dict(batch=256, hw=224, flop=8.178e9, bytes=56.8e6, bound="tensor", cpu_batch=16,
                     config="BASELINE.json configs[1]"),
    "vit_base": dict(batch=64, hw=224, flop=35.13e9, bytes=94.4e6, bound="tensor", cpu_batch=8,
                     config="BASELINE.json configs[2]: 512 images over 8 GPUs = 64 per GPU"),
    "efficientnet_b4": dict(batch=128, hw=224, flop=3.007e9, bytes=70.54e6, bound="hbm", cpu_batch=8,
                            config="BASELINE.json configs[3]"),
    "deeplabv3_resnet50": dict(batch=4, hw=512, flop=346.5e9, bytes=0.69e9, bound="tensor", cpu_batch=1,
                               config="BASELINE.json configs[4]"),
    # configs[0], the README example: ONE image. At batch 1 the step is bound by reading the 61.1 M parameters
    # once (122.2 MB of bf16 filters, 117 MB of them in the three classifier GEMMs) + 1.4 MB of activations
    "alexnet": dict(batch=1, hw=224, flop=1.428e9, bytes=123.6e6, bound="hbm", cpu_batch=1,
                    config="BASELINE.json configs[0] (README example, 1x3x224x224)"),
}
FLOP_PER_IMG = {k: v["flop"] for k, v in MODELS.items()}
BYTES_PER_IMG = {k: v["bytes"] for k, v in MODELS.items()}
PER_GPU_BATCH = {k: v["batch"] for k, v in MODELS.items()}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tflops_burst": d["bf16_tflops"],
                "tflops_sustained": d["bf16_tflops_sustained"], "source": "measured"}
    return {"hbm_gbs": 6650.0, "tflops_burst": 1590.0, "tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [v.strip() for v in ln.split(",")]
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
                for nme, v in zip(names, f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:  # noqa: BLE001
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_state_dict(name: str):
    from tools import synthetic as syn   # workload generator (not the oracle)

    if name == "vit_base":
        return syn.vit_state_dict(embed_dim=768, depth=12, heads=12, num_classes=1000, seed=3)
    if name == "deeplabv3_resnet50":
        return syn.torchvision_model(name, seed=1, calib_hw=64, aux_loss=True).state_dict()
    return syn.torchvision_state_dict(name, seed=1)


def build_model(name: str):
    """seeded synthetic checkpoint -> .pth -> constructor(torch_weights=...) -> inference mode"""
    import torch

    import eqxvision_b200 as eb

    sd = synthetic_state_dict(name)
    f = tempfile.NamedTemporaryFile(suffix=".pth", delete=False)
    torch.save(sd, f.name)
    if name == "vit_base":
        model = eb.models.vit_base(num_classes=1000, torch_weights=f.name)
    elif name == "deeplabv3_resnet50":
        model = eb.models.deeplabv3(intermediate_layers=lambda m: [m.layer3, m.layer4], aux_in_channels=1024,
                                    torch_weights=f.name)
    else:
        model = getattr(eb.models, name)(torch_weights=f.name)
    os.unlink(f.name)
    return eb.tree_inference(model, True), sd


def measure(name, model, batch, steps, warmup, dist, rank):
    """returns dict(ms, e2e_ms, launches, igemm_ms, plan)"""
    import torch

    import eqxvision_b200 as eb
    from eqxvision_b200 import _engine, _lib, ops

    hw = MODELS[name]["hw"]
    in_shape = (3, hw, hw)
    plan = _engine.get_plan(model, "__call__", batch, in_shape, (), {"key": eb.random.PRNGKey(0)})
    st = _engine.stream_handle()
    g = torch.Generator().manual_seed(100 + rank)
    host_in = torch.rand((batch,) + in_shape, generator=g).pin_memory()

    def host_mirrors(pl):   # one pinned host buffer per model output (DeepLabV3 returns (aux, out))
        pairs = []
        for o, _ in pl.outputs:
            assert o.is_contiguous()
            pairs.append((o, torch.empty(tuple(o.shape), dtype=o.dtype).pin_memory()))
        return pairs

    outs1 = host_mirrors(plan)
    plan.x_in.copy_(host_in)
    torch.cuda.synchronize()

    def ev():
        e = C.c_void_p()
        _lib.call("eqxv_event_create", C.byref(e))
        return e

    def barrier():
        _lib.call("eqxv_stream_sync", st)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def timed(step_fn):
        for _ in range(warmup):
            step_fn()
        barrier()
        e0, e1 = ev(), ev()
        _lib.call("eqxv_event_record", e0, st)
        for _ in range(steps):
            step_fn()
        _lib.call("eqxv_event_record", e1, st)
        _lib.call("eqxv_event_sync", e1)
        barrier()
        ms = C.c_float()
        _lib.call("eqxv_event_elapsed_ms", e0, e1, C.byref(ms))
        return ms.value / steps

    # (1) device-resident inputs
    ms = timed(lambda: plan.launch(st))

    # (2) end to end through host buffers: every step copies its own pinned fp32 NCHW batch to the
    # device, replays the graph and copies the logits back. LANES plans (own buffers, own stream) are used
    # round-robin so that the H2D copy of step i+1 overlaps the kernels of step i (copy engine vs SMs);
    # graph launches are chained with events across lanes (compute runs FIFO, never two graphs interleaved:
    # interleaving delays the completion of BOTH graphs and with it the next copies).
    nbytes_in = host_in.numel() * 4
    nbytes_out = sum(o.numel() * o.element_size() for o, _ in outs1)
    n_lanes = int(os.environ.get("EQXV_E2E_LANES", "3"))
    lanes = [(plan, st, outs1)]
    for _ in range(n_lanes - 1):
        pl = _engine.build_plan(model, "__call__", batch, in_shape, (), {"key": eb.random.PRNGKey(0)})
        sx = C.c_void_p()
        _lib.call("eqxv_stream_create", C.byref(sx))
        lanes.append((pl, sx.value, host_mirrors(pl)))
    counter = [0]
    graph_done = [None]
    ev_ring = [ev() for _ in range(8)]

    def e2e_step():
        pl, s, outs = lanes[counter[0] % n_lanes]
        counter[0] += 1
        _lib.call("eqxv_memcpy_h2d_async", pl.x_in.data_ptr(), host_in.data_ptr(), nbytes_in, s)
        if graph_done[0] is not None:
            _lib.call("eqxv_stream_wait_event", s, graph_done[0])   # FIFO on the SMs
        pl.launch(s)
        e = ev_ring[counter[0] % len(ev_ring)]
        _lib.call("eqxv_event_record", e, s)
        graph_done[0] = e
        for o, ho in outs:
            _lib.call("eqxv_memcpy_d2h_async", ho.data_ptr(), o.data_ptr(), o.numel() * o.element_size(), s)

    def sync_lanes():
        for _, s, _ in lanes:
            _lib.call("eqxv_stream_sync", s)

    def timed_lanes():
        for _ in range(max(warmup, n_lanes)):
            e2e_step()
        sync_lanes()
        barrier()
        e0, e1 = ev(), ev()
        _lib.call("eqxv_event_record", e0, st)
        for _, s, _ in lanes[1:]:
            _lib.call("eqxv_stream_wait_event", s, e0)   # every lane starts after e0
        for _ in range(steps):
            e2e_step()
        for _, s, _ in lanes[1:]:                        # e1 is after the last step of EVERY lane
            j = ev()
            _lib.call("eqxv_event_record", j, s)
            _lib.call("eqxv_stream_wait_event", st, j)
        _lib.call("eqxv_event_record", e1, st)
        _lib.call("eqxv_event_sync", e1)
        sync_lanes()
        barrier()
        ms_ = C.c_float()
        _lib.call("eqxv_event_elapsed_ms", e0, e1, C.byref(ms_))
        return ms_.value / steps

    e2e_ms = timed_lanes()
    del lanes[1:]

    # (3) per-launch device time of the igemm (conv / linear) launches, eager replay with events
    igemm_fns = (ops.conv2d, ops.gemm, ops.conv_stem, ops.conv_stem7x7)
    plan.run_steps(st)
    _lib.call("eqxv_stream_sync", st)
    evs = []
    for fn, kw in plan.steps:
        a, b = ev(), ev()
        _lib.call("eqxv_event_record", a, st)
        fn(stream=st, **kw)
        _lib.call("eqxv_event_record", b, st)
        evs.append((fn, a, b))
    _lib.call("eqxv_stream_sync", st)
    igemm_ms, igemm_n, by_kernel = 0.0, 0, {}
    for fn, a, b in evs:
        t = C.c_float()
        _lib.call("eqxv_event_elapsed_ms", a, b, C.byref(t))
        by_kernel[fn.__name__] = by_kernel.get(fn.__name__, 0.0) + t.value
        if fn in igemm_fns:
            igemm_ms += t.value
            igemm_n += 1
    return {"ms": ms, "e2e_ms": e2e_ms, "launches": plan.num_launches, "igemm_ms": igemm_ms,
            "igemm_launches": igemm_n, "by_kernel_ms": {k: round(v, 4) for k, v in by_kernel.items()},
            "h2d": nbytes_in, "d2h": nbytes_out, "act_bytes": plan.act_bytes}


def oracle_forward(name, sd, batch):
    """the CPU oracle (torch fp32 restatement of the reference) as a zero-argument callable on a seeded
    `batch`-image sample; the ONLY place bench.py touches oracle/ (cpu_baseline leg + --impl reference)"""
    from oracle import models as om
    from tools import synthetic as syn

    hw = MODELS[name]["hw"]
    x = syn.synthetic_images(batch, h=hw, w=hw, seed=7)
    if name == "vit_base":
        return lambda: om.vit(sd, x, heads=12)
    if name == "deeplabv3_resnet50":
        return lambda: om.deeplabv3_resnet50(sd, x)
    if name.startswith("efficientnet"):
        return lambda: om.efficientnet(sd, x, name)
    if name == "alexnet":
        return lambda: om.alexnet(sd, x)
    return lambda: om.resnet(sd, x, name)


def cpu_port(name, sd, batch, iters):
    """bounded CPU sample; returns (img/s, cores, sample)"""
    import torch

    torch.set_num_threads(max(1, os.cpu_count() or 1))
    fn = oracle_forward(name, sd, batch)
    with torch.no_grad():
        fn()  # warm-up
        t0 = time.perf_counter()
        for _ in range(iters):
            fn()
        dt = (time.perf_counter() - t0) / iters
    return batch / dt, torch.get_num_threads(), f"{iters} x batch {batch} fp32 forward of {name} (torch CPU oracle)"


def run_reference_arm(args, rank):
    """--impl reference: the reference's CPU path. jax/equinox cannot be installed in this image, so
    the arm times the oracle port (kind "port") with all host threads, rank 0 only."""
    if rank != 0:
        return
    import torch

    # torchrun exports OMP_NUM_THREADS=1 to every rank: the CPU arm uses all host cores explicitly
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    name = args.model
    hw = MODELS[name]["hw"]
    sd = synthetic_state_dict(name)
    sample_batch = MODELS[name]["cpu_batch"]
    fn = oracle_forward(name, sd, sample_batch)
    with torch.no_grad():
        for _ in range(max(1, min(args.warmup, 2))):
            fn()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            fn()
        dt = (time.perf_counter() - t0) / args.steps
    v = sample_batch / dt
    cores = torch.get_num_threads()
    sample = (f"each step = one fp32 forward of {sample_batch} images (bounded sample of the "
              f"{PER_GPU_BATCH[name]}-image batch) on {cores} host threads")
    line = {
        "impl": "reference", "metric": "images/sec", "value": round(v, 2), "unit": "img/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{name} inference 3x{hw}x{hw}, batch {PER_GPU_BATCH[name]} per GPU",
                   "global_batch": PER_GPU_BATCH[name] * args.gpus},
        "cpu_baseline": {"value": round(v, 2), "unit": "img/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": round(v, 2), "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "jax/equinox are not installable offline; this arm is the CPU oracle port of the reference",
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="resnet50", choices=sorted(MODELS))
    ap.add_argument("--no-secondary", action="store_true")
    ap.add_argument("--all-configs", action="store_true",
                    help="also measure EfficientNet-B4 (B=128), DeepLabV3-R50 (4x3x512x512) and AlexNet (1 image) "
                         "as secondary lines")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference_arm(args, rank)
        return

    import torch

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = dist_mod

    peaks = load_peaks()
    name = args.model
    batch = PER_GPU_BATCH[name]
    model, sd = build_model(name)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    m = measure(name, model, batch, args.steps, args.warmup, dist, rank)
    clocks = sampler.stop() if rank == 0 else None

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms = max_over_ranks(m["ms"])
    e2e_ms = max_over_ranks(m["e2e_ms"])
    total_imgs = batch * world
    value = total_imgs / ms * 1e3
    e2e_value = total_imgs / e2e_ms * 1e3

    def roofline_of(nm, mm, step_ms):
        """the model's binding roofline for the dominant kernel family"""
        nb = PER_GPU_BATCH[nm]
        if MODELS[nm]["bound"] == "tensor":
            tf = FLOP_PER_IMG[nm] * nb / mm["igemm_ms"] / 1e9
            r = {"bound": "tensor", "kernel": "tcgen05 implicit-GEMM family (all conv/linear launches of one step: "
                                              "igemm_kernel, pair_kernel, halo_kernel, stem_kernel)",
                 "achieved": round(tf, 1), "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                 "frac": round(tf / peaks["tflops_sustained"], 4), "traffic": None,
                 "peak_source": peaks["source"] + " (sustained cuBLAS bf16)",
                 "launches": mm["igemm_launches"], "kernel_ms_per_step": round(mm["igemm_ms"], 4),
                 "flop_per_image": FLOP_PER_IMG[nm]}
        else:
            gbs = BYTES_PER_IMG[nm] * nb / step_ms / 1e6
            what = ("whole step (batch 1: every filter is read once, the classifier GEMMs are weight-bandwidth bound)"
                    if nm == "alexnet" else
                    "whole step (pointwise 1x1 GEMMs + depthwise stencils + SE; HBM-bound model)")
            r = {"bound": "hbm", "kernel": what,
                 "achieved": round(gbs, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s",
                 "frac": round(gbs / peaks["hbm_gbs"], 4), "traffic": None,
                 "peak_source": peaks["source"] + " (copy bandwidth)", "launches": mm["launches"],
                 "kernel_ms_per_step": round(step_ms, 4), "bytes_per_image": BYTES_PER_IMG[nm]}
        tp = os.path.join(ROOT, "profiles", "traffic.json")   # dram bytes per step from the committed ncu pass
        if os.path.exists(tp):
            t = json.load(open(tp)).get(nm)
            if t:
                r["traffic"] = t["dram_bytes_per_step"]
                r["traffic_source"] = t["source"]
        return r

    secondary = {}
    others = []
    if not args.no_secondary:
        others.append("vit_base" if name == "resnet50" else "resnet50")
    if args.all_configs:
        others += [k for k in ("efficientnet_b4", "deeplabv3_resnet50", "alexnet") if k != name and k not in others]
    for other in others:
        om_model, _ = build_model(other)
        ob = PER_GPU_BATCH[other]
        sm = measure(other, om_model, ob, max(10, args.steps // 2), args.warmup, dist, rank)
        s_ms = max_over_ranks(sm["ms"])
        s_e2e = max_over_ranks(sm["e2e_ms"])
        secondary[other] = {
            "value": round(ob * world / s_ms * 1e3, 1), "unit": "img/s", "ms_per_step": round(s_ms, 4),
            "per_gpu_batch": ob, "e2e": round(ob * world / s_e2e * 1e3, 1), "config": MODELS[other]["config"],
            "step_tflops_per_gpu": round(FLOP_PER_IMG[other] * ob / s_ms / 1e9, 1),
            "step_hbm_gbs_per_gpu": round(BYTES_PER_IMG[other] * ob / s_ms / 1e6, 1),
            "roofline": roofline_of(other, sm, s_ms),
            "by_kernel_ms": sm["by_kernel_ms"], "launches": sm["launches"],
        }
        del om_model
        torch.cuda.empty_cache()

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    hw = MODELS[name]["hw"]
    line = {
        "metric": "images/sec", "value": round(value, 1), "unit": "img/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {
            "workload": f"{name} inference 3x{hw}x{hw}, batch {batch} per GPU ({MODELS[name]['config']})",
            "global_batch": total_imgs, "parallelism": f"dp{world} (batch sharded, weights replicated, no collective)",
            "l2": f"per-step working set {m['act_bytes'] / 2**30:.1f} GiB of activations + "
                  f"{m['h2d'] / 2**20:.0f} MiB input >> 126 MB L2 (no explicit flush needed)",
            "checkpoint": "seeded synthetic state_dict via load_torch_weights",
        },
        "e2e": {"value": round(e2e_value, 1), "unit": "img/s", "ms_per_step": round(e2e_ms, 4),
                "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"]},
        "gpu_launches": m["launches"] * args.steps,
        "launches_per_step": m["launches"],
        "roofline": roofline_of(name, m, ms),
        "step_tflops": round(FLOP_PER_IMG[name] * batch / ms / 1e9, 1),
        "step_hbm": {"achieved": round(BYTES_PER_IMG[name] * batch / ms / 1e6, 1), "peak": peaks["hbm_gbs"],
                     "unit": "GB/s", "frac": round(BYTES_PER_IMG[name] * batch / ms / 1e6 / peaks["hbm_gbs"], 4),
                     "note": "whole step against the layer-wise algorithmic bytes (SURVEY.md §8(d)): the practical "
                             "ceiling of a layer-by-layer bf16 forward"},
        "by_kernel_ms": m["by_kernel_ms"],
        "clocks": clocks,
        "secondary": secondary,
    }
    if not args.no_cpu_baseline:
        v, cores, sample = cpu_port(name, sd, MODELS[name]["cpu_batch"], 3)
        line["cpu_baseline"] = {"value": round(v, 2), "unit": "img/s", "cores": cores, "kind": "port",
                                "sample": sample}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
