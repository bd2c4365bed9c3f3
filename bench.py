#!/usr/bin/env python
"""bench.py — images/s of the B200 forward pass (BASELINE.json metric), one JSON line on rank 0.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...   # CPU arm: the oracle port of the reference
  torchrun --nproc-per-node N bench.py --gpus N ...         # one rank per GPU, weak scaling

Workload at N=1 (BASELINE.json configs[1]): ResNet-50 inference, bf16, batch 256 per GPU, synthetic
3x224x224 images, seeded synthetic checkpoint loaded through load_torch_weights. The other four BASELINE
configs are measured in the same run and reported under "secondary": ViT-B/16 (64 images per GPU = 512 over 8,
configs[2]; with N > 1 its logits all-gather is inside the timed step), EfficientNet-B4 (B=128, configs[3]),
DeepLabV3-ResNet50 (4x3x512x512, configs[4]) and the README's single-image AlexNet forward (configs[0]).
A "step" is one forward pass over one batch: one CUDA-graph replay (56 kernel launches for R50).

  value   : inputs resident in HBM, CUDA events on the launching stream, max over ranks
  e2e     : through the public call `eb.filter_jit(eb.vmap(net, axis_name="batch"))(images, key=keys)` with a PINNED HOST
            batch every step and a device-to-host copy of every model output every step, timed with CUDA events on
            torch's stream (which the public call orders its results on). Headline: uint8 HWC pixels in
            (`transforms.images_u8`: ToTensor + Normalize of the reference's fixture fused on the device);
            `e2e_f32` is the same with the reference's fp32 NCHW batch (4x the bytes over PCIe).
  roofline: tensor-bound models: algorithmic FLOPs (8.178 GFLOP/img for R50) over the WHOLE graph-replay
            step (the tcgen05 implicit-GEMM family is the dominant kernel and is charged with the rest of the step too),
            against the measured sustained cuBLAS bf16 peak; `by_kernel_ms` splits the step by C entry (shares from
            per-launch event pairs of an eager replay, scaled to the step: eager launches lose the PDL overlap, so their
            SUM exceeds the step and only the shares are used); HBM-bound models
            (EfficientNet-B4): whole-step algorithmic bytes against the measured copy bandwidth;
            `traffic` = ncu DRAM bytes of the step (profiles/traffic.json, regenerated per round)
  cpu_baseline: the CPU oracle (torch fp32 restatement of the reference, kind "port") on a bounded sample, timed on
            rank 0 AFTER the process group is gone (no rank spins on a barrier beside it)
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


def config_of(name: str, world: int) -> dict:
    """the `config` object: identical in the B200 arm and in the reference arm (same workload, same batch)"""
    hw = MODELS[name]["hw"]
    b = PER_GPU_BATCH[name]
    return {
        "workload": f"{name} inference 3x{hw}x{hw}, batch {b} per GPU ({MODELS[name]['config']})",
        "global_batch": b * world,
        "parallelism": f"dp{world} (batch sharded, weights replicated)",
        "l2": "every step's input + activations exceed the 126 MB L2 several times over: no explicit flush between "
              "timed iterations (AlexNet at batch 1: 122 MB of filters + L2-resident activations, stated in its line)",
        "checkpoint": "seeded synthetic state_dict via load_torch_weights",
    }


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
        # median over the samples taken under load (idle samples sit at the max clock as well on this part)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa_node(local_rank: int):
    """pin this rank's threads to the host NUMA node its GPU hangs off, BEFORE any pinned buffer is allocated (first
    touch places the pages): eight ranks pulling pinned pages through one remote node halved the per-GPU H2D rate in
    round 1. Best effort; returns a short description for the JSON line."""
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(local_rank), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = "0000:" + bus.split(":", 1)[1]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        if node < 0:
            return "numa: single node"
        cpus = []
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus += list(range(int(lo), int(hi or lo) + 1))
        allowed = sorted(set(cpus) & os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
            return f"numa: rank bound to node {node} ({len(allowed)} cpus)"
        return f"numa: node {node} has no allowed cpus"
    except Exception as e:  # noqa: BLE001
        return f"numa: not bound ({type(e).__name__})"


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


def _leaves(out):
    if out is None:
        return []
    if isinstance(out, (list, tuple)):
        return [t for o in out for t in _leaves(o)]
    return [out]


def measure(name, model, batch, steps, warmup, dist, rank, world, gather=False):
    """returns dict(ms, e2e_ms, e2e_f32_ms, launches, by_kernel_ms, ...)"""
    import torch

    import eqxvision_b200 as eb
    from eqxvision_b200 import _engine, _lib, ops, parallel
    from eqxvision_b200 import transforms as tr

    hw = MODELS[name]["hw"]
    in_shape = (3, hw, hw)
    keys = eb.random.split(eb.random.PRNGKey(0), batch)
    g = torch.Generator().manual_seed(100 + rank)
    pixels = torch.randint(0, 256, (batch, hw, hw, 3), generator=g, dtype=torch.uint8)
    host_u8 = tr.images_u8(pixels).pin_memory()                       # what a data loader hands over
    host_f32 = host_u8.reference_pipeline().pin_memory()              # the reference's fp32 NCHW input
    plan = _engine.get_plan(model, "__call__", batch, in_shape, (), {"key": keys})
    st = plan.stream
    plan.x_in.copy_(host_f32)
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
            torch.cuda.synchronize()

    # (1) device-resident inputs: graph replays on the plan's stream
    for _ in range(warmup):
        plan.launch(st)
    barrier()
    e0, e1 = ev(), ev()
    _lib.call("eqxv_event_record", e0, st)
    for _ in range(steps):
        plan.launch(st)
    _lib.call("eqxv_event_record", e1, st)
    _lib.call("eqxv_event_sync", e1)
    barrier()
    t = C.c_float()
    _lib.call("eqxv_event_elapsed_ms", e0, e1, C.byref(t))
    ms = t.value / steps

    # (2) end to end through the public API: pinned host batch in, every output copied back to pinned host memory
    fwd = eb.filter_jit(eb.vmap(model, axis_name="batch"))
    gatherer = [None]

    def e2e(host_batch):
        probe = _leaves(fwd(host_batch, key=keys))
        if gather and dist is not None and gatherer[0] is None:
            gatherer[0] = parallel.LogitsAllGather(probe[0].shape[0], probe[0].shape[1], probe[0].dtype)
        mirrors = None

        def one_step():
            nonlocal mirrors
            outs = _leaves(fwd(host_batch, key=keys))
            if gatherer[0] is not None:
                outs = [gatherer[0](outs[0])]                         # [world * B, classes] on every rank
            if mirrors is None:
                mirrors = [torch.empty(tuple(o.shape), dtype=o.dtype).pin_memory() for o in outs]
            for o, h in zip(outs, mirrors):
                h.copy_(o, non_blocking=True)

        for _ in range(max(warmup, _engine.PIPELINE_LANES + 1)):
            one_step()
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(steps):
            one_step()
        t1.record()
        t1.synchronize()
        barrier()
        return t0.elapsed_time(t1) / steps, sum(h.numel() * h.element_size() for h in mirrors)

    e2e_ms, d2h = e2e(host_u8)
    e2e_f32_ms, _ = e2e(host_f32)
    if gatherer[0] is not None:
        gatherer[0].close()

    # (3) share of each C entry in the step: per-launch event pairs of an eager replay (no PDL overlap, so the SUM
    # is larger than the graph step; only the shares are used, scaled to the measured step)
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
    eager, igemm_eager, igemm_n, by_kernel = 0.0, 0.0, 0, {}
    for fn, a, b in evs:
        _lib.call("eqxv_event_elapsed_ms", a, b, C.byref(t))
        by_kernel[fn.__name__] = by_kernel.get(fn.__name__, 0.0) + t.value
        eager += t.value
        if fn in igemm_fns:
            igemm_eager += t.value
            igemm_n += 1
    scale = ms / eager
    res = {"ms": ms, "e2e_ms": e2e_ms, "e2e_f32_ms": e2e_f32_ms, "launches": plan.num_launches,
           "igemm_ms": igemm_eager * scale, "igemm_share": igemm_eager / eager, "igemm_launches": igemm_n,
           "by_kernel_ms": {k: round(v * scale, 4) for k, v in by_kernel.items()}, "eager_sum_ms": eager,
           "h2d_u8": pixels.numel(), "h2d_f32": host_f32.numel() * 4, "d2h": d2h, "act_bytes": plan.act_bytes,
           "gathered": gatherer[0] is not None}
    _engine.clear_plans(model)
    return res


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

    try:
        os.sched_setaffinity(0, range(os.cpu_count() or 1))   # undo the NUMA binding of the GPU arm: all host cores
    except OSError:
        pass
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
        "config": config_of(name, args.gpus),
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
    ap.add_argument("--all-configs", action="store_true", help="kept for compatibility: all five BASELINE configs "
                                                               "are measured by default")
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
    numa = bind_to_gpu_numa_node(local_rank)
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
    m = measure(name, model, batch, args.steps, args.warmup, dist, rank, world)
    clocks = sampler.stop() if rank == 0 else None
    del model

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms = max_over_ranks(m["ms"])
    e2e_ms = max_over_ranks(m["e2e_ms"])
    e2e_f32_ms = max_over_ranks(m["e2e_f32_ms"])
    total_imgs = batch * world

    def roofline_of(nm, mm, step_ms):
        """the model's binding roofline for the dominant kernel family"""
        nb = PER_GPU_BATCH[nm]
        if MODELS[nm]["bound"] == "tensor":
            # the WHOLE graph-replay step is the denominator (the conv/linear family is the dominant kernel and the rest
            # of the step is charged to it): a per-kernel time cannot exceed the step it is part of
            tf = FLOP_PER_IMG[nm] * nb / step_ms / 1e9
            r = {"bound": "tensor", "kernel": "tcgen05 implicit-GEMM family (igemm_kernel, pair_kernel, halo_kernel, "
                                              "stem_kernel) charged with the whole step",
                 "achieved": round(tf, 1), "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
                 "frac": round(tf / peaks["tflops_sustained"], 4), "traffic": None,
                 "peak_source": peaks["source"] + " (sustained cuBLAS bf16)",
                 "launches": mm["igemm_launches"], "kernel_ms_per_step": round(step_ms, 4),
                 "family_share_of_eager_launch_time": round(mm["igemm_share"], 4),
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

    def e2e_of(mm, nb, a_ms, f_ms):
        return {"value": round(nb * world / a_ms * 1e3, 1), "unit": "img/s", "ms_per_step": round(a_ms, 4),
                "h2d_bytes_per_step": mm["h2d_u8"], "d2h_bytes_per_step": mm["d2h"],
                "api": "eb.filter_jit(eb.vmap(net, axis_name='batch'))(transforms.images_u8(pinned uint8 NHWC), "
                       "key=keys); every output .copy_() to pinned host memory",
                "input": "uint8 HWC pixels; ToTensor + Normalize of the reference fixture fused on the device"}, \
               {"value": round(nb * world / f_ms * 1e3, 1), "unit": "img/s", "ms_per_step": round(f_ms, 4),
                "h2d_bytes_per_step": mm["h2d_f32"], "d2h_bytes_per_step": mm["d2h"],
                "input": "the reference's fp32 NCHW batch (pinned host memory), same public call"}

    secondary = {}
    others = [] if args.no_secondary else [k for k in ("vit_base", "resnet50", "efficientnet_b4",
                                                       "deeplabv3_resnet50", "alexnet") if k != name]
    for other in others:
        om_model, _ = build_model(other)
        ob = PER_GPU_BATCH[other]
        gathered = other == "vit_base" and world > 1      # configs[2] asks for the gathered logits
        sm = measure(other, om_model, ob, max(10, args.steps // 2), args.warmup, dist, rank, world, gather=gathered)
        s_ms = max_over_ranks(sm["ms"])
        a, f = e2e_of(sm, ob, max_over_ranks(sm["e2e_ms"]), max_over_ranks(sm["e2e_f32_ms"]))
        if gathered:
            a["gathered"] = f["gathered"] = True
            a["collective"] = "one all-gather of the fp32 logits per step inside the timed region (parallel.LogitsAllGather)"
        secondary[other] = {
            "value": round(ob * world / s_ms * 1e3, 1), "unit": "img/s", "ms_per_step": round(s_ms, 4),
            "per_gpu_batch": ob, "e2e": a, "e2e_f32": f, "config": MODELS[other]["config"],
            "step_tflops_per_gpu": round(FLOP_PER_IMG[other] * ob / s_ms / 1e9, 1),
            "step_hbm_gbs_per_gpu": round(BYTES_PER_IMG[other] * ob / s_ms / 1e6, 1),
            "roofline": roofline_of(other, sm, s_ms),
            "by_kernel_ms": sm["by_kernel_ms"], "launches": sm["launches"],
            "act_bytes": sm["act_bytes"],
        }
        del om_model
        torch.cuda.empty_cache()

    # the process group goes away BEFORE the CPU sample: no rank spins on a barrier next to it
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return

    a, f = e2e_of(m, batch, e2e_ms, e2e_f32_ms)
    line = {
        "metric": "images/sec", "value": round(total_imgs / ms * 1e3, 1), "unit": "img/s", "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": config_of(name, world),
        "e2e": a, "e2e_f32": f,
        "gpu_launches": m["launches"] * args.steps,
        "launches_per_step": m["launches"],
        "roofline": roofline_of(name, m, ms),
        "step_tflops": round(FLOP_PER_IMG[name] * batch / ms / 1e9, 1),
        "step_hbm": {"achieved": round(BYTES_PER_IMG[name] * batch / ms / 1e6, 1), "peak": peaks["hbm_gbs"],
                     "unit": "GB/s", "frac": round(BYTES_PER_IMG[name] * batch / ms / 1e6 / peaks["hbm_gbs"], 4),
                     "note": "whole step against the layer-wise algorithmic bytes (SURVEY.md §8(d)): the practical "
                             "ceiling of a layer-by-layer bf16 forward"},
        "by_kernel_ms": m["by_kernel_ms"],
        "eager_sum_ms": round(m["eager_sum_ms"], 4),
        "activation_bytes": m["act_bytes"],
        "clocks": clocks,
        "host": numa,
        "secondary": secondary,
    }
    if not args.no_cpu_baseline:
        v, cores, sample = cpu_port(name, sd, MODELS[name]["cpu_batch"], 3)
        line["cpu_baseline"] = {"value": round(v, 2), "unit": "img/s", "cores": cores, "kind": "port",
                                "sample": sample}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
