#!/usr/bin/env python
"""bench.py — images/s of the B200 forward pass (BASELINE.json metric), one JSON line on rank 0.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...   # CPU arm: the oracle port of the reference
  torchrun --nproc-per-node N bench.py --gpus N ...         # one rank per GPU, weak scaling

Workload at N=1 (BASELINE.json configs[1]): ResNet-50 inference, bf16, batch 256 per GPU, synthetic
3x224x224 images, seeded synthetic checkpoint loaded through load_torch_weights. ViT-B/16 (64 images
per GPU = 512 over 8, configs[2]) is measured in the same run and reported under "secondary".
A "step" is one forward pass over one batch: one CUDA-graph replay (57 kernel launches for R50).

  value  : inputs resident in HBM, CUDA events on the launching stream, max over ranks
  e2e    : pinned host fp32 NCHW batch -> H2D -> graph -> D2H logits, every step, same events
  roofline: the tcgen05 implicit-GEMM kernel (all conv/linear launches of a step), algorithmic FLOPs
           (8.178 GFLOP/img) over the summed per-launch device time of those launches
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

FLOP_PER_IMG = {"resnet50": 8.178e9, "vit_base": 35.13e9}
BYTES_PER_IMG = {"resnet50": 56.8e6, "vit_base": 94.4e6}           # layer-wise bf16 traffic
PER_GPU_BATCH = {"resnet50": 256, "vit_base": 64}


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


def build_model(name: str):
    """seeded synthetic checkpoint -> .pth -> constructor(torch_weights=...) -> inference mode"""
    import torch

    import eqxvision_b200 as eb
    from oracle import checkpoints as ck

    if name == "resnet50":
        sd = ck.torchvision_state_dict("resnet50", seed=1)
    else:
        sd = ck.vit_state_dict(embed_dim=768, depth=12, heads=12, num_classes=1000, seed=3)
    f = tempfile.NamedTemporaryFile(suffix=".pth", delete=False)
    torch.save(sd, f.name)
    model = getattr(eb.models, name)(torch_weights=f.name) if name == "resnet50" else \
        eb.models.vit_base(num_classes=1000, torch_weights=f.name)
    os.unlink(f.name)
    return eb.tree_inference(model, True), sd


def measure(name, model, batch, steps, warmup, dist, rank):
    """returns dict(ms, e2e_ms, launches, igemm_ms, plan)"""
    import torch

    import eqxvision_b200 as eb
    from eqxvision_b200 import _engine, _lib, ops

    plan = _engine.get_plan(model, "__call__", batch, (3, 224, 224), (), {"key": eb.random.PRNGKey(0)})
    st = _engine.stream_handle()
    g = torch.Generator().manual_seed(100 + rank)
    host_in = torch.rand((batch, 3, 224, 224), generator=g).pin_memory()
    out_t, _ = plan.outputs[0]
    host_out = torch.empty(tuple(out_t.shape), dtype=out_t.dtype).pin_memory()
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
    # device, replays the graph and copies the logits back. Two plans (own buffers, own stream) are
    # alternated so that the H2D copy of step i+1 overlaps the kernels of step i (copy engine vs SMs).
    nbytes_in = host_in.numel() * 4
    nbytes_out = out_t.numel() * out_t.element_size()
    assert out_t.is_contiguous()
    plan2 = _engine.build_plan(model, "__call__", batch, (3, 224, 224), (), {"key": eb.random.PRNGKey(0)})
    st2 = C.c_void_p()
    _lib.call("eqxv_stream_create", C.byref(st2))
    st2 = st2.value
    out2, _ = plan2.outputs[0]
    host_out2 = torch.empty(tuple(out2.shape), dtype=out2.dtype).pin_memory()
    lanes = [(plan, st, out_t, host_out), (plan2, st2, out2, host_out2)]
    counter = [0]

    def e2e_step():
        pl, s, o, ho = lanes[counter[0] & 1]
        counter[0] += 1
        _lib.call("eqxv_memcpy_h2d_async", pl.x_in.data_ptr(), host_in.data_ptr(), nbytes_in, s)
        pl.launch(s)
        _lib.call("eqxv_memcpy_d2h_async", ho.data_ptr(), o.data_ptr(), nbytes_out, s)

    def timed_two_streams():
        for _ in range(max(warmup, 2)):
            e2e_step()
        _lib.call("eqxv_stream_sync", st2)
        barrier()
        e0, e1, j = ev(), ev(), ev()
        _lib.call("eqxv_event_record", e0, st)
        _lib.call("eqxv_stream_wait_event", st2, e0)   # both lanes start after e0
        for _ in range(steps):
            e2e_step()
        _lib.call("eqxv_event_record", j, st2)
        _lib.call("eqxv_stream_wait_event", st, j)     # e1 is after the last step of BOTH lanes
        _lib.call("eqxv_event_record", e1, st)
        _lib.call("eqxv_event_sync", e1)
        _lib.call("eqxv_stream_sync", st2)
        barrier()
        ms_ = C.c_float()
        _lib.call("eqxv_event_elapsed_ms", e0, e1, C.byref(ms_))
        return ms_.value / steps

    e2e_ms = timed_two_streams()
    del plan2

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


def cpu_port(name, sd, batch, iters):
    """the CPU oracle (restatement of the reference) on a bounded sample; returns (img/s, cores, sample)"""
    import torch

    from oracle import checkpoints as ck
    from oracle import models as om

    x = ck.synthetic_images(batch, seed=7)
    fn = (lambda: om.resnet(sd, x, "resnet50")) if name == "resnet50" else (lambda: om.vit(sd, x, heads=12))
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

    from oracle import checkpoints as ck

    name = args.model
    sd = ck.torchvision_state_dict("resnet50", seed=1) if name == "resnet50" else \
        ck.vit_state_dict(num_classes=1000, seed=3)
    sample_batch = 16 if name == "resnet50" else 8
    from oracle import models as om

    x = ck.synthetic_images(sample_batch, seed=7)
    fn = (lambda: om.resnet(sd, x, "resnet50")) if name == "resnet50" else (lambda: om.vit(sd, x, heads=12))
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
        "config": {"workload": f"{name} inference 3x224x224, batch {PER_GPU_BATCH[name]} per GPU",
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
    ap.add_argument("--model", default="resnet50", choices=["resnet50", "vit_base"])
    ap.add_argument("--no-secondary", action="store_true")
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

    secondary = {}
    if not args.no_secondary:
        other = "vit_base" if name == "resnet50" else "resnet50"
        om_model, _ = build_model(other)
        ob = PER_GPU_BATCH[other]
        sm = measure(other, om_model, ob, max(10, args.steps // 2), args.warmup, dist, rank)
        s_ms = max_over_ranks(sm["ms"])
        s_e2e = max_over_ranks(sm["e2e_ms"])
        secondary[other] = {
            "value": round(ob * world / s_ms * 1e3, 1), "unit": "img/s", "ms_per_step": round(s_ms, 4),
            "per_gpu_batch": ob, "e2e": round(ob * world / s_e2e * 1e3, 1),
            "tflops_per_gpu": round(FLOP_PER_IMG[other] * ob / s_ms / 1e9, 1),
            "frac_of_sustained_peak": round(FLOP_PER_IMG[other] * ob / s_ms / 1e9 / peaks["tflops_sustained"], 4),
            "igemm_tflops": round(FLOP_PER_IMG[other] * ob / sm["igemm_ms"] / 1e9, 1),
            "by_kernel_ms": sm["by_kernel_ms"], "launches": sm["launches"],
        }

    if rank != 0:
        if dist is not None:
            dist.barrier()
            dist.destroy_process_group()
        return

    igemm_tflops = FLOP_PER_IMG[name] * batch / m["igemm_ms"] / 1e9
    line = {
        "metric": "images/sec", "value": round(value, 1), "unit": "img/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {
            "workload": f"{name} inference 3x224x224, batch {batch} per GPU (BASELINE.json configs[1])",
            "global_batch": total_imgs, "parallelism": f"dp{world} (batch sharded, weights replicated, no collective)",
            "l2": f"per-step working set {m['act_bytes'] / 2**30:.1f} GiB of activations + "
                  f"{m['h2d'] / 2**20:.0f} MiB input >> 126 MB L2 (no explicit flush needed)",
            "checkpoint": "seeded synthetic state_dict via load_torch_weights",
        },
        "e2e": {"value": round(e2e_value, 1), "unit": "img/s", "ms_per_step": round(e2e_ms, 4),
                "h2d_bytes_per_step": m["h2d"], "d2h_bytes_per_step": m["d2h"]},
        "gpu_launches": m["launches"] * args.steps,
        "launches_per_step": m["launches"],
        "roofline": {
            "bound": "tensor", "kernel": "eqxv::igemm_kernel (all conv/linear launches of one step)",
            "achieved": round(igemm_tflops, 1), "peak": peaks["tflops_sustained"], "unit": "TFLOP/s",
            "frac": round(igemm_tflops / peaks["tflops_sustained"], 4), "traffic": None,
            "peak_source": peaks["source"] + " (sustained cuBLAS bf16)",
            "launches": m["igemm_launches"], "kernel_ms_per_step": round(m["igemm_ms"], 4),
            "flop_per_image": FLOP_PER_IMG[name],
        },
        "roofline_hbm": {
            "bound": "hbm", "achieved": round(BYTES_PER_IMG[name] * batch / ms / 1e6, 1), "peak": peaks["hbm_gbs"],
            "unit": "GB/s", "frac": round(BYTES_PER_IMG[name] * batch / ms / 1e6 / peaks["hbm_gbs"], 4),
            "note": "whole step, layer-wise algorithmic bytes (SURVEY.md §8(d)): the practical ceiling of "
                    "unfused ResNet-50 in bf16",
        },
        "step_tflops": round(FLOP_PER_IMG[name] * batch / ms / 1e9, 1),
        "by_kernel_ms": m["by_kernel_ms"],
        "clocks": clocks,
        "secondary": secondary,
    }
    if not args.no_cpu_baseline:
        v, cores, sample = cpu_port(name, sd, 16 if name == "resnet50" else 8, 3)
        line["cpu_baseline"] = {"value": round(v, 2), "unit": "img/s", "cores": cores, "kind": "port",
                                "sample": sample}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
