#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (BASELINE.json configs[1]):

    rgbd_imagenet_adm_128_large_cfg, unconditional view, DDPM 1000 steps + classifier-free guidance, batch 16 per GPU,
    synthetic class labels + synthetic (seeded N(0,1/fan_in)) weights.

One bench "step" = ONE denoising step of the whole batch: batch-2N UNet forward (both CFG halves) + fused eps mix +
x_{t-1} update.  The 1000 steps of a sample are homogeneous (same kernels, same shapes; only the table row differs),
so   samples/s = n_gpus * batch / (1000 * step_seconds)   — `denoise_steps_per_sample` is stated in `config`.
`--full` times complete 1000-step samples instead (one bench step = one batch of finished samples).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--full]

N>1: launched by torchrun, one rank per GPU; samples shard by batch (no data-path collective), weights are packed on
rank 0 and broadcast once with NCCL.  Timing: CUDA events on the launching stream, barrier + synchronize on both
sides, max over ranks.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

LARGE_CFG = dict(image_size=128, in_channels=4, out_channels=4, model_channels=256, num_res_blocks=2, num_classes=1000,
                 has_null_class=True, channel_mult=[1, 1, 2, 3, 4], attention_resolutions=[32, 16, 8], num_groups=32,
                 num_heads=None, num_head_channels=64, dropout=0.0, use_fp16=False)   # configs/rgbd_imagenet_adm_128_large_cfg.json
GUIDANCE = 0.5            # README.md:90 evaluation protocol
DENOISE_STEPS = 1000
GFLOP_PER_FORWARD = 613.8  # per sample


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d.get("hbm_gbs", 6650.0), tf_burst=d.get("bf16_tflops", 1590.0),
                    tf_sust=d.get("bf16_tflops_sustained", 1400.0), src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tf_burst=1590.0, tf_sust=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU oracle port of the reference path (the Python reference cannot travel)
# ----------------------------------------------------------------------------------------------------------------------
def cpu_step_seconds(steps, warmup, budget_s=150.0):
    from oracle import sampler_ref, unet_ref
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    # use the thread count that is actually fastest on this host (oversubscribed SMT threads slow oneDNN down)
    probe_x = torch.randn(1, 256, 128, 128)
    probe_w = torch.randn(256, 256, 3, 3)
    best, cores = None, 1
    for nt in sorted({min(avail, c) for c in (8, 16, 32, 64, 128, avail)}):
        torch.set_num_threads(nt)
        torch.nn.functional.conv2d(probe_x, probe_w, padding=1)
        t0 = time.perf_counter()
        for _ in range(3):
            torch.nn.functional.conv2d(probe_x, probe_w, padding=1)
        dt = time.perf_counter() - t0
        if best is None or dt < best:
            best, cores = dt, nt
    torch.set_num_threads(cores)
    sd = unet_ref.make_synthetic_state_dict(LARGE_CFG, seed=1234)
    tb = sampler_ref.Tables(sampler_ref.get_betas("linear", 1000))
    model = lambda x, t, c: unet_ref.unet_forward(LARGE_CFG, sd, x, t, c)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(1, 4, 128, 128, generator=g)
    classes = torch.tensor([7])

    def one(i):
        nonlocal x
        t = torch.tensor([999 - i])
        eps = sampler_ref.cfg_eps(model, x, t, classes, GUIDANCE)
        x, _ = sampler_ref.ddpm_step(tb, x, t, eps, torch.randn(x.shape, generator=g))

    t0 = time.perf_counter(); one(0); first = time.perf_counter() - t0
    # bound the sample: fit warmup + steps into the budget
    n_warm = max(0, min(warmup, int(budget_s * 0.2 / max(first, 1e-3)) - 1))
    for i in range(n_warm):
        one(1 + i)
    n = max(1, min(steps, int(budget_s * 0.7 / max(first, 1e-3))))
    t0 = time.perf_counter()
    for i in range(n):
        one(1 + n_warm + i)
    dt = (time.perf_counter() - t0) / n
    return dt, n, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    dt, n, cores = cpu_step_seconds(args.steps, args.warmup)
    value = 1.0 / (DENOISE_STEPS * dt)
    line = {
        "impl": "reference", "metric": "128x128 RGBD multiview samples/sec", "value": value, "unit": "samples/s",
        "n_gpus": args.gpus, "steps": n, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "rgbd_imagenet_adm_128_large_cfg uncond, DDPM 1000 steps + CFG(0.5), CPU oracle port of the reference path",
                   "batch": 1, "denoise_steps_per_sample": DENOISE_STEPS, "timed_unit": "one denoising step (2 UNet forwards)"},
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": "port",
                         "sample": f"{n} DDPM+CFG denoising steps at batch 1 (fp32, torch CPU, {cores} threads)"},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------------
def _ncu_traffic():
    """DRAM bytes per launch of the dominant conv kernel from the committed `ncu --set full` capture (profiles/), or None."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "ncu_full_*_conv_summary.json")))
    if not files:
        return None, None
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    tot, n = 0.0, 0
    for k in json.load(open(files[-1])):
        try:
            b = 0.0
            for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                v, u = k[key].split()
                b += float(v.replace(",", "")) * unit[u]
            tot += b
            n += 1
        except (KeyError, ValueError):
            continue
    if n == 0:
        return None, None
    return tot / n, f"{os.path.relpath(files[-1], ROOT)}: mean of {n} captured conv_gemm launches (conv1 / conv2 of the first 128x128 ResBlock)"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="samples per GPU (BASELINE configs[1]: 16)")
    ap.add_argument("--full", action="store_true", help="one bench step = a complete 1000-step sample batch")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3) if not args.full else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback in the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist_
        dist = dist_
        dist.init_process_group("nccl", device_id=dev)

    from ivid_b200 import _lib
    import ivid_b200.backbones as backbones
    import ivid_b200.frameworks as frameworks
    import ivid_b200.samplers as samplers
    from oracle import unet_ref   # only for the synthetic weight generator + cpu_baseline leg

    net = backbones.AdmUnet2d(**LARGE_CFG)
    if rank == 0:
        net.load_state_dict(unet_ref.make_synthetic_state_dict(LARGE_CFG, seed=1234))
    net = net.cuda()
    net.repack()
    bcast_ms = None
    if world > 1:
        # weights: packed on rank 0, ONE NCCL broadcast of the device arena over NVLink (sample.py:186-195 loads per rank)
        ptr, nbytes = net.weight_arena()

        class _Arena:
            __cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3}
        arena = torch.as_tensor(_Arena(), device=dev)
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); dist.broadcast(arena, src=0); e1.record(); torch.cuda.synchronize()
        bcast_ms = e0.elapsed_time(e1)
    fw = frameworks.ClassifierFreeGuidance(net, timesteps=1000, beta_schedule="linear", p_uncond=0.1)
    sampler = samplers.DdpmSampler(fw)

    B = args.batch
    gen = torch.Generator().manual_seed(1000 + rank)
    x_host = torch.randn(B, 4, 128, 128, generator=gen).pin_memory()
    classes_host = (torch.arange(B) + rank * B) % 1000
    x = x_host.to(dev)
    classes = classes_host.to(dev)
    kw = {"strength": GUIDANCE}

    def dstep(xc, i):
        return sampler._native_step(xc, (DENOISE_STEPS - 1 - i) % DENOISE_STEPS, 0, classes, False, 0.0, kw, None, None).pred_x_prev

    def full_sample():
        return sampler.sample(B, noise=x, classes=classes, strength=GUIDANCE, verbose=False).samples

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ----
    xc = x
    for i in range(args.warmup):
        xc = full_sample() if args.full else dstep(xc, i)
    # ---- timed region ----
    clocks = ClockSampler(local) if rank == 0 else None
    barrier()
    if clocks:
        clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.steps):
        xc = full_sample() if args.full else dstep(xc, args.warmup + i)
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    clk = clocks.stop() if clocks else None
    if dist is not None:
        tt = torch.tensor([ms_total], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_total = float(tt.item())
    ms_per_step = ms_total / args.steps
    finite = bool(torch.isfinite(xc).all())
    if args.full:
        value = world * B / (ms_per_step / 1e3)
    else:
        value = world * B / (DENOISE_STEPS * ms_per_step / 1e3)

    # ---- e2e: the public API call a user makes, host buffers, H2D + D2H inside the timed region ----
    out_host = torch.empty(B, 4, 128, 128).pin_memory()
    t_host = torch.full((B,), 500, dtype=torch.int64)

    def e2e_step(i):
        xd = x_host.to(dev, non_blocking=True)
        cd = classes_host.to(dev, non_blocking=True)
        td = t_host.to(dev, non_blocking=True)
        r = sampler.sample_once(xd, td, cd, strength=GUIDANCE)
        out_host.copy_(r.pred_x_prev, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    if args.full:
        def e2e_step(i):  # noqa: F811
            r = sampler.sample(B, noise=x_host.to(dev, non_blocking=True), classes=classes_host.to(dev, non_blocking=True),
                               strength=GUIDANCE, verbose=False)
            out_host.copy_(r.samples, non_blocking=True)
            torch.cuda.current_stream().synchronize()
    for i in range(2 if not args.full else 0):
        e2e_step(i)
    barrier()
    n_e2e = args.steps
    t0 = time.perf_counter()
    for i in range(n_e2e):
        e2e_step(i)
    barrier()
    e2e_s = (time.perf_counter() - t0) / n_e2e
    if dist is not None:
        tt = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_s = float(tt.item())
    e2e_value = world * B / e2e_s if args.full else world * B / (DENOISE_STEPS * e2e_s)
    h2d = x_host.numel() * 4 + classes_host.numel() * 8 + (0 if args.full else t_host.numel() * 8)
    d2h = out_host.numel() * 4

    # ---- roofline of the dominant kernel (per-launch CUDA events inside the library, one profiled step) ----
    L = _lib.lib()
    _lib.check(L.ivid_unet_profile_begin(net._handle))
    dstep(x, 1)
    buf = ctypes.create_string_buffer(1 << 19)
    _lib.check(L.ivid_unet_profile_end(net._handle, buf, len(buf)))
    prof = json.loads(buf.value.decode())
    per_op = prof.pop("_ops", None)
    if per_op is not None and rank == 0:
        with open(os.path.join(ROOT, "gpurun_out", "per_op_profile.json"), "w") as f:
            json.dump(per_op, f)
    peaks = _peaks()
    dom = max((k for k in prof if k.startswith("conv_gemm")), key=lambda k: prof[k]["ms"])
    d = prof[dom]
    ach = d["flops"] / (d["ms"] * 1e-3) / 1e12
    step_ms_prof = sum(v["ms"] for v in prof.values())
    traffic, traffic_src = _ncu_traffic()
    roofline = {"bound": "tensor", "kernel": dom, "achieved": ach, "peak": peaks["tf_sust"], "unit": "TFLOP/s",
                "frac": ach / peaks["tf_sust"], "traffic": traffic, "traffic_unit": "DRAM bytes/launch", "traffic_source": traffic_src,
                "peak_source": peaks["src"] + ", bf16/fp16 dense sustained",
                "launches_per_step": d["launches"], "kernel_share_of_step": d["ms"] / step_ms_prof,
                "families": {k: {"launches": v["launches"], "ms": round(v["ms"], 4),
                                 "tflops": (v["flops"] / (v["ms"] * 1e-3) / 1e12) if v["ms"] > 0 else None,
                                 "gbs": (v["bytes"] / (v["ms"] * 1e-3) / 1e9) if v["ms"] > 0 else None} for k, v in prof.items()}}
    launches_per_step = sum(v["launches"] for v in prof.values()) + 3 + 3   # embed op = 4 kernels; + set_step, classes, step

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        dt, n, cores = cpu_step_seconds(3, 0, budget_s=40.0)
        cpu = {"value": 1.0 / (DENOISE_STEPS * dt), "unit": "samples/s", "cores": cores, "kind": "port",
               "sample": f"{n} DDPM+CFG denoising steps at batch 1 (CPU oracle port of the reference path, fp32, {cores} threads)"}

    line = {
        "metric": "128x128 RGBD multiview samples/sec", "value": value, "unit": "samples/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "fp16 tensor-core operands, fp32 accumulate / residual stream / GroupNorm / softmax / sampler",
        "data": "synthetic",
        "config": {"workload": "rgbd_imagenet_adm_128_large_cfg uncond, DDPM 1000 steps + classifier-free guidance 0.5 (BASELINE configs[1])",
                   "batch_per_gpu": B, "global_batch": B * world, "denoise_steps_per_sample": DENOISE_STEPS,
                   "timed_unit": "complete 1000-step sample batch" if args.full else "one denoising step of the batch (batch-2N UNet forward + fused CFG/DDPM update)",
                   "unet_step_ms": None if args.full else ms_per_step,
                   "model_tflops_per_s": (2 * B * GFLOP_PER_FORWARD / 1e3) / (ms_per_step / 1e3) if not args.full else None,
                   "parallelism": f"dp{world} (samples sharded by batch, no data-path collective)",
                   "l2": "per-step working set (0.84 GB weights + >2 GB activations) exceeds the 126 MB L2; no flush needed",
                   "weights_broadcast_ms": bcast_ms, "finite": finite},
        "roofline": roofline,
        "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "api": "DdpmSampler.sample_once(x_t, t, classes, strength) from pinned host tensors, result read back to host"
                       if not args.full else "DdpmSampler.sample(...) from pinned host noise, samples read back"},
        "gpu_launches": launches_per_step * args.steps * (DENOISE_STEPS if args.full else 1),
        "clocks": clk,
    }
    if cpu:
        line["cpu_baseline"] = cpu
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
