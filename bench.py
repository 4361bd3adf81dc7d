#!/usr/bin/env python
"""bench.py -- the cost-volume hot path on B200: PWC-style flow-decoder pyramid,
fused flow-warp + correlation + LeakyReLU forward, 1024x512 image pair (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A *step* is one pass of the hot path over one image pair's feature pyramid: 5 launches
(level 0 un-warped, levels 1-4 warped; SURVEY.md section 8 table "PWC").  Inputs rotate through a
pool of distinct buffer sets larger than 2x L2 so every step streams from HBM.

Prints ONE JSON line (rank 0): `value` = whole-job Mpix/s of image pixels with inputs resident in
HBM; `e2e` = the same through the host-buffer C ABI call (pinned host memory, H2D + D2H inside the
timed region); `roofline` for the dominant kernel (finest level); `cpu_baseline` = the pure-PyTorch
restatement of the reference op on this box's host cores (north_star asks for exactly that).

`--impl reference` times that CPU implementation alone, on the same workload and metric.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import math
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IMG_W, IMG_H = 1024, 512
# (C, H, W, warped) coarsest first -- PWC FeatureExtractor pyramid at 1024x512, SURVEY.md section 8
PWC_LEVELS = [(192, 8, 16, False), (128, 16, 32, True), (96, 32, 64, True), (64, 64, 128, True), (32, 128, 256, True)]
MD, PAD, D2 = 4, 4, 81
SLOPE = 0.1
N_SETS = 10  # 10 x 29.3 MB of inputs+outputs = 293 MB > 2 x 126 MB L2
FMA_PEAK_TFLOPS = 70.4  # measured on this pool with tools/microbench/pipes.cu (FFMA2, sustained)
# dram__bytes_read.sum + dram__bytes_write.sum of one launch of the finest-level kernel, from the
# ncu --set full capture summarised in profiles/r01_ncu_fwd_finest_level.txt (reads = algorithmic
# input bytes; the 10.6 MB of output is still dirty in L2 when the kernel ends)
NCU_TRAFFIC_BYTES_FINEST = 8930560


def level_bytes(C, H, W, warped, B=1, e=4):
    """Algorithmic HBM bytes of the fused forward: read x1, x2 (+flow), write out."""
    return B * H * W * e * (2 * C + D2 + (2 if warped else 0))


def level_flops(C, H, W, B=1):
    return 2 * B * H * W * C * D2


def synth_level(level_idx, C, H, W, warped, seed_base, device, pin=False):
    """features ~ LeakyReLU(N(0,1)), flow ~ N(0,1.5^2) clipped to +-(md+2)."""
    g = torch.Generator().manual_seed(seed_base + level_idx)
    x1 = torch.nn.functional.leaky_relu(torch.randn(1, C, H, W, generator=g), SLOPE)
    x2 = torch.nn.functional.leaky_relu(torch.randn(1, C, H, W, generator=g), SLOPE)
    fl = (torch.randn(1, 2, H, W, generator=g) * 1.5).clamp_(-(MD + 2), MD + 2) if warped else None
    if pin:
        return tuple(t.pin_memory() if t is not None else None for t in (x1, x2, fl))
    return tuple(t.to(device) if t is not None else None for t in (x1, x2, fl))


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index, period=0.01):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:  # noqa: BLE001
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(self.period)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------ CPU reference arm
def cpu_pyramid_step(levels_host, torch_oracle):
    outs = []
    for (x1, x2, fl) in levels_host:
        outs.append(torch_oracle.level_forward(x1, x2, fl, PAD, 1, MD, 1, 1, torch_oracle.WARP_TORCH, SLOPE))
    return outs


def run_cpu(steps, warmup, budget_s):
    """Pure-PyTorch restatement of the reference op on the host cores.  Returns (Mpix/s, ms/step,
    cores, sample description, steps actually timed)."""
    from oracle import torch_oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    host = [synth_level(i, C, H, W, wp, 1000, "cpu") for i, (C, H, W, wp) in enumerate(PWC_LEVELS)]
    with torch.no_grad():
        t0 = time.perf_counter()
        cpu_pyramid_step(host, torch_oracle)
        t_first = time.perf_counter() - t0
        for _ in range(max(0, min(warmup, 3) - 1)):
            cpu_pyramid_step(host, torch_oracle)
        n = max(1, min(steps, int(budget_s / max(t_first, 1e-3))))
        t0 = time.perf_counter()
        for _ in range(n):
            cpu_pyramid_step(host, torch_oracle)
        dt = (time.perf_counter() - t0) / n
    mpix = IMG_W * IMG_H / dt / 1e6
    sample = f"{n} full pyramid passes (5 levels, batch 1) of the same workload"
    return mpix, dt * 1e3, torch.get_num_threads(), sample, n


def main_reference(args, rank):
    if rank != 0:
        return 0
    mpix, ms, cores, sample, n = run_cpu(args.steps, args.warmup, budget_s=120.0)
    line = {
        "impl": "reference", "metric": "corr+warp Mpix/s", "value": mpix, "unit": "Mpix/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "pwc_pyramid_1024x512_b1_fp32", "levels_CHW": [[c, h, w] for c, h, w, _ in PWC_LEVELS],
                   "timed_steps": n},
        "cpu_baseline": {"value": mpix, "unit": "Mpix/s", "cores": cores, "kind": "port",
                         "sample": sample + " (oracle/torch_oracle.py: pure-PyTorch restatement of "
                                            "CorrelationTorch + flow_warp + leaky_relu)"},
        "e2e": {"value": mpix, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------ GPU arm
def main_gpu(args, rank, world, local_rank):
    import cerberusnet_b200 as cb
    from cerberusnet_b200 import _lib, ops

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    lib = cb.lib()  # raises if the CUDA library is missing: no fallback
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    # ---- buffer pool: N_SETS independent image pairs (seeded differently) + outputs
    sets = []
    for s in range(N_SETS):
        lv = []
        for i, (C, H, W, wp) in enumerate(PWC_LEVELS):
            x1, x2, fl = synth_level(i, C, H, W, wp, 1000 + 100 * s + 7 * rank, dev)
            out = torch.empty(1, D2, H, W, device=dev)
            lv.append((x1, x2, fl, out))
        sets.append(lv)

    def launch_level(lv, variant=0):
        x1, x2, fl, out = lv
        ops.warp_corr_forward(x1, x2, fl, PAD, 1, MD, 1, 1, 1, cb.WARP_TORCH, SLOPE, out=out, variant=variant)

    def launch_step(s):
        for lv in sets[s % N_SETS]:
            launch_level(lv)

    stream = torch.cuda.Stream(device=dev)

    def capture(fn):
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(stream):
            fn()  # eager once: sets the smem attribute, surfaces argument errors
            torch.cuda.synchronize()
            with torch.cuda.graph(g, stream=stream):
                fn()
        return g

    n0 = lib.cerb_launch_count()
    full_graph = capture(lambda: [launch_step(s) for s in range(N_SETS)])
    launches_per_step = (lib.cerb_launch_count() - n0) // (2 * N_SETS)
    rem = args.steps % N_SETS
    rem_graph = capture(lambda: [launch_step(s) for s in range(rem)]) if rem else None

    def run_steps(k):
        with torch.cuda.stream(stream):
            for _ in range(k // N_SETS):
                full_graph.replay()
            if k % N_SETS:
                if k % N_SETS == rem and rem_graph is not None:
                    rem_graph.replay()
                else:
                    for s in range(k % N_SETS):
                        launch_step(s)

    # ---- spin the clocks up (the GPU idles at a few hundred MHz), then W untimed warm-up steps
    t_end = time.perf_counter() + 1.0
    while time.perf_counter() < t_end:
        run_steps(50 * N_SETS)
        torch.cuda.synchronize()
    run_steps(max(args.warmup, 3))
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()

    # ---- timed region: exactly K steps
    sampler = ClockSampler(local_rank)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    sampler.start()
    with torch.cuda.stream(stream):
        e0.record(stream)
    run_steps(args.steps)
    with torch.cuda.stream(stream):
        e1.record(stream)
    torch.cuda.synchronize()
    barrier()
    clocks = sampler.stop()
    ms_total = e0.elapsed_time(e1)
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = world * IMG_W * IMG_H / (ms_per_step * 1e-3) / 1e6

    # ---- per-level launch durations (rotating sets, back to back in one graph, same warm state)
    hbm_peak, peak_src = load_peaks()
    level_stats = []
    reps = 20
    for li, (C, H, W, wp) in enumerate(PWC_LEVELS):
        g = capture(lambda: [launch_level(sets[s % N_SETS][li]) for s in range(N_SETS * 2)])
        samples = []
        for _ in range(5):   # median of 5 windows of 400 launches: one window is ~8 ms, short enough to catch a transient
            with torch.cuda.stream(stream):
                g.replay()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                for _ in range(reps):
                    g.replay()
                b.record(stream)
            torch.cuda.synchronize()
            samples.append(a.elapsed_time(b) * 1e3 / (reps * N_SETS * 2))
        us = float(np.median(samples))
        byts, fl = level_bytes(C, H, W, wp), level_flops(C, H, W)
        level_stats.append({"level": li, "C": C, "H": H, "W": W, "warped": wp, "us_per_launch": round(us, 3),
                            "algorithmic_MB": round(byts / 1e6, 3), "GBps": round(byts / us / 1e3, 1),
                            "hbm_frac": round(byts / us / 1e3 / hbm_peak, 4),
                            "fp32_TFLOPs": round(fl / us / 1e6, 2),
                            "fma_frac": round(fl / us / 1e6 / FMA_PEAK_TFLOPS, 4)})
    dom = max(level_stats, key=lambda d: d["algorithmic_MB"])  # the finest level carries 65% of the bytes
    t_hbm = dom["algorithmic_MB"] * 1e6 / (hbm_peak * 1e9)
    t_fma = level_flops(dom["C"], dom["H"], dom["W"]) / (FMA_PEAK_TFLOPS * 1e12)
    roofline = {
        "bound": "hbm", "achieved": dom["GBps"], "peak": hbm_peak, "unit": "GB/s", "frac": dom["hbm_frac"],
        "traffic": NCU_TRAFFIC_BYTES_FINEST, "traffic_source": "profiles/r01_ncu_fwd_finest_level.txt", "peak_source": peak_src,
        "kernel": f"warp_corr_fwd_kernel<float,8,32,1,4,3> level {dom['level']} (C={dom['C']}, {dom['H']}x{dom['W']})",
        "algorithmic_bytes_per_launch": int(dom["algorithmic_MB"] * 1e6),
        "avg_launch_us": dom["us_per_launch"], "avg_launch_us_method": "median of 5 windows of 400 back-to-back launches (CUDA events on the launching stream)",
        "binding_roof_frac": round(max(t_hbm, t_fma) * 1e6 / dom["us_per_launch"], 4),
        "fma_peak_tflops": FMA_PEAK_TFLOPS, "fma_frac": dom["fma_frac"],
        "sum_level_us": round(sum(d["us_per_launch"] for d in level_stats), 3),
    }

    # ---- end to end with HOST buffers (pinned): every step copies its inputs host->device and its
    # results device->host inside the timed region.  Two public entry points are timed:
    #   (a) cerb_warp_corr_forward_host: the C ABI call, H2D + fused kernel + D2H on one stream;
    #   (b) cerberusnet_b200.HostPipeline: same work on three streams over double-buffered device
    #       staging, so the two PCIe directions and the kernels overlap.  (b) is reported as `e2e`.
    from cerberusnet_b200.host_pipeline import HostPipeline
    host_sets = []
    for s in range(2):
        lv = []
        for i, (C, H, W, wp) in enumerate(PWC_LEVELS):
            x1, x2, fl = synth_level(i, C, H, W, wp, 5000 + 100 * s + 7 * rank, "cpu", pin=True)
            out = torch.empty(1, D2, H, W).pin_memory()
            lv.append((x1, x2, fl, out))
        host_sets.append(lv)
    params, ws_bytes = [], 0
    for (C, H, W, wp), (x1, x2, fl, out) in zip(PWC_LEVELS, host_sets[0]):
        p = _lib.make_params(x1, x2, fl, out, PAD, 1, MD, 1, 1, 1, cb.WARP_TORCH, SLOPE)
        for name in ("x1_stride", "x2_stride", "flow_stride", "out_stride"):
            setattr(p, name, (ctypes.c_int64 * 4)(0, 0, 0, 0))
        params.append(p)
        ws_bytes = max(ws_bytes, lib.cerb_warp_corr_forward_host_workspace(ctypes.byref(p), 1 if wp else 0))
    wss = [torch.empty(ws_bytes, dtype=torch.uint8, device=dev) for _ in PWC_LEVELS]
    h2d = sum(t.numel() * 4 for (x1, x2, fl, _) in host_sets[0] for t in (x1, x2, fl) if t is not None)
    d2h = sum(out.numel() * 4 for (_, _, _, out) in host_sets[0])

    def abi_step(s):
        sp = ctypes.c_void_p(stream.cuda_stream)
        for li, (x1, x2, fl, out) in enumerate(host_sets[s % 2]):
            rc = lib.cerb_warp_corr_forward_host(ctypes.byref(params[li]), _lib.ptr(x1), _lib.ptr(x2),
                                                 _lib.ptr(fl), _lib.ptr(out), _lib.ptr(wss[li]), ws_bytes, sp)
            _lib.check(rc, "cerb_warp_corr_forward_host")

    pipe = HostPipeline(PWC_LEVELS, batch=1, depth=2, device=dev, pad_size=PAD, max_displacement=MD,
                        warp_mode=cb.WARP_TORCH, leaky_slope=SLOPE)

    pipe.enable_arenas()   # one pinned arena per slot and direction: one memcpy each way per step
    for s in range(2):
        for (hx1, hx2, hfl), (x1, x2, fl, _) in zip(pipe.host_inputs(s), host_sets[s]):
            hx1.copy_(x1); hx2.copy_(x2)
            if hfl is not None:
                hfl.copy_(fl)

    def pipe_step(s):
        pipe.submit_packed(s % 2)

    def time_host_path(step_fn, sync_fn, k):
        for s in range(3):
            step_fn(s)
        sync_fn()
        barrier()
        t0 = time.perf_counter()
        for s in range(k):
            step_fn(s)
        sync_fn()                       # results are in host memory
        ms = (time.perf_counter() - t0) * 1e3 / k
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    k_e2e = max(3, min(args.steps, 400))
    ms_abi = time_host_path(abi_step, torch.cuda.synchronize, k_e2e)
    ms_pipe = time_host_path(pipe_step, pipe.synchronize, k_e2e)
    # the pipeline's result must be the kernel's result
    ref_out = ops.warp_corr_forward(*[t.to(dev) for t in host_sets[(k_e2e - 1) % 2][4][:3]], PAD, 1, MD, 1, 1, 1,
                                    cb.WARP_TORCH, SLOPE)
    if not torch.equal(ref_out.cpu(), pipe.host_outputs((k_e2e - 1) % 2)[4]):
        raise RuntimeError("HostPipeline output differs from the device-resident call")
    e2e = {"value": world * IMG_W * IMG_H / (ms_pipe * 1e-3) / 1e6, "unit": "Mpix/s",
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_pipe, "steps": k_e2e,
           "api": "cerberusnet_b200.HostPipeline.submit_packed (pinned host arenas; one H2D copy, the fused kernels and "
                  "one D2H copy per step on three streams, double-buffered); timed host-side until the results "
                  "are in host memory",
           "single_stream_c_abi": {"api": "cerb_warp_corr_forward_host", "ms_per_step": ms_abi,
                                   "value": world * IMG_W * IMG_H / (ms_abi * 1e-3) / 1e6}}

    # ---- CPU baseline beside it (rank 0, N=1 only)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        mpix, ms, cores, sample, n = run_cpu(10_000, 1, budget_s=15.0)
        cpu_baseline = {"value": mpix, "unit": "Mpix/s", "cores": cores, "kind": "port",
                        "sample": sample + ", pure-PyTorch restatement (oracle/torch_oracle.py), fp32",
                        "ms_per_step": ms}

    if rank == 0:
        line = {
            "metric": "corr+warp Mpix/s", "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "pwc_pyramid_1024x512_b1_fp32",
                       "levels_CHW": [[c, h, w] for c, h, w, _ in PWC_LEVELS],
                       "op": "fused flow_warp(mode torch) + correlation(pad4,k1,md4,s1,s2=1) + LeakyReLU(0.1) forward",
                       "launches_per_step": int(launches_per_step),
                       "l2": f"inputs/outputs rotate through {N_SETS} distinct sets (293 MB > 2x L2), no flush",
                       "per_gpu": "one image pair per step per GPU; ranks run independent pairs (no collective)"},
            "roofline": roofline, "levels": level_stats, "cpu_baseline": cpu_baseline, "e2e": e2e,
            "gpu_launches": int(launches_per_step) * args.steps, "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20000)
    ap.add_argument("--warmup", type=int, default=2000)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return main_reference(args, rank)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback "
                         "(use --impl reference for the CPU arm)")
    return main_gpu(args, rank, world, local_rank)


if __name__ == "__main__":
    sys.exit(main())
