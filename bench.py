#!/usr/bin/env python
"""bench.py -- the cost-volume hot path on B200: PWC-style flow-decoder pyramid, fused flow-warp +
correlation + LeakyReLU, 1024x512 image pair (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

A *step* is one pass of the hot path over one image pair's feature pyramid, BOTH flow directions -- what the
reference's forward computes by default (`consistency=True`: nnet_models/pwcnet.py:108-113, cerberus.py:131-135).
One launch per level covers both directions: the level's features of [image 1; image 2] are one (2,C,H,W) tensor
that serves as both correlation inputs (`x2_batch_roll = 1`), so a step is 5 launches (level 0 un-warped, levels
1-4 warped; SURVEY.md section 8 table "PWC").  Inputs rotate through a pool of distinct buffer sets larger than
2x L2, so every step streams from HBM.

Prints ONE JSON line (rank 0):
  value        whole-job Mpix/s of image pixels per flow direction (2 x 1024 x 512 per step), inputs resident in HBM;
  e2e          the same through cerberusnet_b200.HostPipeline with pinned HOST buffers (H2D + kernels + D2H timed);
  roofline     dominant kernel (finest level, both directions) against the measured HBM peak; roofline_bwd for the
               backward of the same level and of the HRNet training shape;
  levels / levels_smooth   per-launch device time per level for the iid flow model of SURVEY 8d and for a
               decoder-like smooth flow (x2 up-sampled coarse flow, sigma 3 px), with the staging-path mix;
  train        DDP training step (flow network on the fused op + parameter ballast sized like CerberusBase's
               287 MB gradient all-reduce), batch 8 per GPU -- the split north_star names for multi-GPU;
  cpu_baseline the pure-PyTorch restatement of the reference op on this box's host cores (N=1 only).

`--impl reference` times that CPU implementation alone on the same workload, metric and config.
"""
from __future__ import annotations

import argparse
import ctypes
import json
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
HRNET_TRAIN_LEVEL = (48, 128, 256, 8)   # C, H, W, batch: finest HRNetV2-W48 level at batch 8 (BASELINE configs[3])
MD, PAD, D2 = 4, 4, 81
SLOPE = 0.1
DIRS = 2      # flow directions per step (forward and backward flow of the pair)
N_SETS = 6    # 6 x 58.7 MB of inputs+outputs = 352 MB > 2 x 126 MB L2
# dram__bytes_read.sum + dram__bytes_write.sum of one launch of the finest-level kernel (both directions), from this
# round's `ncu --set full` capture summarised in profiles/r02_ncu_fwd_finest_level.txt
NCU_TRAFFIC_BYTES_FINEST = 9004288   # 9.003 MB read + 1 KB written (the 21.2 MB of output is still dirty in L2 at kernel end)
NCU_TRAFFIC_SOURCE = "profiles/r02_ncu_fwd_finest_level.txt"


def make_config():
    """The workload description -- identical in both arms (the driver compares the dicts)."""
    return {"workload": "pwc_pyramid_1024x512_b1_fp32",
            "levels_CHW": [[c, h, w] for c, h, w, _ in PWC_LEVELS],
            "flow_directions": DIRS,
            "op": "fused flow_warp(mode torch) + correlation(pad4,k1,md4,s1,s2=1) + LeakyReLU(0.1) forward, "
                  "forward and backward flow of one image pair per step (reference default consistency=True)",
            "flow_model": "iid N(0,1.5^2) px clipped to +-6 per pixel (SURVEY 8d)",
            "l2": f"inputs/outputs rotate through {N_SETS} distinct sets (352 MB > 2x L2), no flush",
            "per_gpu": "one image pair per step per GPU; ranks run independent pairs (no collective on the path)"}


def level_bytes(C, H, W, warped, B=1, e=4):
    """Algorithmic HBM bytes of the fused forward per direction: read x1, x2 (+flow), write out."""
    return B * H * W * e * (2 * C + D2 + (2 if warped else 0))


def level_bytes_bwd(C, H, W, warped, B=1, e=4):
    """bytes = B*H*W*e*(D2 + 4C [+ D2 for the saved activation, + 4 for flow and g_flow])."""
    return B * H * W * e * (2 * D2 + 4 * C + (4 if warped else 0))


def level_flops(C, H, W, B=1):
    return 2 * B * H * W * C * D2


def synth_flow(kind, B, H, W, g):
    if kind == "iid":
        return (torch.randn(B, 2, H, W, generator=g) * 1.5).clamp_(-(MD + 2), MD + 2)
    # what a decoder produces: the next-coarser flow, doubled and bilinearly up-sampled (sigma = 3 px)
    coarse = (torch.randn(B, 2, H // 2, W // 2, generator=g) * 1.5).clamp_(-(MD + 2), MD + 2)
    return torch.nn.functional.interpolate(coarse * 2, scale_factor=2, mode="bilinear", align_corners=True)


def synth_level(level_idx, C, H, W, warped, seed_base, device, pin=False, flow_kind="iid", B=DIRS):
    """Features of [image 1; image 2] ~ LeakyReLU(N(0,1)) (post-activation, pwcnet_modules.py:13) and the
    [forward; backward] flows."""
    g = torch.Generator().manual_seed(seed_base + level_idx)
    f = torch.nn.functional.leaky_relu(torch.randn(B, C, H, W, generator=g), SLOPE)
    fl = synth_flow(flow_kind, B, H, W, g) if warped else None
    if pin:
        return tuple(t.pin_memory() if t is not None else None for t in (f, fl))
    return tuple(t.to(device) if t is not None else None for t in (f, fl))


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
    """Both flow directions of one pair with the pure-PyTorch restatement: direction 1 swaps the feature maps."""
    outs = []
    for (f, fl) in levels_host:
        for d in range(DIRS):
            x1, x2 = f[d:d + 1], f[(d + 1) % DIRS:(d + 1) % DIRS + 1]
            outs.append(torch_oracle.level_forward(x1, x2, None if fl is None else fl[d:d + 1], PAD, 1, MD, 1, 1,
                                                   torch_oracle.WARP_TORCH, SLOPE))
    return outs


def run_cpu(steps, warmup, budget_s):
    """Pure-PyTorch restatement of the reference op on the host cores.  Returns (Mpix/s, ms/step, cores, sample
    description, steps actually timed)."""
    from oracle import torch_oracle
    cores = os.cpu_count() or 1
    try:
        cores = min(cores, len(os.sched_getaffinity(0)))
    except AttributeError:
        pass
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
    mpix = DIRS * IMG_W * IMG_H / dt / 1e6
    sample = f"{n} full steps (5 levels x {DIRS} flow directions, one image pair) of the same workload"
    return mpix, dt * 1e3, torch.get_num_threads(), sample, n


def main_reference(args, rank):
    if rank != 0:
        return 0
    mpix, ms, cores, sample, n = run_cpu(args.steps, args.warmup, budget_s=120.0)
    line = {
        "impl": "reference", "metric": "corr+warp Mpix/s", "value": mpix, "unit": "Mpix/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": make_config(), "timed_steps": n,
        "cpu_baseline": {"value": mpix, "unit": "Mpix/s", "cores": cores, "kind": "port",
                         "sample": sample + " (oracle/torch_oracle.py: pure-PyTorch restatement of CorrelationTorch + "
                                            "flow_warp + leaky_relu; the reference's own Python cannot travel to the box)"},
        "e2e": {"value": mpix, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------ training block (DDP)
def run_train_block(dev, rank, world, local_rank, steps=8, warmup=3, batch=8):
    """BASELINE configs[3]: a flow-network training step on the fused op, forward and backward flow, Adam, batch 8
    per GPU of 1024x512 pairs, DistributedDataParallel over NCCL.  The flow network is this repo's harness
    (encoder + PWC-style decoder, 2.2 M parameters); the reference's CerberusBase has 71.72 M parameters
    (cerberus.py:88-146), so a parameter ballast brings the bucketed gradient all-reduce to the same ~287 MB --
    the ballast takes part in the graph (its gradient is produced by autograd and reduced by DDP) but adds no math."""
    import torch.distributed as dist
    import cerberusnet_b200 as cb
    from cerberusnet_b200.decoder import FlowNetLite, photometric_loss

    target_params = 71_720_000

    class Ballasted(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.net = FlowNetLite()
            have = sum(p.numel() for p in self.net.parameters())
            n = max(0, target_params - have)
            # 16 tensors, so DDP spreads them over its 25 MB buckets and overlaps the reduction with backward
            self.ballast = torch.nn.ParameterList(torch.nn.Parameter(torch.zeros(n // 16)) for _ in range(16))

        def forward(self, a, b):
            out = self.net(a, b, consistency=True)
            z = sum(p[0] for p in self.ballast) * 0.0      # every ballast tensor gets a (zero) gradient
            out["flow"] = [f + z for f in out["flow"]]
            return out

    torch.manual_seed(0)
    model = Ballasted().to(dev)
    n_params = sum(p.numel() for p in model.parameters())
    net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank]) if world > 1 else model
    opt = torch.optim.Adam(net.parameters(), lr=1e-4, betas=(0.9, 0.99), weight_decay=1e-6)
    g = torch.Generator().manual_seed(4321 + rank)
    host = [(torch.rand(batch, 3, IMG_H, IMG_W, generator=g).pin_memory(),
             torch.rand(batch, 3, IMG_H, IMG_W, generator=g).pin_memory()) for _ in range(2)]
    lib = cb.lib()
    op_ms = [0.0]

    def step(i, time_op=False):
        h1, h2 = host[i % 2]
        img1, img2 = h1.to(dev, non_blocking=True), h2.to(dev, non_blocking=True)   # H2D inside the step
        out = net(img1, img2)
        loss = photometric_loss(img1, img2, out["flow"]) + photometric_loss(img2, img1, out["flow_b"])
        opt.zero_grad(set_to_none=True)
        loss.backward()
        opt.step()
        return loss

    for i in range(warmup):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    n0 = lib.cerb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        loss = step(i)
    last = float(loss.item())            # D2H of the step result
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    launches = int(lib.cerb_launch_count() - n0)

    # share of the step spent in the hot path: the op's fwd+bwd on the decoder's level shapes at this batch, both
    # directions, timed alone on the device
    from cerberusnet_b200 import ops
    op_us = 0.0
    for (C, H, W, wp) in PWC_LEVELS:
        f1 = torch.randn(batch, C, H, W, device=dev)
        f2 = torch.randn(batch, C, H, W, device=dev)
        fl = torch.randn(batch, 2, H, W, device=dev) if wp else None
        go = torch.randn(batch, D2, H, W, device=dev)
        for _ in range(2):
            o = ops.warp_corr_forward(f1, f2, fl, PAD, 1, MD, 1, 1, 1, cb.WARP_TORCH, SLOPE)
            ops.warp_corr_backward(f1, f2, fl, o, go, PAD, 1, MD, 1, 1, 1, cb.WARP_TORCH, SLOPE)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3):
            o = ops.warp_corr_forward(f1, f2, fl, PAD, 1, MD, 1, 1, 1, cb.WARP_TORCH, SLOPE)
            ops.warp_corr_backward(f1, f2, fl, o, go, PAD, 1, MD, 1, 1, 1, cb.WARP_TORCH, SLOPE)
        b.record()
        torch.cuda.synchronize()
        op_us += a.elapsed_time(b) * 1e3 / 3
    del model, net, opt
    torch.cuda.empty_cache()
    return {"metric": "flow-network training image pairs/s (whole job)", "value": world * batch / (ms * 1e-3),
            "unit": "pairs/s", "ms_per_step": ms, "steps": steps, "batch_per_gpu": batch, "params": n_params,
            "allreduce_bytes_per_step": n_params * 4 if world > 1 else 0, "parallelism": f"ddp{world}",
            "model": "PyramidEncoder + FlowDecoder on the fused op (2.2 M params) + parameter ballast to CerberusBase's "
                     "71.72 M (gradient all-reduce ~287 MB); unFlowLoss stand-in (L1 photometric + smoothness), Adam",
            "corr_warp_fwd_bwd_ms_per_step": 2 * op_us * 1e-3, "corr_warp_share_of_step": 2 * op_us * 1e-3 / ms,
            "costvolume_launches_per_step": launches / steps, "h2d_bytes_per_step": 2 * batch * 3 * IMG_H * IMG_W * 4,
            "last_loss": last}


# ------------------------------------------------------------------------------ BASELINE configs[2] and [4]
def run_other_configs(dev, lib, fma_peak, hbm_peak):
    """The other BASELINE configurations, in the driver-run record (rank 0, N=1):
    configs[2] -- the part that runs without TensorRT: the plugin-`enqueue`-shaped entries on the HRNetV2-W48 level
                  shapes at 1024x512, batch 1, kFLOAT and kHALF, CUDA-graph replays (no syncs, no private streams);
    configs[4] -- KITTI-shaped 1248x384 (HRNet pyramid, incl. the unaligned W = 78 level) and 1280x384 (PWC) finest
                  levels at max_displacement 8 (289 planes), batch 32, forward and backward."""
    import cerberusnet_b200 as cb
    from cerberusnet_b200 import _lib, ops
    F = torch.nn.functional
    f = _lib.TrtCorrFields()
    lib.cerb_trt_corr_default_fields(ctypes.byref(f))

    def graph_time(fn, reps=100):
        st = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(st):
            sp = ctypes.c_void_p(st.cuda_stream)
            fn(sp)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=st):
                for _ in range(10):
                    fn(sp)
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for _ in range(reps // 10):
                g.replay()
            e1.record(st)
            torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / reps

    trt = {"levels": [], "note": "plugin enqueue entries (cerb_trt_corr_enqueue / cerb_trt_warp_corr_enqueue, TensorRT "
                                 "un-normalise), HRNetV2-W48 level shapes, batch 1, CUDA-graph replays; TensorRT itself is "
                                 "not in the image, so this is the plugin half of configs[2] only"}
    tot = {"kFLOAT": 0.0, "kHALF": 0.0}
    for (C, H, W, wp) in ((384, 16, 32, False), (192, 32, 64, True), (96, 64, 128, True), (48, 128, 256, True)):
        row = {"C": C, "H": H, "W": W}
        for name, dt, code in (("kFLOAT", torch.float32, 0), ("kHALF", torch.float16, 1)):
            x1 = F.leaky_relu(torch.randn(1, C, H, W, device=dev), 0.1).to(dt)
            x2 = F.leaky_relu(torch.randn(1, C, H, W, device=dev), 0.1).to(dt)
            fl = (torch.randn(1, 2, H, W, device=dev) * 1.5).clamp_(-6, 6)
            out = torch.empty(1, 81, H, W, device=dev, dtype=dt)
            descs = (_lib.TrtTensorDesc * 4)()
            for i, dims in enumerate(((1, C, H, W), (1, C, H, W), (1, 2, H, W), (1, 81, H, W))):
                descs[i].dims.nbDims = 4
                for j, v in enumerate(dims):
                    descs[i].dims.d[j] = v
                descs[i].type = code if i != 2 else 0
            ins = (ctypes.c_void_p * 3)(x1.data_ptr(), x2.data_ptr(), fl.data_ptr())
            outs = (ctypes.c_void_p * 1)(out.data_ptr())

            def plain(sp):
                assert lib.cerb_trt_corr_enqueue(ctypes.byref(f), descs, ctypes.byref(descs[3]), ins, outs, None, sp) == 0

            def fused(sp):
                assert lib.cerb_trt_warp_corr_enqueue(ctypes.byref(f), 1, 0.1, descs, ctypes.byref(descs[3]), ins, outs, None, sp) == 0

            row[name + "_corr_node_us"] = round(graph_time(plain), 2)
            if wp:
                row[name + "_fused_node_us"] = round(graph_time(fused), 2)
            tot[name] += row[name + ("_fused_node_us" if wp else "_corr_node_us")]
        trt["levels"].append(row)
    trt["pyramid_us"] = {k: round(v, 1) for k, v in tot.items()}

    def timeit(fn, reps):
        fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / reps

    kitti = []
    for (C, H, W, B) in ((32, 96, 320, 32), (48, 96, 312, 32), (192, 24, 78, 32)):
        x1 = F.leaky_relu(torch.randn(B, C, H, W, device=dev), 0.1)
        x2 = F.leaky_relu(torch.randn(B, C, H, W, device=dev), 0.1)
        fl = (torch.randn(B, 2, H, W, device=dev) * 1.5).clamp_(-6, 6)
        out = torch.empty(B, 289, H, W, device=dev)
        tf = timeit(lambda: ops.warp_corr_forward(x1, x2, fl, 8, 1, 8, 1, 1, 1, cb.WARP_TORCH, SLOPE, out=out), 5)
        gg = torch.randn_like(out)
        tb = timeit(lambda: ops.warp_corr_backward(x1, x2, fl, out, gg, 8, 1, 8, 1, 1, 1, cb.WARP_TORCH, SLOPE), 3)
        flops, byts = 2 * B * H * W * C * 289, B * H * W * 4 * (2 * C + 289 + 2)
        bbytes = B * H * W * 4 * (2 * 289 + 4 * C + 4)
        kitti.append({"C": C, "H": H, "W": W, "batch": B, "max_displacement": 8, "fwd_us": round(tf, 1),
                      "fwd_TFLOPs": round(flops / tf / 1e6, 2), "fwd_fma_frac": round(flops / tf / 1e6 / fma_peak, 4),
                      "fwd_hbm_frac": round(byts / tf / 1e3 / hbm_peak, 4), "bwd_us": round(tb, 1),
                      "bwd_TFLOPs": round(2 * flops / tb / 1e6, 2), "bwd_hbm_frac": round(bbytes / tb / 1e3 / hbm_peak, 4),
                      "tma_rows_aligned": W % 4 == 0})
        del x1, x2, fl, out, gg
    torch.cuda.empty_cache()
    return {"configs2_trt_plugin_entries": trt, "configs4_kitti_md8": kitti}


# ------------------------------------------------------------------------------ GPU arm
def main_gpu(args, rank, world, local_rank):
    from cerberusnet_b200.parallel import bind_to_gpu_numa
    numa = bind_to_gpu_numa(local_rank, local_rank, int(os.environ.get("LOCAL_WORLD_SIZE", world)))
    import cerberusnet_b200 as cb
    from cerberusnet_b200 import _lib, ops

    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    lib = cb.lib()  # raises if the CUDA library is missing: no fallback
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()

    def max_ranks(v):
        if world > 1:
            import torch.distributed as dist
            t = torch.tensor([v], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return float(v)

    # ---- buffer pool: N_SETS independent image pairs (seeded differently) + outputs
    def make_sets(flow_kind, n_sets, levels=PWC_LEVELS, B=DIRS, seed=1000):
        sets = []
        for s in range(n_sets):
            lv = []
            for i, (C, H, W, wp) in enumerate(levels):
                f, fl = synth_level(i, C, H, W, wp, seed + 100 * s + 7 * rank, dev, flow_kind=flow_kind, B=B)
                lv.append((f, fl, torch.empty(B, D2, H, W, device=dev)))
            sets.append(lv)
        return sets

    sets = make_sets("iid", N_SETS)

    def launch_level(lv, roll=DIRS // 2):
        f, fl, out = lv
        ops.warp_corr_forward(f, f, fl, PAD, 1, MD, 1, 1, 1, cb.WARP_TORCH, SLOPE, out=out, x2_roll=roll)

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
    ms_per_step = max_ranks(e0.elapsed_time(e1)) / args.steps
    value = world * DIRS * IMG_W * IMG_H / (ms_per_step * 1e-3) / 1e6

    # ---- measured roofs
    hbm_peak, peak_src = load_peaks()
    tf = ctypes.c_double(0.0)
    _lib.check(lib.cerb_measure_fma_peak(ctypes.byref(tf), ctypes.c_void_p(stream.cuda_stream)), "cerb_measure_fma_peak")
    fma_peak = float(tf.value)

    def time_launches(fn_graph_body, n_launches, reps=20, windows=5):
        g = capture(fn_graph_body)
        samples = []
        for _ in range(windows):   # median of short windows: short enough to catch a transient
            with torch.cuda.stream(stream):
                g.replay()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                for _ in range(reps):
                    g.replay()
                b.record(stream)
            torch.cuda.synchronize()
            samples.append(a.elapsed_time(b) * 1e3 / (reps * n_launches))
        return float(np.median(samples))

    ctr = torch.zeros(4, dtype=torch.int64, device=dev)

    def level_table(the_sets, roll=DIRS // 2, B=DIRS):
        stats = []
        for li, (C, H, W, wp) in enumerate(PWC_LEVELS):
            ns = len(the_sets)
            us = time_launches(lambda: [launch_level(the_sets[s % ns][li], roll) for s in range(ns * 2)], ns * 2)
            # staging-path mix of the warped gather (one eager launch with the counters on)
            ctr.zero_()
            lib.cerb_debug_set_path_counters(ctypes.c_void_p(ctr.data_ptr()))
            with torch.cuda.stream(stream):
                launch_level(the_sets[0][li], roll)
            torch.cuda.synchronize()
            lib.cerb_debug_set_path_counters(None)
            c = ctr.tolist()
            tiles = max(1, c[1] + c[2] + c[3])
            byts, fl = level_bytes(C, H, W, wp, B), level_flops(C, H, W, B)
            stats.append({"level": li, "C": C, "H": H, "W": W, "warped": wp, "directions": B,
                          "us_per_launch": round(us, 3), "algorithmic_MB": round(byts / 1e6, 3),
                          "GBps": round(byts / us / 1e3, 1), "hbm_frac": round(byts / us / 1e3 / hbm_peak, 4),
                          "fp32_TFLOPs": round(fl / us / 1e6, 2), "fma_frac": round(fl / us / 1e6 / fma_peak, 4),
                          "kernel": "tensor-core (tcgen05, 8x16 tiles)" if (wp and B * ((H + 7) // 8) * ((W + 15) // 16) >= 64) else "CUDA-core (cluster-split 4x16 / 8x16 tiles)",
                          "tiles_small_box": c[1], "tiles_large_box": c[3], "tiles_fallback_direct_gather": c[2],
                          "fallback_rate": round(c[2] / tiles, 4) if wp else 0.0})
        return stats

    level_stats = level_table(sets)
    smooth_sets = make_sets("smooth", N_SETS, seed=3000)
    level_stats_smooth = level_table(smooth_sets)
    del smooth_sets
    # one flow direction per launch (the round-1 definition of the step), for continuity
    one_sets = make_sets("iid", N_SETS, B=1, seed=1000)
    level_stats_one = level_table(one_sets, roll=0, B=1)
    del one_sets

    dom = max(level_stats, key=lambda d: d["algorithmic_MB"])  # the finest level carries 65% of the bytes
    t_hbm = dom["algorithmic_MB"] * 1e6 / (hbm_peak * 1e9)
    t_fma = level_flops(dom["C"], dom["H"], dom["W"], DIRS) / (fma_peak * 1e12)
    roofline = {
        "bound": "hbm", "achieved": dom["GBps"], "peak": hbm_peak, "unit": "GB/s", "frac": dom["hbm_frac"],
        "traffic": NCU_TRAFFIC_BYTES_FINEST, "traffic_source": NCU_TRAFFIC_SOURCE, "peak_source": peak_src,
        "kernel": f"warp_corr_fwd_tc_kernel<float> (tcgen05.mma.kind::tf32 into TMEM, 3xTF32 split of the fp32 operands; 8x16 tiles "
                  f"as 128x384 Gram tiles) level {dom['level']} (C={dom['C']}, {dom['H']}x{dom['W']}, {DIRS} flow directions per launch)",
        "tensor_pipe": {"tiles_per_launch": DIRS * ((dom["H"] + 7) // 8) * ((dom["W"] + 15) // 16),
                        "mma_per_tile": 6 * ((dom["C"] + 7) // 8), "mma_shape": "M128 N192 K8 tf32",
                        "issued_tf32_TFLOPs": round(DIRS * ((dom["H"] + 7) // 8) * ((dom["W"] + 15) // 16) * 6 * ((dom["C"] + 7) // 8)
                                                    * 2 * 128 * 192 * 8 / dom["us_per_launch"] / 1e6, 1),
                        "useful_fraction_of_gram_tile": round(81 / 384, 3),
                        "note": "not the binding roof: shared-memory bandwidth (tap gather + operand stores + MMA operand reads + raw-box TMA writes) is, see DESIGN 4.1b"},
        "algorithmic_bytes_per_launch": int(dom["algorithmic_MB"] * 1e6),
        "algorithmic_bytes_formula": "directions * H*W*4*(2C + 81 + 2) (SURVEY 8d per direction)",
        "avg_launch_us": dom["us_per_launch"],
        "avg_launch_us_method": "median of 5 windows of back-to-back launches over rotating buffer sets (CUDA events on the launching stream)",
        "binding_roof_frac": round(max(t_hbm, t_fma) * 1e6 / dom["us_per_launch"], 4),
        "fma_peak_tflops": round(fma_peak, 2), "fma_peak_source": "measured in this run (cerb_measure_fma_peak: FFMA2 loop on every SM)",
        "fma_frac": dom["fma_frac"],
        "sum_level_us": round(sum(d["us_per_launch"] for d in level_stats), 3),
        "sum_level_us_per_direction": round(sum(d["us_per_launch"] for d in level_stats) / DIRS, 3),
        "one_direction_per_launch": {"sum_level_us": round(sum(d["us_per_launch"] for d in level_stats_one), 3),
                                     "finest_us": level_stats_one[-1]["us_per_launch"],
                                     "finest_hbm_frac": level_stats_one[-1]["hbm_frac"]},
        "smooth_flow_finest_us": level_stats_smooth[-1]["us_per_launch"],
    }

    # ---- backward of the fused level op: dominant level of this workload and the HRNet training shape
    def time_backward(C, H, W, B, roll, kernel=-1):
        lib.cerb_debug_set_backward_kernel(kernel)   # -1 automatic (what the product runs), 1 tensor-core kernel + window splat
        nset = max(2, int(300e6 // level_bytes_bwd(C, H, W, True, B)) + 1)
        bs = []
        for s in range(nset):
            g = torch.Generator(device=dev).manual_seed(77 + s)
            f1 = torch.nn.functional.leaky_relu(torch.randn(B, C, H, W, device=dev, generator=g), SLOPE)
            f2 = f1 if roll else torch.nn.functional.leaky_relu(torch.randn(B, C, H, W, device=dev, generator=g), SLOPE)
            fl = (torch.randn(B, 2, H, W, device=dev, generator=g) * 1.5).clamp_(-6, 6)
            out = ops.warp_corr_forward(f1, f2, fl, PAD, 1, MD, 1, 1, 1, cb.WARP_TORCH, SLOPE, x2_roll=roll)
            go = torch.randn(B, D2, H, W, device=dev, generator=g)
            bs.append((f1, f2, fl, out, go))
        n_before = lib.cerb_launch_count()
        with torch.cuda.stream(stream):
            for (f1, f2, fl, out, go) in bs:
                ops.warp_corr_backward(f1, f2, fl, out, go, PAD, 1, MD, 1, 1, 1, cb.WARP_TORCH, SLOPE, x2_roll=roll)
            torch.cuda.synchronize()
            kern = (lib.cerb_launch_count() - n_before) // nset
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3
            a.record(stream)
            for _ in range(reps):
                for (f1, f2, fl, out, go) in bs:
                    ops.warp_corr_backward(f1, f2, fl, out, go, PAD, 1, MD, 1, 1, 1, cb.WARP_TORCH, SLOPE, x2_roll=roll)
            b.record(stream)
        torch.cuda.synchronize()
        us = a.elapsed_time(b) * 1e3 / (reps * nset)
        lib.cerb_debug_set_backward_kernel(-1)
        byts = level_bytes_bwd(C, H, W, True, B)
        fl_ = 2 * level_flops(C, H, W, B)
        return {"C": C, "H": H, "W": W, "batch": B, "us_per_call": round(us, 2), "kernels_per_call": int(kern),
                "algorithmic_MB": round(byts / 1e6, 3), "GBps": round(byts / us / 1e3, 1),
                "frac": round(byts / us / 1e3 / hbm_peak, 4), "fp32_TFLOPs": round(fl_ / us / 1e6, 2),
                "fma_frac": round(fl_ / us / 1e6 / fma_peak, 4)}

    Cd, Hd, Wd = dom["C"], dom["H"], dom["W"]
    bwd_dom = time_backward(Cd, Hd, Wd, DIRS, DIRS // 2)
    bwd_train = time_backward(*HRNET_TRAIN_LEVEL[:3], HRNET_TRAIN_LEVEL[3], 0)
    # the tensor-core backward (costvolume_bwd_tc.cu, opt-in: cerb_debug_set_backward_kernel(1)) on the same shapes
    bwd_tc = {"dominant_level": time_backward(Cd, Hd, Wd, DIRS, DIRS // 2, kernel=1),
              "hrnet_train_level_b8": time_backward(*HRNET_TRAIN_LEVEL[:3], HRNET_TRAIN_LEVEL[3], 0, kernel=1),
              "note": "tcgen05 banded GEMMs (3xTF32) + shared-memory-window splat kernel; parity-tested, not the default"}

    # stand-alone flow_warp backward (cerb_flow_warp_backward: shared-memory-window splat, costvolume_splat.cu)
    def time_flow_warp_backward(C, H, W, B):
        sets_ = []
        for s_ in range(4):
            g = torch.Generator(device=dev).manual_seed(99 + s_)
            img = torch.randn(B, C, H, W, device=dev, generator=g)
            fl = (torch.randn(B, 2, H, W, device=dev, generator=g) * 1.5).clamp_(-6, 6)
            sets_.append((img, fl, torch.randn(B, C, H, W, device=dev, generator=g)))
        ctr = torch.zeros(4, dtype=torch.int64, device=dev)
        lib.cerb_debug_set_path_counters(ctypes.c_void_p(ctr.data_ptr()))
        with torch.cuda.stream(stream):
            for t in sets_:
                ops.flow_warp_backward(*t)
            torch.cuda.synchronize()
            lib.cerb_debug_set_path_counters(None)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5
            a.record(stream)
            for _ in range(reps):
                for t in sets_:
                    ops.flow_warp_backward(*t)
            b.record(stream)
        torch.cuda.synchronize()
        us = a.elapsed_time(b) * 1e3 / (reps * len(sets_))
        byts = B * C * H * W * 4 * 3 + B * H * W * 16
        c_ = ctr.cpu().tolist()
        return {"C": C, "H": H, "W": W, "batch": B, "us_per_call": round(us, 2), "launches": "memset + window-splat kernel",
                "algorithmic_MB": round(byts / 1e6, 3), "GBps": round(byts / us / 1e3, 1), "hbm_frac": round(byts / us / 1e3 / hbm_peak, 4),
                "window_tiles": int(c_[1]), "scattered_atomics_tiles": int(c_[2]),
                "scattered_atomics_kernel_us_r02_microbench": 207.9}
    fwb = time_flow_warp_backward(*HRNET_TRAIN_LEVEL[:3], HRNET_TRAIN_LEVEL[3])
    roofline_bwd = {"bound": "hbm", "achieved": bwd_dom["GBps"], "peak": hbm_peak, "unit": "GB/s", "frac": bwd_dom["frac"],
                    "traffic": None, "algorithmic_bytes_formula": "B*H*W*4*(2*81 + 4C + 4) (SURVEY 8d)",
                    "dominant_level": bwd_dom, "hrnet_train_level_b8": bwd_train, "tensor_core_path": bwd_tc,
                    "flow_warp_backward_hrnet_b8": fwb,
                    "method": "CUDA events around back-to-back cerb_warp_corr_backward calls over rotating buffer sets (stream launches)"}

    # ---- end to end with HOST buffers (pinned): every step copies its inputs host->device and its
    # results device->host inside the timed region.
    from cerberusnet_b200.host_pipeline import HostPipeline
    pipe = HostPipeline(PWC_LEVELS, batch=DIRS, depth=2, device=dev, pad_size=PAD, max_displacement=MD,
                        warp_mode=cb.WARP_TORCH, leaky_slope=SLOPE, x2_roll=DIRS // 2, shared_features=True)
    pipe.enable_arenas()   # one pinned arena per slot and direction: one memcpy each way per step
    host_sets = []
    for s in range(2):
        lv = [synth_level(i, C, H, W, wp, 5000 + 100 * s + 7 * rank, "cpu") for i, (C, H, W, wp) in enumerate(PWC_LEVELS)]
        host_sets.append(lv)
        for (hf, _, hfl), (f, fl) in zip(pipe.host_inputs(s), lv):
            hf.copy_(f)
            if hfl is not None:
                hfl.copy_(fl)
    h2d, d2h = pipe.h2d_bytes, pipe.d2h_bytes

    def time_host_path(step_fn, sync_fn, k):
        for s in range(3):
            step_fn(s)
        sync_fn()
        barrier()
        t0 = time.perf_counter()
        for s in range(k):
            step_fn(s)
        sync_fn()                       # results are in host memory
        return max_ranks((time.perf_counter() - t0) * 1e3 / k)

    k_e2e = max(3, min(args.steps, 400))
    ms_pipe = time_host_path(lambda s: pipe.submit_packed(s % 2), pipe.synchronize, k_e2e)
    # the pipeline's result must be the kernel's result
    f_l, fl_l = host_sets[(k_e2e - 1) % 2][4]
    ref_out = ops.warp_corr_forward(f_l.to(dev), f_l.to(dev), fl_l.to(dev), PAD, 1, MD, 1, 1, 1, cb.WARP_TORCH, SLOPE,
                                    x2_roll=DIRS // 2)
    if not torch.equal(ref_out.cpu(), pipe.host_outputs((k_e2e - 1) % 2)[4]):
        raise RuntimeError("HostPipeline output differs from the device-resident call")
    e2e = {"value": world * DIRS * IMG_W * IMG_H / (ms_pipe * 1e-3) / 1e6, "unit": "Mpix/s",
           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_pipe, "steps": k_e2e,
           "h2d_GBps_per_gpu": h2d / (ms_pipe * 1e-3) / 1e9, "d2h_GBps_per_gpu": d2h / (ms_pipe * 1e-3) / 1e9,
           "numa": numa,
           "api": "cerberusnet_b200.HostPipeline.submit_packed (pinned host arenas; one H2D copy, the fused kernels and "
                  "one D2H copy per step on three streams, double-buffered; each feature map crosses PCIe once and "
                  "serves both flow directions); timed host-side until the results are in host memory"}
    del pipe

    # ---- CPU baseline beside it (rank 0, N=1 only)
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        mpix, ms, cores, sample, n = run_cpu(10_000, 1, budget_s=15.0)
        cpu_baseline = {"value": mpix, "unit": "Mpix/s", "cores": cores, "kind": "port",
                        "sample": sample + ", pure-PyTorch restatement (oracle/torch_oracle.py), fp32",
                        "ms_per_step": ms}

    other = None
    if rank == 0 and world == 1 and not args.no_other_configs:
        other = run_other_configs(dev, lib, fma_peak, hbm_peak)

    # ---- the multi-GPU split north_star names: DDP training step
    train = None
    if not args.no_train:
        del sets, full_graph, rem_graph
        torch.cuda.empty_cache()
        train = run_train_block(dev, rank, world, local_rank)

    if rank == 0:
        cfg = make_config()
        line = {
            "metric": "corr+warp Mpix/s", "value": value, "unit": "Mpix/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "launches_per_step": int(launches_per_step),
            "roofline": roofline, "roofline_bwd": roofline_bwd, "levels": level_stats, "levels_smooth": level_stats_smooth,
            "levels_one_direction": level_stats_one, "cpu_baseline": cpu_baseline, "e2e": e2e, "train": train,
            "other_configs": other,
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
    ap.add_argument("--no-train", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true")
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
