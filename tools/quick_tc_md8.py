#!/usr/bin/env python
"""Tensor-core forward at max_displacement 8 (two accumulator passes per tile): parity vs the generic kernel, then device
time next to the CUDA-core displacement-window kernel on KITTI-shaped levels (BASELINE configs[4])."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.nn.functional as F
import cerberusnet_b200 as cb
from cerberusnet_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(5)


def rel(x, r):
    return float((x.double() - r.double()).abs().max() / r.double().abs().max().clamp_min(1e-30))


worst = 0.0
for (B, C, H, W, pad) in [(1, 16, 24, 64, 8), (2, 32, 48, 160, 8), (1, 20, 19, 37, 8), (1, 8, 32, 48, 6), (2, 40, 16, 32, 8)]:
    x1 = F.leaky_relu(torch.randn(B, C, H, W, device=dev, generator=g), 0.1)
    x2 = F.leaky_relu(torch.randn(B, C, H, W, device=dev, generator=g), 0.1)
    for sigma in (None, 1.5, 10.0):
        fl = None if sigma is None else torch.randn(B, 2, H, W, device=dev, generator=g) * sigma
        ref = ops.warp_corr_forward(x1, x2, fl, pad, 1, 8, 1, 1, 1, cb.WARP_TORCH, 0.1, variant=5)
        out = ops.warp_corr_forward(x1, x2, fl, pad, 1, 8, 1, 1, 1, cb.WARP_TORCH, 0.1, variant=7)
        torch.cuda.synchronize()
        e = rel(out, ref)
        worst = max(worst, e / 1e-5)
        print(f"parity md=8 B={B} C={C} {H}x{W} pad={pad} sigma={sigma} rel={e:.2e}{'' if e <= 1e-5 else '   <-- FAIL'}", flush=True)
if H % 2 == 0:
    a = torch.randn(2, 16, 32, 64, device=dev); coarse = torch.randn(2, 2, 16, 32, device=dev)
    ref_flow = F.interpolate(coarse * 2, scale_factor=2, mode="bilinear", align_corners=True)
    ref = ops.warp_corr_forward(a, a, ref_flow, 8, 1, 8, 1, 1, 1, cb.WARP_TORCH, 0.1, variant=5)
    os.environ["X"] = "1"
    out, up = ops.warp_corr_forward_upflow(a, a, coarse, 8, 1, 8, 1, 1, 1, cb.WARP_TORCH, 0.1)
    torch.cuda.synchronize()
    print("upflow md=8:", rel(out, ref), bool(torch.equal(up, ref_flow)))
    worst = max(worst, rel(out, ref) / 1e-5)
print("PARITY", "OK" if worst <= 1.0 else "FAIL", flush=True)
if "--notime" in sys.argv:
    sys.exit(0 if worst <= 1.0 else 1)
t_end = time.perf_counter() + 1.0
x = torch.randn(4096, 4096, device=dev)
while time.perf_counter() < t_end:
    (x @ x).sum().item()
for (C, H, W, B) in [(32, 96, 320, 32), (48, 96, 312, 32), (64, 48, 160, 32), (32, 96, 320, 4)]:
    f1 = F.leaky_relu(torch.randn(B, C, H, W, device=dev, generator=g), 0.1)
    f2 = F.leaky_relu(torch.randn(B, C, H, W, device=dev, generator=g), 0.1)
    fl = (torch.randn(B, 2, H, W, device=dev, generator=g) * 1.5).clamp_(-6, 6)
    out = torch.empty(B, 289, H, W, device=dev)
    for variant in (7, 1):
        for _ in range(3):
            ops.warp_corr_forward(f1, f2, fl, 8, 1, 8, 1, 1, 1, cb.WARP_TORCH, 0.1, out=out, variant=variant)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.warp_corr_forward(f1, f2, fl, 8, 1, 8, 1, 1, 1, cb.WARP_TORCH, 0.1, out=out, variant=variant)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 100
        fl_ = 2 * B * H * W * C * 289
        print(f"TIME md=8 variant={variant} C={C} {H}x{W} B={B}: {us:9.1f} us  {fl_/us/1e6:6.2f} TFLOP/s (useful fp32)", flush=True)
