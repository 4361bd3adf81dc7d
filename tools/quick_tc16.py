#!/usr/bin/env python
"""16-bit tensor-core forward (kind::f16): error against the float64 C oracle on the same 16-bit inputs next to the CUDA-core
16-bit kernel's, then device time.  Bars (tests): fp16 6e-4, bf16 5e-3 of max|ref|."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import torch.nn.functional as F
import cerberusnet_b200 as cb
from cerberusnet_b200 import ops
from oracle import c_oracle as co
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(5)


def rel(x, r):
    return float(np.abs(x.astype(np.float64) - r).max() / max(np.abs(r).max(), 1e-30))


worst = 0.0
for dt, tol in ((torch.float16, 6e-4), (torch.bfloat16, 5e-3)):
    for (B, C, H, W, pad, md) in [(2, 32, 24, 64, 4, 4), (1, 64, 40, 96, 4, 4), (2, 20, 19, 37, 4, 4), (1, 48, 32, 64, 8, 8), (1, 24, 30, 52, 2, 4), (1, 96, 16, 48, 4, 4)]:
        x1 = F.leaky_relu(torch.randn(B, C, H, W, device=dev, generator=g), 0.1).to(dt)
        x2 = F.leaky_relu(torch.randn(B, C, H, W, device=dev, generator=g), 0.1).to(dt)
        for sigma in (None, 1.5, 12.0):
            fl = None if sigma is None else torch.randn(B, 2, H, W, device=dev, generator=g) * sigma
            ref = co.level_forward(x1.float().cpu().numpy(), x2.float().cpu().numpy(), None if fl is None else fl.cpu().numpy(),
                                   pad, 1, md, 1, 1, co.WARP_TORCH, 0.1)
            o7 = ops.warp_corr_forward(x1, x2, fl, pad, 1, md, 1, 1, 1, cb.WARP_TORCH, 0.1, variant=7)
            o1 = ops.warp_corr_forward(x1, x2, fl, pad, 1, md, 1, 1, 1, cb.WARP_TORCH, 0.1, variant=1)
            torch.cuda.synchronize()
            e7, e1 = rel(o7.float().cpu().numpy(), ref), rel(o1.float().cpu().numpy(), ref)
            worst = max(worst, e7 / tol)
            print(f"{str(dt)[6:]:9s} B={B} C={C} {H}x{W} pad={pad} md={md} sigma={sigma}: tc {e7:.2e}  cuda-core {e1:.2e}{'' if e7 <= tol else '   <-- FAIL'}", flush=True)
print("PARITY", "OK" if worst <= 1.0 else "FAIL", flush=True)
if "--notime" in sys.argv:
    sys.exit(0 if worst <= 1.0 else 1)
t_end = time.perf_counter() + 1.0
x = torch.randn(4096, 4096, device=dev)
while time.perf_counter() < t_end:
    (x @ x).sum().item()
for dt in (torch.float16, torch.bfloat16, torch.float32):
    for (C, H, W, B, md) in [(32, 128, 256, 8, 4), (48, 128, 256, 8, 4), (96, 64, 128, 8, 4), (32, 96, 320, 32, 8), (48, 128, 256, 1, 4)]:
        f1 = F.leaky_relu(torch.randn(B, C, H, W, device=dev, generator=g), 0.1).to(dt)
        f2 = F.leaky_relu(torch.randn(B, C, H, W, device=dev, generator=g), 0.1).to(dt)
        fl = (torch.randn(B, 2, H, W, device=dev, generator=g) * 1.5).clamp_(-6, 6)
        out = torch.empty(B, (2 * md + 1) ** 2, H, W, device=dev, dtype=dt)
        res = []
        for variant in (7, 1):
            for _ in range(3):
                ops.warp_corr_forward(f1, f2, fl, md, 1, md, 1, 1, 1, cb.WARP_TORCH, 0.1, out=out, variant=variant)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                ops.warp_corr_forward(f1, f2, fl, md, 1, md, 1, 1, 1, cb.WARP_TORCH, 0.1, out=out, variant=variant)
            e1.record()
            torch.cuda.synchronize()
            res.append(e0.elapsed_time(e1) * 50)
        print(f"TIME {str(dt)[6:]:9s} C={C} {H}x{W} B={B} md={md}: tc {res[0]:8.1f} us   cuda-core {res[1]:8.1f} us", flush=True)
