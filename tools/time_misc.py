#!/usr/bin/env python
"""Device time of the less common configurations: 16-bit dtypes and max_displacement=8 (KITTI-shaped)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, torch.nn.functional as F
import cerberusnet_b200 as cb
from cerberusnet_b200 import ops
dev = torch.device("cuda:0")
x = torch.randn(4096, 4096, device=dev); t_end = time.perf_counter() + 1.0
while time.perf_counter() < t_end: (x @ x).sum().item()
def timeit(fn, reps=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps
for B in (1, 8):
    for (C, H, W) in ((32, 128, 256), (64, 64, 128)):
        for dt in (torch.float32, torch.bfloat16, torch.float16):
            x1 = F.leaky_relu(torch.randn(B, C, H, W, device=dev), 0.1).to(dt); x2 = F.leaky_relu(torch.randn(B, C, H, W, device=dev), 0.1).to(dt)
            fl = (torch.randn(B, 2, H, W, device=dev) * 1.5).clamp_(-6, 6)
            out = torch.empty(B, 81, H, W, device=dev, dtype=dt)
            t = timeit(lambda: ops.warp_corr_forward(x1, x2, fl, 4, 1, 4, 1, 1, 1, 0, 0.1, out=out))
            print(f"fwd B={B} C={C} {H}x{W} {str(dt):16s} {t:8.1f} us")
# BASELINE configs[4] / SURVEY 8d config 5: KITTI-shaped pairs padded to 1280x384 (PWC pyramid) or 1248x384 (HRNet),
# finest level, max_displacement = pad = 8 (289 planes), batch sweep
for (C, H, W) in ((32, 96, 320), (48, 96, 312)):
    for B in (1, 4, 32):
        x1 = F.leaky_relu(torch.randn(B, C, H, W, device=dev), 0.1); x2 = F.leaky_relu(torch.randn(B, C, H, W, device=dev), 0.1)
        fl = (torch.randn(B, 2, H, W, device=dev) * 1.5).clamp_(-6, 6)
        out = torch.empty(B, 289, H, W, device=dev)
        t8 = timeit(lambda: ops.warp_corr_forward(x1, x2, fl, 8, 1, 8, 1, 1, 1, 0, 0.1, out=out), reps=5)
        tg = timeit(lambda: ops.warp_corr_forward(x1, x2, fl, 8, 1, 8, 1, 1, 1, 0, 0.1, out=out, variant=5), reps=2)
        fl8 = 2 * B * H * W * C * 289
        by8 = B * H * W * 4 * (2 * C + 289 + 2)
        print(f"md=8 B={B} C={C} {H}x{W}: windowed fast path {t8:9.1f} us ({fl8 / t8 / 1e6:5.2f} TFLOP/s, {by8 / t8 / 1e3:6.0f} GB/s)   generic kernel {tg:9.1f} us")
# backward at md = 8 (tiled kernel, 4 displacement windows)
for (B, C, H, W) in ((4, 32, 96, 320),):
    x1 = F.leaky_relu(torch.randn(B, C, H, W, device=dev), 0.1); x2 = F.leaky_relu(torch.randn(B, C, H, W, device=dev), 0.1)
    fl = (torch.randn(B, 2, H, W, device=dev) * 1.5).clamp_(-6, 6)
    out = ops.warp_corr_forward(x1, x2, fl, 8, 1, 8, 1, 1, 1, 0, 0.1); gg = torch.randn_like(out)
    tb = timeit(lambda: ops.warp_corr_backward(x1, x2, fl, out, gg, 8, 1, 8, 1, 1, 1, 0, 0.1), reps=5)
    print(f"md=8 backward B={B} C={C} {H}x{W}: {tb:9.1f} us ({4 * B * H * W * C * 289 / tb / 1e6:5.2f} TFLOP/s)")
