#!/usr/bin/env python
"""Device time of forward + backward of the fused level op: python tools/time_backward.py [--batch B] [--pyramid pwc|hrnet]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.nn.functional as F
import cerberusnet_b200 as cb
from cerberusnet_b200 import ops
PYR = {"pwc": [(192, 8, 16, False), (128, 16, 32, True), (96, 32, 64, True), (64, 64, 128, True), (32, 128, 256, True)],
       "hrnet": [(384, 16, 32, False), (192, 32, 64, True), (96, 64, 128, True), (48, 128, 256, True)]}
ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=8); ap.add_argument("--pyramid", default="hrnet")
a = ap.parse_args(); B = a.batch; dev = torch.device("cuda:0")
x = torch.randn(4096, 4096, device=dev); t_end = time.perf_counter() + 1.0
while time.perf_counter() < t_end: (x @ x).sum().item()
tf = tb = 0.0
for li, (C, H, W, wp) in enumerate(PYR[a.pyramid]):
    x1 = F.leaky_relu(torch.randn(B, C, H, W, device=dev), 0.1); x2 = F.leaky_relu(torch.randn(B, C, H, W, device=dev), 0.1)
    fl = (torch.randn(B, 2, H, W, device=dev) * 1.5).clamp_(-6, 6) if wp else None
    out = ops.warp_corr_forward(x1, x2, fl, 4, 1, 4, 1, 1, 1, 0, 0.1); g = torch.randn_like(out)
    def timeit(fn, reps=10):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps): fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) * 1e3 / reps
    f_us = timeit(lambda: ops.warp_corr_forward(x1, x2, fl, 4, 1, 4, 1, 1, 1, 0, 0.1, out=out))
    b_us = timeit(lambda: ops.warp_corr_backward(x1, x2, fl, out, g, 4, 1, 4, 1, 1, 1, 0, 0.1))
    bb = B * H * W * 4 * (2 * 81 + 4 * C + 4); bf = 4 * B * H * W * C * 81
    tf += f_us; tb += b_us
    print(f"L{li} B={B} C={C} {H}x{W}: fwd {f_us:8.1f} us   bwd {b_us:8.1f} us  ({bb/b_us/1e3:6.0f} GB/s, {bf/b_us/1e6:5.1f} TFLOP/s)")
print(f"total fwd {tf:.0f} us, bwd {tb:.0f} us")
