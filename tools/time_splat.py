#!/usr/bin/env python
"""Device time of the stand-alone flow_warp backward (cerb_flow_warp_backward): python tools/time_splat.py [--tag T] [--smooth]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.nn.functional as F
from cerberusnet_b200 import ops
ap = argparse.ArgumentParser()
ap.add_argument("--tag", default="")
a = ap.parse_args()
dev = torch.device("cuda:0")
t_end = time.perf_counter() + 1.0
x = torch.randn(4096, 4096, device=dev)
while time.perf_counter() < t_end:
    (x @ x).sum().item()
for (B, C, H, W) in [(8, 48, 128, 256), (2, 32, 128, 256), (8, 96, 64, 128)]:
    for kind in ("iid", "smooth"):
        g = torch.Generator(device=dev).manual_seed(1)
        sets = []
        for s in range(4):
            img = torch.randn(B, C, H, W, device=dev, generator=g)
            if kind == "iid":
                fl = (torch.randn(B, 2, H, W, device=dev, generator=g) * 1.5).clamp_(-6, 6)
            else:
                co = (torch.randn(B, 2, H // 2, W // 2, device=dev, generator=g) * 1.5).clamp_(-6, 6)
                fl = F.interpolate(co * 2, scale_factor=2, mode="bilinear", align_corners=True)
            go = torch.randn(B, C, H, W, device=dev, generator=g)
            sets.append((img, fl, go))
        for t in sets:
            ops.flow_warp_backward(*t)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        e0.record()
        for _ in range(reps):
            for t in sets:
                ops.flow_warp_backward(*t)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (reps * len(sets))
        byts = B * C * H * W * 4 * 3 + B * H * W * 16
        print(f"TIMESPLAT {a.tag:10s} B={B} C={C} {H}x{W} {kind:6s}: {us:8.1f} us (memset + kernel)  {byts/us/1e3:7.1f} GB/s algorithmic")
