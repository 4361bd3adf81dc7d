#!/usr/bin/env python
"""Per-level device time of the fused forward for other batch sizes / flow statistics / pyramids.
    python tools/time_levels.py [--batch B] [--flow iid|smooth|zero] [--pyramid pwc|hrnet]"""
import argparse, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.nn.functional as F
import cerberusnet_b200 as cb
from cerberusnet_b200 import ops

PYR = {"pwc": [(192, 8, 16, False), (128, 16, 32, True), (96, 32, 64, True), (64, 64, 128, True), (32, 128, 256, True)],
       "hrnet": [(384, 16, 32, False), (192, 32, 64, True), (96, 64, 128, True), (48, 128, 256, True)]}
ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1)
ap.add_argument("--flow", default="iid")
ap.add_argument("--pyramid", default="pwc")
ap.add_argument("--sets", type=int, default=0)
ap.add_argument("--variant", type=int, default=0)
a = ap.parse_args()
dev = torch.device("cuda:0")
B = a.batch
hbm = 6461.5
t_end = time.perf_counter() + 1.0
x = torch.randn(4096, 4096, device=dev)
while time.perf_counter() < t_end:
    (x @ x).sum().item()
tot = 0.0
for li, (C, H, W, wp) in enumerate(PYR[a.pyramid]):
    byts = B * H * W * 4 * (2 * C + 81 + (2 if wp else 0))
    nsets = a.sets or max(2, int(300e6 // byts) + 1)
    sets = []
    for s in range(nsets):
        g = torch.Generator(device=dev).manual_seed(100 * s + li)
        x1 = F.leaky_relu(torch.randn(B, C, H, W, device=dev, generator=g), 0.1)
        x2 = F.leaky_relu(torch.randn(B, C, H, W, device=dev, generator=g), 0.1)
        fl = None
        if wp:
            if a.flow == "iid":
                fl = (torch.randn(B, 2, H, W, device=dev, generator=g) * 1.5).clamp_(-6, 6)
            elif a.flow == "smooth":  # what a decoder produces: coarser flow * 2, bilinearly up-sampled
                coarse = (torch.randn(B, 2, H // 2, W // 2, device=dev, generator=g) * 1.5).clamp_(-6, 6)
                fl = F.interpolate(coarse * 2, scale_factor=2, mode="bilinear", align_corners=True)
            else:
                fl = torch.zeros(B, 2, H, W, device=dev)
        sets.append((x1, x2, fl, torch.empty(B, 81, H, W, device=dev)))
    st = torch.cuda.Stream()
    gr = torch.cuda.CUDAGraph()
    def run():
        for (x1, x2, fl, out) in sets:
            ops.warp_corr_forward(x1, x2, fl, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1, out=out, variant=a.variant)
    with torch.cuda.stream(st):
        run(); torch.cuda.synchronize()
        with torch.cuda.graph(gr, stream=st):
            run()
        gr.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record(st)
        for _ in range(reps):
            gr.replay()
        e1.record(st)
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * nsets)
    fl_ = 2 * B * H * W * C * 81
    tot += us
    print(f"L{li} B={B} C={C} {H}x{W} flow={a.flow if wp else 'none'}: {us:8.2f} us  {byts/us/1e3:7.1f} GB/s ({byts/us/1e3/hbm:.3f} of HBM)  {fl_/us/1e6:6.2f} TFLOP/s ({fl_/us/1e6/70.4:.3f} of FMA)")
print(f"pyramid total {tot:.2f} us -> {B*1024*512/tot:.0f} Mpix/s")
