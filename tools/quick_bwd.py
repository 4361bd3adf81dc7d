#!/usr/bin/env python
"""Fast developer loop for the backward on a GPU box: parity against autograd of the pure-PyTorch oracle run on the
GPU (tests/ pin that oracle to the reference), x2_roll against an explicit roll, then device time of training shapes.

    [CERB_LIB_OVERRIDE=...] python tools/quick_bwd.py [--notime]
"""
import argparse, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.nn.functional as F
import cerberusnet_b200 as cb
from cerberusnet_b200 import ops
from oracle import torch_oracle as to

ap = argparse.ArgumentParser()
ap.add_argument("--notime", action="store_true")
ap.add_argument("--tag", default="")
ap.add_argument("--cases", default="H3:8,H2:8,L4:2,L4:8")
a = ap.parse_args()
dev = torch.device("cuda:0")
HBM = 6461.5


def rel(x, r):
    return float((x.double() - r.double()).abs().max() / r.double().abs().max().clamp_min(1e-30))


worst = 0.0
import numpy as np
from oracle import c_oracle as co
g = torch.Generator(device=dev).manual_seed(11)
for (B, C, H, W, md, flow_on) in [(2, 16, 24, 64, 4, True), (2, 16, 24, 64, 4, False), (1, 48, 40, 96, 4, True), (3, 20, 17, 36, 4, True),
                                  (2, 8, 24, 64, 8, True), (2, 33, 16, 32, 4, False), (4, 32, 32, 64, 4, True)]:
    x1 = torch.randn(B, C, H, W, device=dev, generator=g)
    x2 = torch.randn(B, C, H, W, device=dev, generator=g)
    fl = (torch.randn(B, 2, H, W, device=dev, generator=g) * 2.0) if flow_on else None
    go = torch.randn(B, (2 * md + 1) ** 2, H, W, device=dev, generator=g)
    # the C oracle (fp32 sample positions like the kernels, fp64 accumulation): what tests/ compare against
    r1, r2, rf = co.level_backward(x1.cpu().numpy(), x2.cpu().numpy(), fl.cpu().numpy() if flow_on else None, go.cpu().numpy(),
                                   md, 1, md, 1, 1, 0, 0.1)
    out = ops.warp_corr_forward(x1, x2, fl, md, 1, md, 1, 1, 1, cb.WARP_TORCH, 0.1)
    g1, g2, gf = ops.warp_corr_backward(x1, x2, fl, out, go, md, 1, md, 1, 1, 1, cb.WARP_TORCH, 0.1)
    e = [rel(g1.cpu(), torch.from_numpy(r1)), rel(g2.cpu(), torch.from_numpy(r2))] + ([rel(gf.cpu(), torch.from_numpy(rf))] if flow_on else [])
    # x2_roll against an explicit roll of x2 (the second gradient comes back in x2's own order)
    roll = B // 2
    if roll:
        x2r = torch.roll(x2, -roll, 0).contiguous()
        out_r = ops.warp_corr_forward(x1, x2, fl, md, 1, md, 1, 1, 1, cb.WARP_TORCH, 0.1, x2_roll=roll)
        h1, h2, hf = ops.warp_corr_backward(x1, x2, fl, out_r, go, md, 1, md, 1, 1, 1, cb.WARP_TORCH, 0.1, x2_roll=roll)
        out_e = ops.warp_corr_forward(x1, x2r, fl, md, 1, md, 1, 1, 1, cb.WARP_TORCH, 0.1)
        k1, k2, kf = ops.warp_corr_backward(x1, x2r, fl, out_e, go, md, 1, md, 1, 1, 1, cb.WARP_TORCH, 0.1)
        e += [rel(out_r, out_e), rel(h1, k1), rel(h2, torch.roll(k2, roll, 0))] + ([rel(hf, kf)] if flow_on else [])
    torch.cuda.synchronize()
    worst = max(worst, max(e) / 2e-5)
    print(f"bwd parity B={B} C={C} {H}x{W} md={md} flow={int(flow_on)}: " + " ".join(f"{v:.1e}" for v in e) +
          ("" if max(e) <= 2e-5 else "   <-- FAIL"))
print("PARITY", "OK" if worst <= 1.0 else "FAIL", a.tag)
if a.notime:
    sys.exit(0 if worst <= 1.0 else 1)
t_end = time.perf_counter() + 1.0
x = torch.randn(4096, 4096, device=dev)
while time.perf_counter() < t_end:
    (x @ x).sum().item()
SH = {"L4": (32, 128, 256), "L3": (64, 64, 128), "H3": (48, 128, 256), "H2": (96, 64, 128), "H1": (192, 32, 64)}
for case in a.cases.split(","):
    name, B = case.split(":")
    B = int(B)
    C, H, W = SH[name]
    byts = B * H * W * 4 * (2 * 81 + 4 * C + 4)
    nset = max(2, int(300e6 // byts) + 1)
    bs = []
    for s in range(nset):
        gg = torch.Generator(device=dev).manual_seed(5 + s)
        x1 = F.leaky_relu(torch.randn(B, C, H, W, device=dev, generator=gg), 0.1)
        x2 = F.leaky_relu(torch.randn(B, C, H, W, device=dev, generator=gg), 0.1)
        fl = (torch.randn(B, 2, H, W, device=dev, generator=gg) * 1.5).clamp_(-6, 6)
        out = ops.warp_corr_forward(x1, x2, fl, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1)
        bs.append((x1, x2, fl, out, torch.randn(B, 81, H, W, device=dev, generator=gg)))
    for t in bs:
        ops.warp_corr_backward(*t, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    e0.record()
    for _ in range(reps):
        for t in bs:
            ops.warp_corr_backward(*t, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (reps * nset)
    print(f"TIMEBWD {a.tag:10s} {name} B={B}: {us:9.2f} us  {byts/us/1e3:7.1f} GB/s  frac={byts/us/1e3/HBM:.3f}  {4*B*H*W*C*81/us/1e6:.2f} TFLOP/s")
    del bs
