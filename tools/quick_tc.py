#!/usr/bin/env python
"""Developer loop for the tensor-core forward (variant 7) on a GPU box: parity against the generic
one-thread-per-output kernel (pinned to the oracle by tests/) and the C oracle, then device time next to the
CUDA-core path on the bench's shapes.

    python tools/quick_tc.py [--notime] [--cases L4:2,L4:1,...]
"""
import argparse, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.nn.functional as F
import cerberusnet_b200 as cb
from cerberusnet_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--notime", action="store_true")
ap.add_argument("--noparity", action="store_true")
ap.add_argument("--cases", default="L4:2,L4:1,L3:2,L2:2,L4:8,H3:8")
ap.add_argument("--flows", default="iid,smooth")
ap.add_argument("--variants", default="7,0")
a = ap.parse_args()
dev = torch.device("cuda:0")
HBM = 6461.5
TC = 7


def mkflow(kind, B, H, W, g):
    if kind == "iid":
        return (torch.randn(B, 2, H, W, device=dev, generator=g) * 1.5).clamp_(-6, 6)
    if kind == "smooth":
        coarse = (torch.randn(B, 2, H // 2, W // 2, device=dev, generator=g) * 1.5).clamp_(-6, 6)
        return F.interpolate(coarse * 2, scale_factor=2, mode="bilinear", align_corners=True)
    if kind == "big":
        return torch.randn(B, 2, H, W, device=dev, generator=g) * 12.0
    if kind == "zero":
        return torch.zeros(B, 2, H, W, device=dev)
    return None


def rel(x, r):
    return float((x.double() - r.double()).abs().max() / r.double().abs().max().clamp_min(1e-30))


worst = 0.0
g = torch.Generator(device=dev).manual_seed(7)
for (B, C, H, W, pad) in [] if a.noparity else [(1, 32, 16, 32, 4), (1, 32, 128, 256, 4), (2, 64, 64, 128, 4), (3, 20, 24, 72, 4), (2, 48, 40, 100, 4),
                          (1, 192, 8, 16, 4), (2, 128, 16, 32, 4), (2, 17, 19, 37, 4), (1, 8, 24, 48, 6), (1, 40, 33, 50, 2)]:
    x1 = F.leaky_relu(torch.randn(B, C, H, W, device=dev, generator=g), 0.1)
    x2 = F.leaky_relu(torch.randn(B, C, H, W, device=dev, generator=g), 0.1)
    for kind in ("none", "iid", "smooth", "big"):
        if kind == "smooth" and (H % 2 or W % 2):
            continue
        fl = mkflow(kind, B, H, W, g)
        for slope in (0.1, None):
            ref = ops.warp_corr_forward(x1, x2, fl, pad, 1, 4, 1, 1, 1, cb.WARP_TORCH, slope, variant=5)
            out = ops.warp_corr_forward(x1, x2, fl, pad, 1, 4, 1, 1, 1, cb.WARP_TORCH, slope, variant=TC)
            buf = torch.zeros(B, ref.shape[1] + 8, ref.shape[2], ref.shape[3], device=dev)
            ops.warp_corr_forward(x1, x2, fl, pad, 1, 4, 1, 1, 1, cb.WARP_TORCH, slope, out=buf[:, 4:4 + ref.shape[1]], variant=TC)
            torch.cuda.synchronize()
            e = max(rel(out, ref), rel(buf[:, 4:4 + ref.shape[1]], ref))
            worst = max(worst, e / 1e-5)
            print(f"parity B={B} C={C} {H}x{W} pad={pad} flow={kind:6s} slope={slope} rel={e:.2e}{'' if e <= 1e-5 else '   <-- FAIL'}", flush=True)
for (B, C, H, W) in [] if a.noparity else [(2, 32, 128, 256), (4, 64, 32, 64), (6, 24, 40, 72)]:
    f = F.leaky_relu(torch.randn(B, C, H, W, device=dev, generator=g), 0.1)
    fl = mkflow("iid", B, H, W, g)
    both = ops.warp_corr_forward(f, f, fl, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1, x2_roll=B // 2, variant=TC)
    ref = ops.warp_corr_forward(f, f, fl, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1, x2_roll=B // 2, variant=5)
    torch.cuda.synchronize()
    e = rel(both, ref)
    worst = max(worst, e / 1e-5)
    print(f"roll   B={B} C={C} {H}x{W} rel={e:.2e}{'' if e <= 1e-5 else '   <-- FAIL'}", flush=True)
print("PARITY", "OK" if worst <= 1.0 else "FAIL", flush=True)
if a.notime:
    sys.exit(0 if worst <= 1.0 else 1)

t_end = time.perf_counter() + 1.0
x = torch.randn(4096, 4096, device=dev)
while time.perf_counter() < t_end:
    (x @ x).sum().item()
SH = {"L4": (32, 128, 256), "L3": (64, 64, 128), "L2": (96, 32, 64), "L1": (128, 16, 32), "H3": (48, 128, 256), "H2": (96, 64, 128), "H1": (192, 32, 64), "H0": (384, 16, 32)}
for case in a.cases.split(","):
    name, B = case.split(":")
    B = int(B)
    C, H, W = SH[name]
    for kind in a.flows.split(","):
        byts = B * H * W * 4 * (2 * C + 81 + 2)
        nsets = max(2, int(300e6 // byts) + 1)
        sets = []
        for s in range(nsets):
            gg = torch.Generator(device=dev).manual_seed(100 * s + 1)
            f = F.leaky_relu(torch.randn(B, C, H, W, device=dev, generator=gg), 0.1)
            sets.append((f, mkflow(kind, B, H, W, gg), torch.empty(B, 81, H, W, device=dev)))
        for variant in [int(v) for v in a.variants.split(",")]:
            st = torch.cuda.Stream()
            gr = torch.cuda.CUDAGraph()

            def run():
                for (f, fl, out) in sets:   # both flow directions of B/2 pairs (x2 = x1 rolled by B/2), as the bench does
                    ops.warp_corr_forward(f, f, fl, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1, out=out, variant=variant, x2_roll=B // 2)
            with torch.cuda.stream(st):
                run(); torch.cuda.synchronize()
                with torch.cuda.graph(gr, stream=st):
                    run()
                gr.replay(); torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                reps = max(4, int(4000 / nsets))
                e0.record(st)
                for _ in range(reps):
                    gr.replay()
                e1.record(st)
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / (reps * nsets)
            print(f"TIME variant={variant} {name} B={B} flow={kind:6s}: {us:8.2f} us  {byts/us/1e3:7.1f} GB/s  frac={byts/us/1e3/HBM:.3f}", flush=True)
        del sets
