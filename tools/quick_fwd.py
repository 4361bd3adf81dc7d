#!/usr/bin/env python
"""Fast developer loop for the fused forward on a GPU box: parity of the auto path against the
generic one-thread-per-output kernel (itself pinned to the oracle by tests/) on a few shapes, then
device time of the finest PWC levels at several batch sizes / flow models.

    [CERB_LIB_OVERRIDE=path/to/lib_variant.so] python tools/quick_fwd.py [--notime] [--tag name]
"""
import argparse, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.nn.functional as F
import cerberusnet_b200 as cb
from cerberusnet_b200 import ops

ap = argparse.ArgumentParser()
ap.add_argument("--notime", action="store_true")
ap.add_argument("--tag", default="")
ap.add_argument("--cases", default="L4:1,L4:2,L4:8,L3:1,L3:2")
ap.add_argument("--variant", type=int, default=0)
ap.add_argument("--flows", default="iid,smooth")
a = ap.parse_args()
dev = torch.device("cuda:0")
HBM = 6461.5


def mkflow(kind, B, H, W, g):
    if kind == "iid":
        return (torch.randn(B, 2, H, W, device=dev, generator=g) * 1.5).clamp_(-6, 6)
    if kind == "smooth":
        coarse = (torch.randn(B, 2, H // 2, W // 2, device=dev, generator=g) * 1.5).clamp_(-6, 6)
        return F.interpolate(coarse * 2, scale_factor=2, mode="bilinear", align_corners=True)
    if kind == "zero":
        return torch.zeros(B, 2, H, W, device=dev)
    return None


def rel(x, r):
    return float((x.double() - r.double()).abs().max() / r.double().abs().max().clamp_min(1e-30))


worst = 0.0
g = torch.Generator(device=dev).manual_seed(7)
for (B, C, H, W, md) in [(1, 32, 128, 256, 4), (2, 64, 64, 128, 4), (3, 20, 24, 72, 4), (2, 48, 40, 100, 4), (1, 16, 24, 64, 8),
                         (1, 192, 8, 16, 4), (2, 128, 16, 32, 4)]:
    x1 = F.leaky_relu(torch.randn(B, C, H, W, device=dev, generator=g), 0.1)
    x2 = F.leaky_relu(torch.randn(B, C, H, W, device=dev, generator=g), 0.1)
    for kind in ("none", "iid", "smooth"):
        if kind == "smooth" and (H % 2 or W % 2):
            continue
        fl = mkflow(kind, B, H, W, g)
        for dt in (torch.float32, torch.float16):
            if dt != torch.float32 and (C > 64 or md != 4):
                continue
            a1, a2 = x1.to(dt), x2.to(dt)
            ref = ops.warp_corr_forward(a1, a2, fl, md, 1, md, 1, 1, 1, cb.WARP_TORCH, 0.1, variant=5)
            out = ops.warp_corr_forward(a1, a2, fl, md, 1, md, 1, 1, 1, cb.WARP_TORCH, 0.1)
            # strided output (concat-buffer slice)
            buf = torch.zeros(B, ref.shape[1] + 8, H, W, device=dev, dtype=dt)
            ops.warp_corr_forward(a1, a2, fl, md, 1, md, 1, 1, 1, cb.WARP_TORCH, 0.1, out=buf[:, 4:4 + ref.shape[1]])
            torch.cuda.synchronize()
            e = max(rel(out, ref), rel(buf[:, 4:4 + ref.shape[1]], ref))
            tol = 1e-5 if dt == torch.float32 else 2e-3
            worst = max(worst, e / tol)
            print(f"parity B={B} C={C} {H}x{W} md={md} flow={kind:6s} {str(dt)[6:]:8s} rel={e:.2e}{'' if e <= tol else '   <-- FAIL'}")
# x2_roll: both flow directions in one launch == two launches with the roles swapped
for (B, C, H, W) in [(2, 32, 128, 256), (4, 64, 32, 64), (2, 128, 16, 32), (6, 24, 40, 72)]:
    f = F.leaky_relu(torch.randn(B, C, H, W, device=dev, generator=g), 0.1)
    fl = mkflow("iid", B, H, W, g)
    for flow in (fl, None):
        both = ops.warp_corr_forward(f, f, flow, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1, x2_roll=B // 2)
        ref = ops.warp_corr_forward(f, torch.roll(f, -(B // 2), 0).contiguous(), flow, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1)
        gen = ops.warp_corr_forward(f, f, flow, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1, x2_roll=B // 2, variant=5)
        torch.cuda.synchronize()
        e = max(rel(both, ref), rel(gen, ref))
        worst = max(worst, e / 1e-5)
        print(f"roll   B={B} C={C} {H}x{W} flow={'iid' if flow is not None else 'none'} rel={e:.2e}{'' if e <= 1e-5 else '   <-- FAIL'}")
print("PARITY", "OK" if worst <= 1.0 else "FAIL", a.tag)
if a.notime:
    sys.exit(0 if worst <= 1.0 else 1)

t_end = time.perf_counter() + 1.0
x = torch.randn(4096, 4096, device=dev)
while time.perf_counter() < t_end:
    (x @ x).sum().item()
SH = {"L4": (32, 128, 256), "L3": (64, 64, 128), "L2": (96, 32, 64), "L1": (128, 16, 32), "H3": (48, 128, 256), "H2": (96, 64, 128)}
for case in a.cases.split(","):
    name, B = case.split(":")
    B = int(B)
    C, H, W = SH[name]
    for kind in a.flows.split(","):
        byts = B * H * W * 4 * (2 * C + 81 + (2 if kind != "none" else 0))
        nsets = max(2, int(300e6 // byts) + 1)
        sets = []
        for s in range(nsets):
            gg = torch.Generator(device=dev).manual_seed(100 * s + 1)
            x1 = F.leaky_relu(torch.randn(B, C, H, W, device=dev, generator=gg), 0.1)
            x2 = F.leaky_relu(torch.randn(B, C, H, W, device=dev, generator=gg), 0.1)
            sets.append((x1, x2, mkflow(kind, B, H, W, gg), torch.empty(B, 81, H, W, device=dev)))
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
            reps = max(4, int(4000 / nsets))
            e0.record(st)
            for _ in range(reps):
                gr.replay()
            e1.record(st)
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / (reps * nsets)
        print(f"TIME {a.tag:12s} {name} B={B} flow={kind:6s}: {us:8.2f} us  {byts/us/1e3:7.1f} GB/s  frac={byts/us/1e3/HBM:.3f}")
        del sets
