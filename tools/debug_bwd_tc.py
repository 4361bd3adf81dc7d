#!/usr/bin/env python
"""Structured-input probe of the tensor-core backward (CERB_DEBUG_BWD_TC=1): delta gradients against index-coded feature
maps, so that a wrong shift / lane mapping / operand layout shows up as a readable pattern."""
import os, sys
os.environ.setdefault("CERB_DEBUG_BWD_TC", "1")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import cerberusnet_b200 as cb
from cerberusnet_b200 import ops

torch.set_printoptions(linewidth=250, precision=1, sci_mode=False)
dev = torch.device("cuda:0")


def run(H, W, C, d, B=1):
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    x1 = torch.stack([1000.0 * (c + 1) + 32 * yy + xx for c in range(C)]).float()[None].repeat(B, 1, 1, 1).to(dev)
    x2 = torch.stack([-1000.0 * (c + 1) - 32 * yy - xx for c in range(C)]).float()[None].repeat(B, 1, 1, 1).to(dev)
    go = torch.zeros(B, 81, H, W, device=dev)
    go[:, d] = 1.0
    out = torch.ones(B, 81, H, W, device=dev)
    g1, g2, _ = ops.warp_corr_backward(x1, x2, None, out, go, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1)
    torch.cuda.synchronize()
    dy, dx = d // 9 - 4, d % 9 - 4
    # expected: g1[c, p] = x2[c, p + d] / C (zero outside), g2[c, q] = x1[c, q - d] / C
    e1 = torch.zeros_like(x1)
    e2 = torch.zeros_like(x1)
    ys0, ys1 = max(0, -dy), min(H, H - dy)
    xs0, xs1 = max(0, -dx), min(W, W - dx)
    e1[:, :, ys0:ys1, xs0:xs1] = x2[:, :, ys0 + dy:ys1 + dy, xs0 + dx:xs1 + dx] / C
    e2[:, :, ys0 + dy:ys1 + dy, xs0 + dx:xs1 + dx] = x1[:, :, ys0:ys1, xs0:xs1] / C
    print(f"=== H={H} W={W} C={C} d=({dy},{dx})  max|g1-e1|={float((g1 - e1).abs().max()):.3g}  max|g2-e2|={float((g2 - e2).abs().max()):.3g}")
    return g1, e1, g2, e2


g1, e1, g2, e2 = run(8, 16, 16, 40)
print("g1*C channel 0 (expected x2 ch0 = -1000 - 32y - x):")
print((g1[0, 0] * 16).cpu())
print("g1*C channel 1:")
print((g1[0, 1] * 16).cpu())
print("g2*C channel 0 (expected x1 ch0 = 1000 + 32y + x):")
print((g2[0, 0] * 16).cpu())
g1, e1, g2, e2 = run(8, 16, 16, 0)
print("d=(-4,-4): g1*C channel 0, expected x2[p - 4] ; got:")
print((g1[0, 0] * 16).cpu())
print("expected:")
print((e1[0, 0] * 16).cpu())
for (H, W, C, d) in [(16, 32, 16, 40), (16, 32, 16, 44), (16, 32, 48, 76), (24, 64, 32, 3)]:
    run(H, W, C, d)

print("d=(-4,-4): g2*C channel 0, expected x1[q + 4]; got:")
g1, e1, g2, e2 = run(8, 16, 16, 0)
print((g2[0, 0] * 16).cpu())
print("expected:")
print((e2[0, 0] * 16).cpu())

# random inputs against the C oracle, one feature at a time
import numpy as np
from oracle import c_oracle as co
gen = torch.Generator(device=dev).manual_seed(3)


def rel(x, r):
    r = torch.from_numpy(r)
    return float((x.cpu().double() - r.double()).abs().max() / r.double().abs().max().clamp_min(1e-30))


for name, B, C, H, W, use_mask, use_flow in [("plain", 1, 16, 8, 16, False, False), ("plain 2 tiles", 1, 16, 8, 32, False, False),
                                              ("mask", 1, 16, 8, 16, True, False), ("B=2", 2, 16, 8, 16, False, False),
                                              ("flow", 1, 16, 8, 16, False, True), ("flow+mask B=2 24x64", 2, 16, 24, 64, True, True)]:
    x1 = torch.randn(B, C, H, W, device=dev, generator=gen)
    x2 = torch.randn(B, C, H, W, device=dev, generator=gen)
    fl = torch.randn(B, 2, H, W, device=dev, generator=gen) * 2 if use_flow else None
    go = torch.randn(B, 81, H, W, device=dev, generator=gen)
    out = ops.warp_corr_forward(x1, x2, fl, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1)
    if not use_mask:
        out = out.abs() + 1.0
    # oracle: LeakyReLU mask from the sign of `out` -- emulate "no mask" by a positive out (slope irrelevant then)
    r1, r2, rf = co.level_backward(x1.cpu().numpy(), x2.cpu().numpy(), fl.cpu().numpy() if use_flow else None, go.cpu().numpy(),
                                   4, 1, 4, 1, 1, 0, 0.1 if use_mask else 1.0)
    g1, g2, gf = ops.warp_corr_backward(x1, x2, fl, out, go, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1)
    torch.cuda.synchronize()
    print(f"{name}: g1 {rel(g1, r1):.2e} g2 {rel(g2, r2):.2e}" + (f" gflow {rel(gf, rf):.2e}" if use_flow else ""))
