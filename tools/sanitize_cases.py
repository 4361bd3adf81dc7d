#!/usr/bin/env python
"""Small cases that touch every kernel path, for compute-sanitizer (memcheck / racecheck / synccheck)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import cerberusnet_b200 as cb
from cerberusnet_b200 import ops
dev = torch.device("cuda:0")
torch.manual_seed(0)
def run(B, C, H, W, flow, variant=0, pad=4, sigma=2.0, md=4, dtype=torch.float32):
    x1 = torch.randn(B, C, H, W, device=dev).to(dtype); x2 = torch.randn(B, C, H, W, device=dev).to(dtype)
    fl = torch.randn(B, 2, H, W, device=dev) * sigma if flow else None
    out = ops.warp_corr_forward(x1, x2, fl, pad, 1, md, 1, 1, 1, cb.WARP_TORCH, 0.1, variant=variant)
    g = torch.randn_like(out)
    ops.warp_corr_backward(x1, x2, fl, out, g, pad, 1, md, 1, 1, 1, cb.WARP_TORCH, 0.1)
    torch.cuda.synchronize()
for flow in (False, True):
    run(1, 32, 128, 256, flow)            # 8x32 config, TMA everywhere (one tile per CTA)
    run(2, 20, 24, 64, flow, variant=1)   # 8x32, partial channel chunk, several tiles per CTA not needed
    run(1, 64, 16, 32, flow)              # 4x16 config with cluster split
    run(1, 192, 8, 16, flow)              # 4x16, 8-CTA clusters
    run(3, 7, 9, 50, flow)                # ragged: no TMA (W % 4 != 0)
    run(2, 12, 26, 28, flow, pad=2)       # pad != md: LDG staging of x1
run(1, 16, 32, 64, True, sigma=20.0)      # raw box does not fit: direct-gather fallback
for flow in (False, True):
    run(1, 12, 24, 64, flow, md=8, pad=8)                 # displacement windows, forward and backward
    run(1, 12, 26, 44, flow, md=10, variant=3)            # three windows per axis, pad != md, 4x16 kernel
    for dt in (torch.float16, torch.bfloat16):
        run(1, 32, 40, 64, flow, dtype=dt, variant=1)     # 16-bit TMA raw path, 8x32
        run(1, 24, 16, 32, flow, dtype=dt)                # 16-bit, 4x16 with cluster split
        run(2, 8, 19, 37, flow, dtype=dt)                 # 16-bit, no TMA (W % 8 != 0)
run(1, 16, 32, 64, True, sigma=20.0, dtype=torch.bfloat16, variant=1)   # 16-bit raw box misfit
# fused flow up-sampling (plain and cluster-split kernels, md 4 and 8), writing into a concat-buffer slice
for (C, H, W, md) in ((16, 32, 64, 4), (48, 16, 32, 4), (8, 24, 40, 8)):
    a = torch.randn(1, C, H, W, device=dev); coarse = torch.randn(1, 2, H // 2, W // 2, device=dev)
    cat = torch.zeros(1, (2 * md + 1) ** 2 + 2, H, W, device=dev)
    ops.warp_corr_forward_upflow(a, a, coarse, md, 1, md, 1, 1, 1, cb.WARP_TORCH, 0.1, out=cat[:, :-2], flow_up=cat[:, -2:])
# more tiles than CTAs: persistent loop, per-column stores, 16-bit without a flow
for dt in (torch.float32, torch.float16):
    a = torch.randn(5, 4, 128, 256, device=dev).to(dt); fl = torch.randn(5, 2, 128, 256, device=dev)
    ops.warp_corr_forward(a, a, fl, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1, variant=1)
    ops.warp_corr_forward(a, a, None, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1, variant=1)
torch.cuda.synchronize()
# tensor-core forward (variant 7): persistent loop over several tiles, 48 K steps, ragged / unaligned, pad != md, direct gather,
# no flow, fused up-sampling; and the fused backward at a shape with several tiles per image
a = torch.randn(3, 16, 64, 160, device=dev); fl = torch.randn(3, 2, 64, 160, device=dev) * 2
ops.warp_corr_forward(a, a, fl, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1, variant=7, x2_roll=1)   # 120 tiles > CTAs on small parts only
ops.warp_corr_forward(torch.randn(6, 8, 128, 256, device=dev), torch.randn(6, 8, 128, 256, device=dev),
                      torch.randn(6, 2, 128, 256, device=dev), 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1, variant=7)   # 768 tiles: 5-6 per CTA
run(1, 384, 16, 32, True, variant=7)
run(2, 13, 21, 45, True, variant=7)
run(1, 24, 30, 52, True, variant=7, pad=2)
run(1, 24, 30, 52, True, variant=7, pad=6)
run(1, 16, 32, 64, True, variant=7, sigma=20.0)
run(2, 32, 24, 64, False, variant=7)
a = torch.randn(2, 16, 64, 128, device=dev); coarse = torch.randn(2, 2, 32, 64, device=dev)
cat = torch.zeros(2, 83, 64, 128, device=dev)
ops.warp_corr_forward_upflow(a, a, coarse, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1, out=cat[:, :-2], flow_up=cat[:, -2:])   # 64 tiles -> tensor-core path
torch.cuda.synchronize()
x = torch.randn(1, 6, 12, 20, device=dev); f = torch.randn(1, 2, 12, 20, device=dev) * 3
o = ops.flow_warp_forward(x, f); ops.flow_warp_backward(x, f, torch.randn_like(o))
ops.warp_corr_forward(x, x, f, 3, 3, 4, 2, 2); torch.cuda.synchronize()
# tensor-core backward (banded GEMMs in TMEM) + shared-memory-window splat: windows that fit, a wild flow (scattered atomics),
# ragged tiles, C = 64 / 96 (two / one partial accumulators), no flow, the stand-alone flow_warp backward through the windows
cb.lib().cerb_debug_set_backward_kernel(1)
run(2, 16, 24, 64, True)
run(1, 48, 40, 96, True, sigma=1.5)
run(3, 20, 17, 36, True)
run(1, 32, 16, 48, True, sigma=14.0)
run(1, 64, 16, 32, True)
run(1, 96, 16, 32, False)
cb.lib().cerb_debug_set_backward_kernel(-1)
x = torch.randn(2, 12, 24, 48, device=dev); f = torch.randn(2, 2, 24, 48, device=dev) * 2
ops.flow_warp_backward(x, f, torch.randn_like(x)); torch.cuda.synchronize()
print("sanitize cases done")
