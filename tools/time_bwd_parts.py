#!/usr/bin/env python
"""Backward of one level split into its launches (CUDA events around each variant of the call):
    python tools/time_bwd_parts.py [B] [C] [H] [W]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, torch.nn.functional as F
import cerberusnet_b200 as cb
from cerberusnet_b200 import ops
B, C, H, W = [int(v) for v in (sys.argv[1:5] + ["8", "48", "128", "256"][len(sys.argv) - 1:])]
dev = torch.device("cuda:0")
x1 = F.leaky_relu(torch.randn(B, C, H, W, device=dev), 0.1); x2 = F.leaky_relu(torch.randn(B, C, H, W, device=dev), 0.1)
fl = (torch.randn(B, 2, H, W, device=dev) * 1.5).clamp_(-6, 6)
out = ops.warp_corr_forward(x1, x2, fl, 4, 1, 4, 1, 1, 1, 0, 0.1); g = torch.randn_like(out)
def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps
gw = torch.randn_like(x2)
print(f"B={B} C={C} {H}x{W}")
print(f"  fused backward (flow, leaky)    {timeit(lambda: ops.warp_corr_backward(x1, x2, fl, out, g, 4, 1, 4, 1, 1, 1, 0, 0.1)):8.1f} us")
print(f"  corr backward only (leaky)      {timeit(lambda: ops.warp_corr_backward(x1, x2, None, out, g, 4, 1, 4, 1, 1, 1, 0, 0.1)):8.1f} us")
print(f"  corr backward only (no mask)    {timeit(lambda: ops.warp_corr_backward(x1, x2, None, None, g, 4, 1, 4, 1, 1, 1, 0, None)):8.1f} us")
print(f"  flow_warp forward               {timeit(lambda: ops.flow_warp_forward(x2, fl)):8.1f} us")
print(f"  flow_warp backward (+memset)    {timeit(lambda: ops.flow_warp_backward(x2, fl, gw)):8.1f} us")
