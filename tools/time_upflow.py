#!/usr/bin/env python
"""Op-level benefit of fusing the flow up-sampling: per PWC level, CUDA-graph replays of
   [flow*2 -> F.interpolate -> fused warp+corr+act -> copy flow into the concat buffer]   vs
   [cerb_warp_corr_forward_upflow writing cost volume and flow into the concat buffer]."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, torch.nn.functional as F
import cerberusnet_b200 as cb
from cerberusnet_b200 import ops
dev = torch.device("cuda:0")
x = torch.randn(4096, 4096, device=dev); t_end = time.perf_counter() + 1.0
while time.perf_counter() < t_end: (x @ x).sum().item()
def graph_time(fn, reps=200):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn(); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(10): fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps // 10): g.replay()
        e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps
tot = [0.0, 0.0]
for (C, H, W) in ((128, 16, 32), (96, 32, 64), (64, 64, 128), (32, 128, 256)):
    x1 = F.leaky_relu(torch.randn(1, C, H, W, device=dev), 0.1); x2 = F.leaky_relu(torch.randn(1, C, H, W, device=dev), 0.1)
    coarse = torch.randn(1, 2, H // 2, W // 2, device=dev)
    cat = torch.empty(1, 81 + 32 + 2, H, W, device=dev)
    def unfused():
        fl = F.interpolate(coarse * 2, scale_factor=2, mode="bilinear", align_corners=True)
        ops.warp_corr_forward(x1, x2, fl, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1, out=cat[:, :81])
        cat[:, -2:] = fl
    def fused():
        ops.warp_corr_forward_upflow(x1, x2, coarse, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1, out=cat[:, :81], flow_up=cat[:, -2:])
    tu, tf = graph_time(unfused), graph_time(fused)
    tot[0] += tu; tot[1] += tf
    print(f"C={C} {H}x{W}: interpolate + op + copy {tu:6.2f} us   fused {tf:6.2f} us")
print(f"levels 1-4 total: {tot[0]:.1f} us -> {tot[1]:.1f} us")
