#!/usr/bin/env python
"""A/B timing on the GPU box: the reference's own CUDA op (oracle/_ref, compiled unmodified for
sm_100a) and the stock decoder composition (ATen grid_sample -> reference op -> leaky_relu_)
against the fused kernels, same shapes.  Test/bench infrastructure only.
    python tools/time_reference_op.py [--batch B] [--pyramid pwc|hrnet]"""
import argparse, json, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import torch, torch.nn.functional as F
torch.ops.load_library(os.path.join(ROOT, "oracle", "_ref", "correlation_ref.so"))
import cerberusnet_b200 as cb
from cerberusnet_b200 import ops
from oracle import torch_oracle as to
PYR = {"pwc": [(192, 8, 16, False), (128, 16, 32, True), (96, 32, 64, True), (64, 64, 128, True), (32, 128, 256, True)],
       "hrnet": [(384, 16, 32, False), (192, 32, 64, True), (96, 64, 128, True), (48, 128, 256, True)]}
ap = argparse.ArgumentParser(); ap.add_argument("--batch", type=int, default=1); ap.add_argument("--pyramid", default="pwc")
a = ap.parse_args(); B = a.batch; dev = torch.device("cuda:0")
x = torch.randn(4096, 4096, device=dev); t_end = time.perf_counter() + 1.0
while time.perf_counter() < t_end: (x @ x).sum().item()
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps
res = []
for li, (C, H, W, wp) in enumerate(PYR[a.pyramid]):
    x1 = F.leaky_relu(torch.randn(B, C, H, W, device=dev), 0.1); x2 = F.leaky_relu(torch.randn(B, C, H, W, device=dev), 0.1)
    fl = (torch.randn(B, 2, H, W, device=dev) * 1.5).clamp_(-6, 6) if wp else None
    out = torch.empty(B, 81, H, W, device=dev); g = torch.randn_like(out)
    def stock():
        w = to.flow_warp(x2, fl, to.WARP_TORCH) if wp else x2
        o = torch.ops.cerberus.correlation(x1, w, 4, 1, 4, 1, 1, 1)
        return F.leaky_relu(o, 0.1, inplace=True)
    r = {"level": li, "B": B, "C": C, "H": H, "W": W,
         "ours_fused_fwd_us": timeit(lambda: ops.warp_corr_forward(x1, x2, fl, 4, 1, 4, 1, 1, 1, 0, 0.1, out=out)),
         "ref_op_fwd_us": timeit(lambda: torch.ops.cerberus.correlation(x1, x2, 4, 1, 4, 1, 1, 1)),
         "stock_composition_fwd_us": timeit(stock),
         "ours_corr_bwd_us": timeit(lambda: ops.warp_corr_backward(x1, x2, None, None, g, 4, 1, 4, 1, 1, 1, 0, None)),
         "ref_op_bwd_us": timeit(lambda: torch.ops.cerberus.correlation_backward(x1, x2, g, 4, 1, 4, 1, 1, 1), reps=5),
         "ours_fused_bwd_us": timeit(lambda: ops.warp_corr_backward(x1, x2, fl, out, g, 4, 1, 4, 1, 1, 1, 0, 0.1))}
    res.append(r)
    print(json.dumps({k: (round(v, 1) if isinstance(v, float) else v) for k, v in r.items()}))
tot = {k: round(sum(r[k] for r in res), 1) for k in res[0] if k.endswith("_us")}
print("TOTAL", json.dumps(tot))
