import os, sys, time
sys.path.insert(0, "/root/repo")
import torch, torch.nn.functional as F
from cerberusnet_b200 import ops
dev = torch.device("cuda:0")
x = torch.randn(4096, 4096, device=dev); t_end = time.perf_counter() + 1.0
while time.perf_counter() < t_end: (x @ x).sum().item()
def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps
B, C, H, W = 8, 48, 128, 256
for dt in (torch.float32, torch.bfloat16, torch.float16):
    x1 = F.leaky_relu(torch.randn(B, C, H, W, device=dev), 0.1).to(dt); x2 = F.leaky_relu(torch.randn(B, C, H, W, device=dev), 0.1).to(dt)
    fl = (torch.randn(B, 2, H, W, device=dev) * 1.5).clamp_(-6, 6)
    out = ops.warp_corr_forward(x1, x2, fl, 4, 1, 4, 1, 1, 1, 0, 0.1); g = torch.randn_like(out)
    tb = timeit(lambda: ops.warp_corr_backward(x1, x2, fl, out, g, 4, 1, 4, 1, 1, 1, 0, 0.1))
    tc = timeit(lambda: ops.warp_corr_backward(x1, x2, None, out, g, 4, 1, 4, 1, 1, 1, 0, 0.1))
    print(dt, "fused bwd", round(tb, 1), "us   corr-only bwd", round(tc, 1), "us")
