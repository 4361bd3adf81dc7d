import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import cerberusnet_b200 as cb
from cerberusnet_b200 import ops
dev = torch.device("cuda:0")
for (C, H, W) in ((64, 16, 32), (192, 8, 16), (96, 32, 64)):
    x1 = torch.randn(1, C, H, W, device=dev); x2 = torch.randn(1, C, H, W, device=dev); fl = torch.randn(1, 2, H, W, device=dev)
    ops.warp_corr_forward(x1, x2, fl, 4, 1, 4, 1, 1, 1, 0, 0.1)
    ops.warp_corr_forward(x1, x2, None, 4, 1, 4, 1, 1, 1, 0, 0.1)
# tensor-core forward: a few tiles per CTA, and the fused backward
x = torch.randn(2, 16, 32, 64, device=dev); fl = torch.randn(2, 2, 32, 64, device=dev)
out = ops.warp_corr_forward(x, x, fl, 4, 1, 4, 1, 1, 1, 0, 0.1, variant=7)
ops.warp_corr_backward(x, x, fl, out, torch.randn_like(out), 4, 1, 4, 1, 1, 1, 0, 0.1)
torch.cuda.synchronize(); print("done")
# tensor-core backward + window splat (shared-memory reductions, TMA reduce)
cb.lib().cerb_debug_set_backward_kernel(1)
ops.warp_corr_backward(x, x, fl, out, torch.randn_like(out), 4, 1, 4, 1, 1, 1, 0, 0.1)
cb.lib().cerb_debug_set_backward_kernel(-1)
torch.cuda.synchronize(); print("done")
