import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, torch.nn.functional as F
import cerberusnet_b200 as cb
from cerberusnet_b200 import ops
cb.lib().cerb_debug_set_backward_kernel(0)   # the fused CUDA-core kernel (AUTO takes the tensor-core kernel at this shape)
B, C, H, W = 8, 48, 128, 256
dev = torch.device("cuda:0")
x1 = F.leaky_relu(torch.randn(B, C, H, W, device=dev), 0.1); x2 = F.leaky_relu(torch.randn(B, C, H, W, device=dev), 0.1)
fl = (torch.randn(B, 2, H, W, device=dev) * 1.5).clamp_(-6, 6)
out = ops.warp_corr_forward(x1, x2, fl, 4, 1, 4, 1, 1, 1, 0, 0.1); g = torch.randn_like(out)
for _ in range(3):
    ops.warp_corr_backward(x1, x2, fl, out, g, 4, 1, 4, 1, 1, 1, 0, 0.1)
torch.cuda.synchronize()
