import os, sys
sys.path.insert(0, "/root/repo")
import torch
import cerberusnet_b200 as cb
from cerberusnet_b200 import ops
dev = torch.device("cuda:0")
for (C, H, W) in ((64, 16, 32), (192, 8, 16), (96, 32, 64)):
    x1 = torch.randn(1, C, H, W, device=dev); x2 = torch.randn(1, C, H, W, device=dev); fl = torch.randn(1, 2, H, W, device=dev)
    ops.warp_corr_forward(x1, x2, fl, 4, 1, 4, 1, 1, 1, 0, 0.1)
    ops.warp_corr_forward(x1, x2, None, 4, 1, 4, 1, 1, 1, 0, 0.1)
torch.cuda.synchronize(); print("done")
