import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, torch.nn.functional as F
import cerberusnet_b200 as cb
from cerberusnet_b200 import ops
B, C, H, W = int(sys.argv[1]), 48, 128, 256
SLOPE = None if sys.argv[2] == 'none' else 0.1
FLOW = sys.argv[3] == 'flow'
dev = torch.device("cuda:0"); lib = cb.lib()
x1 = F.leaky_relu(torch.randn(B, C, H, W, device=dev), 0.1); x2 = F.leaky_relu(torch.randn(B, C, H, W, device=dev), 0.1)
fl = (torch.randn(B, 2, H, W, device=dev) * 1.5).clamp_(-6, 6) if FLOW else None
out = ops.warp_corr_forward(x1, x2, fl, 4, 1, 4, 1, 1, 1, 0, SLOPE); g = torch.randn_like(out)
for _ in range(20): ops.warp_corr_backward(x1, x2, fl, out, g, 4, 1, 4, 1, 1, 1, 0, SLOPE)
torch.cuda.synchronize()
tr = torch.zeros(64, dtype=torch.int64, device=dev)
lib.cerb_debug_set_trace_buffer(ctypes.c_void_p(tr.data_ptr()))
ops.warp_corr_backward(x1, x2, fl, out, g, 4, 1, 4, 1, 1, 1, 0, SLOPE); torch.cuda.synchronize()
lib.cerb_debug_set_trace_buffer(None)
t = tr.cpu().numpy()
names = ["start", "G tile built", "S chunk0 staged", "chunk0 contraction done", "chunk0 out staged", "all chunks written", "G batch0", "G batch1", "G batch2"]
for w in (0, 1):
    print("kernel WHICH =", w)
    for i, nm in enumerate(names):
        print(f"   {nm:26s} +{int(t[w*32+i]-t[w*32]):8d} cyc")
