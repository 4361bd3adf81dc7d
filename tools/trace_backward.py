"""clock64 trace of one CTA of each correlation-backward kernel: python tools/trace_backward.py B none|leaky flow|noflow"""
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
names = {0: "start (window 0)", 6: "G tile fetched + masked", 1: "chunk 0: barrier passed", 2: "chunk 0: S halo staged",
         3: "chunk 0: contraction done", 4: "chunk 0: output staged", 5: "all windows / chunks written"}
for w in (0, 1):
    print("kernel WHICH =", w, "(grad wrt x1)" if w == 0 else "(grad wrt the second correlation input)")
    for i, nm in sorted(names.items(), key=lambda kv: t[w * 32 + kv[0]]):
        print(f"   {nm:30s} +{int(t[w*32+i]-t[w*32]):8d} cyc")
