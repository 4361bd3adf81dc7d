#!/usr/bin/env python
"""Per-CTA timeline of the tensor-core backward (clock64 trace; build with -DCERB_BTC_TRACE):
   python tools/trace_bwd_tc.py [C H W B]"""
import ctypes, os, sys
os.environ.setdefault("CERB_DEBUG_BWD_TC", "1")
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.nn.functional as F
import cerberusnet_b200 as cb
from cerberusnet_b200 import ops
C, H, W, B = (int(v) for v in sys.argv[1:5]) if len(sys.argv) > 4 else (48, 128, 256, 8)
dev = torch.device("cuda:0")
lib = cb.lib()
g = torch.Generator(device=dev).manual_seed(5)
x1 = F.leaky_relu(torch.randn(B, C, H, W, device=dev, generator=g), 0.1)
x2 = F.leaky_relu(torch.randn(B, C, H, W, device=dev, generator=g), 0.1)
fl = (torch.randn(B, 2, H, W, device=dev, generator=g) * 1.5).clamp_(-6, 6)
out = ops.warp_corr_forward(x1, x2, fl, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1)
go = torch.randn(B, 81, H, W, device=dev, generator=g)
for _ in range(5):
    ops.warp_corr_backward(x1, x2, fl, out, go, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1)
torch.cuda.synchronize()
SLOTS = 192
trace = torch.zeros(148 * SLOTS, dtype=torch.int64, device=dev)
lib.cerb_debug_set_trace_buffer(ctypes.c_void_p(trace.data_ptr()))
ops.warp_corr_backward(x1, x2, fl, out, go, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1)
torch.cuda.synchronize()
lib.cerb_debug_set_trace_buffer(None)
t = trace.cpu().numpy().reshape(148, SLOTS)
names = {}
for X, gname in ((0, "gA"), (1, "gB")):
    for s, nm in enumerate(["build start", "chunk 3 built", "build end", "taps / bbox done", "next Gx staged", "d_full ok", "drain end"]):
        names[X * 12 + s] = f"{gname}: {nm}"
    names[24 + X] = f"mma {gname}: r0 operands ready"
    names[26 + X] = f"mma {gname}: r8 operands ready"
    names[28 + X] = f"mma {gname}: r15 operands ready"
    names[30 + X] = f"mma {gname}: d_full committed"
    names[32 + X] = f"tma {gname}: r0 issue"
    names[34 + X] = f"tma {gname}: r15 issue"
    names[36 + X] = f"split {gname}: r0 landed"
    names[38 + X] = f"split {gname}: r15 landed"
for cta in (0, 77):
    row = t[cta]
    print(f"CTA {cta}: end +{int(row[191] - row[0])} cyc")
    ev = []
    for ti in range(3):
        for s, nm in names.items():
            v = row[1 + ti * 40 + s]
            if v:
                ev.append((int(v - row[0]), f"t{ti} {nm}"))
    for c, nm in sorted(ev):
        print(f"   +{c:7d}  {nm}")
