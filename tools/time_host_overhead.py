#!/usr/bin/env python
"""Host-side cost of one forward call (no sync inside the loop, tiny problem so the GPU is never the limit):
C ABI through ctypes with a prebuilt parameter block vs the Python ops wrapper."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import cerberusnet_b200 as cb
from cerberusnet_b200 import ops, _lib
dev = torch.device("cuda:0"); lib = cb.lib()
for (C, H, W) in ((32, 128, 256), (192, 8, 16)):
    x1 = torch.randn(1, C, H, W, device=dev); x2 = torch.randn(1, C, H, W, device=dev); fl = torch.randn(1, 2, H, W, device=dev)
    out = torch.empty(1, 81, H, W, device=dev)
    p = _lib.make_params(x1, x2, fl, out, 4, 1, 4, 1, 1, 1, 0, 0.1)
    sp = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    args = (ctypes.byref(p), _lib.ptr(x1), _lib.ptr(x2), _lib.ptr(fl), _lib.ptr(out), sp)
    for name, fn in (("C ABI via ctypes", lambda: lib.cerb_warp_corr_forward(*args)),
                     ("ops.warp_corr_forward(out=)", lambda: ops.warp_corr_forward(x1, x2, fl, 4, 1, 4, 1, 1, 1, 0, 0.1, out=out)),
                     ("cb.warp_correlation (allocating)", lambda: cb.warp_correlation(x1, x2, fl))):
        for _ in range(200): fn()
        torch.cuda.synchronize()
        n = 3000
        t0 = time.perf_counter()
        for _ in range(n): fn()
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        print(f"C={C} {H}x{W} {name:34s} host {1e6 * (t1 - t0) / n:6.2f} us/call   (incl. drain {1e6 * (t2 - t0) / n:6.2f})")
