#!/usr/bin/env python
"""Per-CTA timeline of the fast forward kernel (clock64 trace): python tools/trace_level.py [level] [variant]"""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
import bench, cerberusnet_b200 as cb
from cerberusnet_b200 import ops
li = int(sys.argv[1]) if len(sys.argv) > 1 else 4
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0
BATCH = int(sys.argv[3]) if len(sys.argv) > 3 else 1
TITER = int(sys.argv[4]) if len(sys.argv) > 4 else 0
C, H, W, wp = bench.PWC_LEVELS[li]
dev = torch.device("cuda:0")
lib = cb.lib()
x1, x2, fl = bench.synth_level(li, C, H, W, wp, 1000, dev)
if os.environ.get("TRACE_FLOW") == "none":
    fl = None
elif os.environ.get("TRACE_FLOW") == "zero" and fl is not None:
    fl = torch.zeros_like(fl)
x1, x2 = x1.repeat(BATCH, 1, 1, 1), x2.repeat(BATCH, 1, 1, 1)
fl = fl.repeat(BATCH, 1, 1, 1) if fl is not None else None
out = torch.empty(BATCH, 81, H, W, device=dev)
lib.cerb_debug_set_trace_iter(TITER)
for _ in range(200):
    ops.warp_corr_forward(x1, x2, fl, 4, 1, 4, 1, 1, 1, 0, 0.1, out=out, variant=variant)
torch.cuda.synchronize()
SLOTS = 192
trace = torch.zeros(148 * SLOTS, dtype=torch.int64, device=dev)
lib.cerb_debug_set_trace_buffer(ctypes.c_void_p(trace.data_ptr()))
ops.warp_corr_forward(x1, x2, fl, 4, 1, 4, 1, 1, 1, 0, 0.1, out=out, variant=variant)
torch.cuda.synchronize()
lib.cerb_debug_set_trace_buffer(None)
t = trace.cpu().numpy().reshape(148, SLOTS)
names = {0: "start", 1: "prod:taps ready", 2: "prod:chunk0 full", 3: "prod:chunk1 full", 4: "prod:chunk2 full", 5: "prod:chunk3 full",
         16: "cons:begin", 17: "cons:c0 got", 18: "cons:c0 done", 19: "cons:c1 got", 20: "cons:c1 done", 21: "cons:c2 got",
         22: "cons:c2 done", 23: "cons:c3 got", 24: "cons:c3 done", 44: "prod:c0 start wait raw", 45: "prod:c0 raw landed", 46: "prod:c0 gather done", 47: "prod:c1 start wait raw", 48: "prod:c1 raw landed", 49: "prod:c1 gather done", 50: "prod:c2 start wait raw", 51: "prod:c2 raw landed", 52: "prod:c2 gather done", 60: "pt0:c1 loop top", 61: "pt0:c1 raw issued", 62: "pt0:c1 x1 issued", 56: "warp1:c1 loop top", 57: "warp1:c1 empty ok", 55: "warp1:c1 gather done", 58: "prod:flow+taps math done", 59: "prod:bbox reduced", 35: "epi:first barrier passed", 36: "epi:acc stored", 37: "epi:barrier 2 passed", 38: "epi:group reduce done", 39: "cons:partial written", 43: "cons:all partials in", 40: "cons:mainloop end", 41: "cons:store issued/reduced", 42: "cons:exit"}
for ck in range(6):
    for k, nm in enumerate(("loop top", "raw_empty ok", "raw issued", "x empty ok", "x1 issued")):
        names[64 + 8 * ck + k] = f"tma:c{ck} {nm}"
for ck in (2, 3):
    for gw in range(6):
        for k, nm in enumerate(("top", "empty ok", "raw_full ok", "gathered")):
            names[112 + (ck - 2) * 40 + gw * 6 + k] = f"g{gw}:c{ck} {nm}"
for cta in (0, 1):
    row = t[cta]
    if row[0] == 0:
        continue
    print(f"CTA {cta}:")
    for k in sorted(names, key=lambda k: row[k]):
        if row[k]:
            print(f"   {names[k]:24s} +{int(row[k] - row[0]):7d} cyc")
