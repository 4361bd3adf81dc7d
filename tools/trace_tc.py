#!/usr/bin/env python
"""Per-CTA timeline of the tensor-core forward (clock64 trace): python tools/trace_tc.py [level] [batch]"""
import ctypes, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench, cerberusnet_b200 as cb
from cerberusnet_b200 import ops
li = int(sys.argv[1]) if len(sys.argv) > 1 else 4
BATCH = int(sys.argv[2]) if len(sys.argv) > 2 else 2
C, H, W, wp = bench.PWC_LEVELS[li]
dev = torch.device("cuda:0")
lib = cb.lib()
x1, fl = bench.synth_level(li, C, H, W, wp, 1000, dev, B=BATCH)
out = torch.empty(BATCH, 81, H, W, device=dev)
for _ in range(200):
    ops.warp_corr_forward(x1, x1, fl, 4, 1, 4, 1, 1, 1, 0, 0.1, out=out, variant=7, x2_roll=BATCH // 2)
torch.cuda.synchronize()
SLOTS = 192
trace = torch.zeros(148 * SLOTS, dtype=torch.int64, device=dev)
lib.cerb_debug_set_trace_buffer(ctypes.c_void_p(trace.data_ptr()))
ops.warp_corr_forward(x1, x1, fl, 4, 1, 4, 1, 1, 1, 0, 0.1, out=out, variant=7, x2_roll=BATCH // 2)
torch.cuda.synchronize()
lib.cerb_debug_set_trace_buffer(None)
t = trace.cpu().numpy().reshape(148, SLOTS)
names = {0: "gather: tile start", 1: "gather: taps done", 2: "gather: bbox done", 15: "tma: bbox read", 20: "mma: d_empty ok",
         25: "mma: d_full committed", 30: "epi: d_full ok", 31: "epi: tmem drained (d_empty)", 32: "epi: stores done"}
for ks in range(4):
    names[3 + 3 * ks] = f"gather: k{ks} raw landed"
    names[4 + 3 * ks] = f"gather: k{ks} slot free"
    names[5 + 3 * ks] = f"gather: k{ks} published"
    names[16 + ks] = f"tma: k{ks} issued"
    names[21 + ks] = f"mma: k{ks} operands ready"
    names[26 + ks] = f"epi: k{ks} x1 staged"
for cta in (0, 77):
    row = t[cta]
    print(f"CTA {cta}: end +{int(row[191] - row[0])} cyc")
    ev = []
    for ti in range(4):
        for s, nm in names.items():
            v = row[1 + ti * 40 + s]
            if v:
                ev.append((int(v - row[0]), f"t{ti} {nm}"))
    for c, nm in sorted(ev):
        print(f"   +{c:7d}  {nm}")
