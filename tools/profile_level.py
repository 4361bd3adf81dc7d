#!/usr/bin/env python
"""Run one pyramid level a few times for ncu (both flow directions per launch, as the bench does).
    python tools/profile_level.py LEVEL VARIANT ITERS [rotate=1] [flow=1]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import bench
from cerberusnet_b200 import ops
import cerberusnet_b200 as cb

li = int(sys.argv[1]) if len(sys.argv) > 1 else 4
variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 30
rotate = int(sys.argv[4]) if len(sys.argv) > 4 else 1
use_flow = int(sys.argv[5]) if len(sys.argv) > 5 else 1
C, H, W, wp = bench.PWC_LEVELS[li]
dev = torch.device("cuda:0")
sets = []
for s in range(10 if rotate else 1):
    f, fl = bench.synth_level(li, C, H, W, True, 1000 + 100 * s, dev)
    sets.append((f, fl if use_flow else None, torch.empty(bench.DIRS, 81, H, W, device=dev)))
for i in range(iters):
    f, fl, out = sets[i % len(sets)]
    ops.warp_corr_forward(f, f, fl, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1, out=out, variant=variant, x2_roll=bench.DIRS // 2)
torch.cuda.synchronize()
print("done")
