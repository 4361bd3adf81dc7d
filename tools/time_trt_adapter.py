#!/usr/bin/env python
"""BASELINE configs[2] / SURVEY 8d config 3, the part that can run without TensorRT: the plugin-`enqueue`-shaped entry
(cerb_trt_corr_enqueue: plain correlation node; cerb_trt_warp_corr_enqueue: fused warp + correlation + LeakyReLU node,
TensorRT un-normalisation) on the HRNetV2-W48 level shapes at 1024x512, batch 1, kFLOAT and kHALF, CUDA-graph replays
(the entry issues no synchronisation and no private streams, so an engine built on it is graph-capturable)."""
import ctypes, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch, torch.nn.functional as F
import cerberusnet_b200 as cb
from cerberusnet_b200 import _lib
dev = torch.device("cuda:0"); lib = cb.lib()
x = torch.randn(4096, 4096, device=dev); t_end = time.perf_counter() + 1.0
while time.perf_counter() < t_end: (x @ x).sum().item()
f = _lib.TrtCorrFields(); lib.cerb_trt_corr_default_fields(ctypes.byref(f))
def graph_time(fn, reps=200):
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        sp = ctypes.c_void_p(st.cuda_stream)
        fn(sp); torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(10): fn(sp)
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(reps // 10): g.replay()
        e1.record(st); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / reps
tot = {}
for (C, H, W, wp) in ((384, 16, 32, False), (192, 32, 64, True), (96, 64, 128, True), (48, 128, 256, True)):
    for name, dt, code in (("kFLOAT", torch.float32, 0), ("kHALF", torch.float16, 1)):
        x1 = F.leaky_relu(torch.randn(1, C, H, W, device=dev), 0.1).to(dt); x2 = F.leaky_relu(torch.randn(1, C, H, W, device=dev), 0.1).to(dt)
        fl = (torch.randn(1, 2, H, W, device=dev) * 1.5).clamp_(-6, 6)
        out = torch.empty(1, 81, H, W, device=dev, dtype=dt)
        descs = (_lib.TrtTensorDesc * 4)()
        for i, dims in enumerate(((1, C, H, W), (1, C, H, W), (1, 2, H, W), (1, 81, H, W))):
            descs[i].dims.nbDims = 4
            for j, v in enumerate(dims): descs[i].dims.d[j] = v
            descs[i].type = code if i != 2 else 0
        ins = (ctypes.c_void_p * 3)(x1.data_ptr(), x2.data_ptr(), fl.data_ptr()); outs = (ctypes.c_void_p * 1)(out.data_ptr())
        def plain(sp): assert lib.cerb_trt_corr_enqueue(ctypes.byref(f), descs, ctypes.byref(descs[3]), ins, outs, None, sp) == 0
        def fused(sp): assert lib.cerb_trt_warp_corr_enqueue(ctypes.byref(f), 1, 0.1, descs, ctypes.byref(descs[3]), ins, outs, None, sp) == 0
        tp = graph_time(plain); tfu = graph_time(fused) if wp else float("nan")
        tot[name] = tot.get(name, 0.0) + (tfu if wp else tp)
        print(f"C={C:3d} {H}x{W} {name}: correlation node {tp:6.2f} us   fused warp+corr+LeakyReLU node {tfu:6.2f} us")
print("pyramid (level 0 plain, levels 1-3 fused): " + ", ".join(f"{k} {v:.1f} us" for k, v in tot.items()))
