#!/usr/bin/env python
"""Developer check on a GPU box: every forward variant, the backward and the stand-alone warp
against the CPU oracle, with errors printed per case (pytest gives pass/fail; this gives numbers).

    python tools/gpu_quickcheck.py [--time]
"""
from __future__ import annotations

import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))

import cerberusnet_b200 as cb  # noqa: E402
from cerberusnet_b200 import _lib, ops  # noqa: E402
from oracle import c_oracle as co  # noqa: E402

VARIANTS = {1: "fast8x32+tma", 2: "fast8x32", 3: "small4x16+tma", 4: "small4x16", 5: "generic"}


def rel(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--time", action="store_true")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    print(torch.cuda.get_device_name(0))
    rs = np.random.RandomState(0)
    worst = 0.0
    shapes = [(1, 32, 16, 32), (2, 20, 13, 37), (1, 64, 64, 128), (1, 192, 8, 16), (3, 7, 9, 50), (1, 48, 24, 64)]
    for (B, C, H, W) in shapes:
        x1 = rs.standard_normal((B, C, H, W)).astype(np.float32)
        x2 = rs.standard_normal((B, C, H, W)).astype(np.float32)
        fl = (rs.standard_normal((B, 2, H, W)) * 2.5).astype(np.float32)
        t1, t2, tf = (torch.from_numpy(a).to(dev) for a in (x1, x2, fl))
        for flow_on in (False, True):
            for mode in ((0, 1) if flow_on else (0,)):
                ref = co.level_forward(x1, x2, fl if flow_on else None, 4, 1, 4, 1, 1, mode, 0.1)
                for v, name in VARIANTS.items():
                    try:
                        out = ops.warp_corr_forward(t1, t2, tf if flow_on else None, 4, 1, 4, 1, 1, 1, mode, 0.1,
                                                    variant=v)
                        torch.cuda.synchronize()
                        e = rel(out.cpu().numpy(), ref)
                    except Exception as ex:  # noqa: BLE001
                        print(f"  {name}: EXC {ex}")
                        e = float("inf")
                    worst = max(worst, e)
                    flag = "" if e < 1e-5 else "   <-- FAIL"
                    print(f"fwd {B}x{C}x{H}x{W} flow={int(flow_on)} mode={mode} {name:14s} rel={e:.2e}{flag}")
        # backward (auto path) vs oracle adjoint
        for flow_on in (False, True):
            g = rs.standard_normal(ref.shape).astype(np.float32)
            tg = torch.from_numpy(g).to(dev)
            out = ops.warp_corr_forward(t1, t2, tf if flow_on else None, 4, 1, 4, 1, 1, 1, 0, 0.1)
            g1, g2, gf = ops.warp_corr_backward(t1, t2, tf if flow_on else None, out, tg, 4, 1, 4, 1, 1, 1, 0, 0.1)
            torch.cuda.synchronize()
            r1, r2, rf = co.level_backward(x1, x2, fl if flow_on else None, g, 4, 1, 4, 1, 1, 0, 0.1)
            e1, e2 = rel(g1.cpu().numpy(), r1), rel(g2.cpu().numpy(), r2)
            ef = rel(gf.cpu().numpy(), rf) if flow_on else 0.0
            worst = max(worst, e1, e2, ef)
            print(f"bwd {B}x{C}x{H}x{W} flow={int(flow_on)} g1={e1:.2e} g2={e2:.2e} gflow={ef:.2e}")
    # generic parameters
    for (p, k, md, s1, s2) in [(4, 1, 10, 1, 1), (3, 3, 4, 2, 2), (2, 1, 4, 1, 2), (6, 3, 4, 1, 1), (5, 1, 4, 2, 1), (8, 1, 8, 1, 1)]:
        B, C, H, W = 2, 12, 20, 28
        x1 = rs.standard_normal((B, C, H, W)).astype(np.float32)
        x2 = rs.standard_normal((B, C, H, W)).astype(np.float32)
        fl = (rs.standard_normal((B, 2, H, W)) * 2.5).astype(np.float32)
        t1, t2, tf = (torch.from_numpy(a).to(dev) for a in (x1, x2, fl))
        for flow_on in (False, True):
            ref = co.level_forward(x1, x2, fl if flow_on else None, p, k, md, s1, s2, 0, 0.1)
            out = ops.warp_corr_forward(t1, t2, tf if flow_on else None, p, k, md, s1, s2, 1, 0, 0.1)
            g = rs.standard_normal(ref.shape).astype(np.float32)
            g1, g2, gf = ops.warp_corr_backward(t1, t2, tf if flow_on else None, out, torch.from_numpy(g).to(dev), p, k,
                                                md, s1, s2, 1, 0, 0.1)
            torch.cuda.synchronize()
            r1, r2, rf = co.level_backward(x1, x2, fl if flow_on else None, g, p, k, md, s1, s2, 0, 0.1)
            e = [rel(out.cpu().numpy(), ref), rel(g1.cpu().numpy(), r1), rel(g2.cpu().numpy(), r2),
                 rel(gf.cpu().numpy(), rf) if flow_on else 0.0]
            worst = max(worst, *e)
            print(f"generic p{p} k{k} md{md} s1{s1} s2{s2} flow={int(flow_on)} out={e[0]:.2e} g1={e[1]:.2e} g2={e[2]:.2e} gf={e[3]:.2e}")
    # stand-alone warp
    for mode in (0, 1):
        img = rs.standard_normal((2, 6, 11, 23)).astype(np.float32)
        fl = (rs.standard_normal((2, 2, 11, 23)) * 4).astype(np.float32)
        g = rs.standard_normal(img.shape).astype(np.float32)
        o = ops.flow_warp_forward(torch.from_numpy(img).to(dev), torch.from_numpy(fl).to(dev), mode)
        gi, gf = ops.flow_warp_backward(torch.from_numpy(img).to(dev), torch.from_numpy(fl).to(dev),
                                        torch.from_numpy(g).to(dev), mode)
        ri, rf = co.flow_warp_backward(img, fl, g, mode)
        e = [rel(o.cpu().numpy(), co.flow_warp_forward(img, fl, mode)), rel(gi.cpu().numpy(), ri), rel(gf.cpu().numpy(), rf)]
        worst = max(worst, *e)
        print(f"warp mode={mode} out={e[0]:.2e} gimg={e[1]:.2e} gflow={e[2]:.2e}")
    print(f"WORST rel error {worst:.3e}")

    if args.time:
        pyr = [(192, 8, 16), (128, 16, 32), (96, 32, 64), (64, 64, 128), (32, 128, 256)]
        for (C, H, W) in pyr:
            t1 = torch.randn(1, C, H, W, device=dev)
            t2 = torch.randn(1, C, H, W, device=dev)
            tf = torch.randn(1, 2, H, W, device=dev) * 1.5
            for v in (0, 1, 2, 3, 4):
                out = ops.warp_corr_forward(t1, t2, tf, 4, 1, 4, 1, 1, 1, 0, 0.1, variant=v)
                g = torch.cuda.CUDAGraph()
                s = torch.cuda.Stream()
                with torch.cuda.stream(s):
                    for _ in range(3):
                        ops.warp_corr_forward(t1, t2, tf, 4, 1, 4, 1, 1, 1, 0, 0.1, out=out, variant=v)
                    torch.cuda.synchronize()
                    with torch.cuda.graph(g, stream=s):
                        for _ in range(20):
                            ops.warp_corr_forward(t1, t2, tf, 4, 1, 4, 1, 1, 1, 0, 0.1, out=out, variant=v)
                g.replay()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    g.replay()
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) * 1000 / 200
                byts = H * W * 4 * (2 * C + 81 + 2)
                print(f"time C{C} {H}x{W} variant {v}: {us:.2f} us  ({byts / us / 1e3:.0f} GB/s algorithmic, L2-resident)")
    return 0 if worst < 1e-5 else 1


if __name__ == "__main__":
    sys.exit(main())
