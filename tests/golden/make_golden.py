#!/usr/bin/env python
"""Generate tests/golden/*.npz by EXECUTING the reference's own Python for the hot path.

Run in the build container only (needs /root/reference; the GPU box never has it):

    python tests/golden/make_golden.py [--ref /root/reference]

Nothing is copied from the reference: its source is loaded from where it lies --
  * ``CorrelationTorch``   nnet_training/correlation_package/correlation.py:4-21, taken out by AST
    (importing that module fails: its line 2 loads a py3.8 .so by relative path);
  * ``flow_warp`` & co     nnet_training/loss_functions/UnFlowLoss.py:11-32,83-94, imported as a
    stand-alone two-file package (UnFlowLoss.py + loss_functions.py, both pure torch).
Outputs are the reference results (and autograd gradients of the reference composition
warp -> correlation -> leaky_relu(0.1), nnet_models/pwcnet_sfd.py:171-182) on seeded inputs.
Inputs are NOT stored: tests regenerate them with ``golden_inputs`` below (numpy RandomState is
bit-stable across platforms and versions).
"""
from __future__ import annotations

import argparse
import ast
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))

# name -> (seed, B, C, H, W, flow_sigma or None)
CASES = {
    "corr_small": (11, 1, 16, 12, 20, None),
    "corr_c48": (12, 2, 48, 9, 16, None),          # C not a multiple of 32
    "corr_pwc_l0": (13, 1, 192, 8, 16, None),      # coarsest PWC level, no warp
    "corr_config1": (0, 1, 64, 64, 128, None),     # BASELINE.json configs[0]; stored subsampled
    "warp_small": (21, 2, 8, 12, 20, 3.0),
    "warp_zero_flow": (22, 1, 4, 9, 16, 0.0),
    "level_small": (31, 1, 32, 16, 32, 1.5),
    "level_stress": (32, 2, 24, 10, 24, 12.0),     # flow well past every border (3 * md)
}


def golden_inputs(name: str):
    """Seeded inputs of a case: (x1, x2, flow-or-None) float32 numpy, features post-LeakyReLU-like."""
    seed, B, C, H, W, sigma = CASES[name]
    rs = np.random.RandomState(seed)
    x1 = rs.standard_normal((B, C, H, W)).astype(np.float32)
    x2 = rs.standard_normal((B, C, H, W)).astype(np.float32)
    flow = None
    if sigma is not None:
        flow = (rs.standard_normal((B, 2, H, W)) * sigma).astype(np.float32)
    return x1, x2, flow


def load_reference(ref_root: str):
    corr_py = os.path.join(ref_root, "nnet_training", "correlation_package", "correlation.py")
    tree = ast.parse(open(corr_py).read())
    cls = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "CorrelationTorch"]
    assert len(cls) == 1
    ns = {"torch": torch}
    exec(compile(ast.Module(body=cls, type_ignores=[]), corr_py, "exec"), ns)
    CorrelationTorch = ns["CorrelationTorch"]

    lf_dir = os.path.join(ref_root, "nnet_training", "loss_functions")
    pkg = types.ModuleType("_ref_lf")
    pkg.__path__ = [lf_dir]
    sys.modules["_ref_lf"] = pkg
    for sub in ("loss_functions", "UnFlowLoss"):
        spec = importlib.util.spec_from_file_location(f"_ref_lf.{sub}", os.path.join(lf_dir, sub + ".py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"_ref_lf.{sub}"] = mod
        spec.loader.exec_module(mod)
    flow_warp = sys.modules["_ref_lf.UnFlowLoss"].flow_warp
    return CorrelationTorch, flow_warp


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    args = ap.parse_args()
    torch.manual_seed(0)
    torch.set_num_threads(8)
    CorrelationTorch, flow_warp = load_reference(args.ref)
    corr = CorrelationTorch(max_displacement=4)

    for name in CASES:
        x1, x2, flow = golden_inputs(name)
        out = {}
        if name.startswith("corr_"):
            y = corr(torch.from_numpy(x1), torch.from_numpy(x2)).numpy()
            if name == "corr_config1":
                out["out_sub4"] = y[:, :, ::4, ::4].copy()
                out["plane_sums_f64"] = y.astype(np.float64).sum(axis=(2, 3))
            else:
                out["out"] = y
        elif name.startswith("warp_"):
            img = torch.from_numpy(x1).requires_grad_()
            fl = torch.from_numpy(flow).requires_grad_()
            y = flow_warp(img, fl)
            g = np.random.RandomState(1000 + CASES[name][0]).standard_normal(y.shape).astype(np.float32)
            y.backward(torch.from_numpy(g))
            out["out"] = y.detach().numpy()
            out["grad_image"] = img.grad.numpy()
            out["grad_flow"] = fl.grad.numpy()
        else:  # level_*: the decoder's hot-path composition
            t1 = torch.from_numpy(x1).requires_grad_()
            t2 = torch.from_numpy(x2).requires_grad_()
            fl = torch.from_numpy(flow).requires_grad_()
            y = torch.nn.functional.leaky_relu(corr(t1, flow_warp(t2, fl)), 0.1)
            g = np.random.RandomState(1000 + CASES[name][0]).standard_normal(y.shape).astype(np.float32)
            y.backward(torch.from_numpy(g))
            out["out"] = y.detach().numpy()
            out["grad_x1"] = t1.grad.numpy()
            out["grad_x2"] = t2.grad.numpy()
            out["grad_flow"] = fl.grad.numpy()
        path = os.path.join(HERE, name + ".npz")
        np.savez(path, **out)
        print(f"{name}: {', '.join(f'{k}{tuple(v.shape)}' for k, v in out.items())} -> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
