"""CPU restatement of the reference's TensorRT grid-sampler plugin -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the product
(cerberusnet_b200/) never does.

Follows runtime/cerberus_net/trt_plugins/grid_sampler.cu:
    grid_sampler_unnormalize        :48-59    (align_corners: ((g+1)/2)*(size-1);  else  ((g+1)*(size-1))/2  -- NOT ATen's)
    clip_coordinates                :62-64
    reflect_coordinates             :73-86
    grid_sampler_compute_source_index :126-141
    bilinear / nearest bodies       :181-233  (taps outside the image are skipped; nearest uses roundf)
`convention="aten"` swaps in what the same ONNX node computes in PyTorch (F.grid_sample): ((g+1)*size-1)/2 and
round-half-to-even.  Pinned in tests/test_oracle_golden.py against torch.nn.functional.grid_sample on the CPU for
every (mode, padding, align_corners) combination; the TRT convention has no executable reference here (TensorRT
7.2 is absent), so it is pinned by the formulas above and by the identity  p = clamp(x + u)  that the flow-warp
oracle already uses (mode R).
All coordinate arithmetic is float32, like the kernel's.
"""
from __future__ import annotations

import numpy as np

BILINEAR, NEAREST = 0, 1
ZEROS, BORDER, REFLECTION = 0, 1, 2
f32 = np.float32


def _unnormalize(c, size, align, convention):
    c = c.astype(f32)
    if align:
        return ((c + f32(1)) * f32(0.5)) * f32(size - 1)
    if convention == "trt":
        return ((c + f32(1)) * f32(size - 1)) * f32(0.5)
    return ((c + f32(1)) * f32(size) - f32(1)) * f32(0.5)


def _reflect(x, twice_low, twice_high):
    if twice_low == twice_high:
        return np.zeros_like(x)
    mn = f32(twice_low) * f32(0.5)
    span = f32(twice_high - twice_low) * f32(0.5)
    x = np.abs(x - mn).astype(f32)
    extra = np.fmod(x, span).astype(f32)
    flips = np.floor(x / span).astype(np.int64)
    return np.where(flips % 2 == 0, extra + mn, span - extra + mn).astype(f32)


def _source_index(c, size, padding, align, convention):
    c = _unnormalize(c, size, align, convention)
    if padding == BORDER:
        c = np.clip(c, f32(0), f32(size - 1))
    elif padding == REFLECTION:
        c = _reflect(c, 0, 2 * (size - 1)) if align else _reflect(c, -1, 2 * size - 1)
        c = np.clip(c, f32(0), f32(size - 1))
    bad = ~np.isfinite(c) | (c > f32(2147483646.0)) | (c < f32(-2147483648.0))
    return np.where(bad, f32(-100.0), c).astype(f32)


def grid_sample(inp: np.ndarray, grid: np.ndarray, interpolation=BILINEAR, padding=BORDER, align_corners=False,
                convention="trt") -> np.ndarray:
    """inp (N,C,H,W), grid (N,oH,oW,2) with x first; returns (N,C,oH,oW) float64-accumulated, float32 coordinates."""
    N, C, H, W = inp.shape
    ix = _source_index(grid[..., 0], W, padding, align_corners, convention)
    iy = _source_index(grid[..., 1], H, padding, align_corners, convention)
    out = np.zeros((N, C) + grid.shape[1:3], np.float64)
    x = inp.astype(np.float64)
    n_idx = np.arange(N)[:, None, None]
    if interpolation == BILINEAR:
        x0 = np.floor(ix).astype(np.int64); y0 = np.floor(iy).astype(np.int64)
        x1, y1 = x0 + 1, y0 + 1
        wx1 = (x1.astype(f32) - ix).astype(f32); wx0 = (ix - x0.astype(f32)).astype(f32)
        wy1 = (y1.astype(f32) - iy).astype(f32); wy0 = (iy - y0.astype(f32)).astype(f32)
        for (xx, yy, w) in ((x0, y0, wx1 * wy1), (x1, y0, wx0 * wy1), (x0, y1, wx1 * wy0), (x1, y1, wx0 * wy0)):
            ok = (xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)
            xc, yc = np.clip(xx, 0, W - 1), np.clip(yy, 0, H - 1)
            v = x[n_idx, :, yc, xc]                    # (N, oH, oW, C)
            out += np.moveaxis(v * (w.astype(np.float64) * ok)[..., None], -1, 1)
    else:
        if convention == "trt":   # roundf: half away from zero
            rx = np.where(ix >= 0, np.floor(ix + f32(0.5)), np.ceil(ix - f32(0.5)))
            ry = np.where(iy >= 0, np.floor(iy + f32(0.5)), np.ceil(iy - f32(0.5)))
        else:                     # nearbyint: half to even
            rx, ry = np.rint(ix), np.rint(iy)
        xn, yn = rx.astype(np.int64), ry.astype(np.int64)
        ok = (xn >= 0) & (xn < W) & (yn >= 0) & (yn < H)
        v = x[n_idx, :, np.clip(yn, 0, H - 1), np.clip(xn, 0, W - 1)]
        out = np.moveaxis(v * ok[..., None], -1, 1)
    return out
