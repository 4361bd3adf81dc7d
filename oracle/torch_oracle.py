"""Pure-PyTorch restatement of the cost-volume hot path (CPU baseline + autograd oracle).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  This is also the "pure-PyTorch
implementation of the same op" that BASELINE.json's north_star asks to be timed on the host
cores (bench.py ``cpu_baseline`` and ``--impl reference``).

Written from SURVEY.md section 8(a); reference sites (relative to /root/reference):
  correlation semantics   nnet_training/correlation_package/correlation_cuda_kernel.cu:29-95,
                          output dims correlation_cuda.cpp:6-14, CPU form correlation.py:4-21
  flow_warp               nnet_training/loss_functions/UnFlowLoss.py:11-32,83-94
  TRT grid sampler        runtime/cerberus_net/trt_plugins/grid_sampler.cu:48-59,181-217
  LeakyReLU(0.1)          nnet_training/nnet_models/pwcnet_sfd.py:182
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn.functional as F

WARP_TORCH = 0      # whatever ATen does on the tensor's device (CUDA: reciprocal multiply, CPU: division)
WARP_TRT = 1
WARP_TORCH_CPU = 2  # alias of WARP_TORCH for CPU tensors


def corr_out_dims(H: int, W: int, pad: int, k: int, md: int, s1: int, s2: int):
    kr = (k - 1) // 2
    border = kr + md
    r = md // s2
    d = 2 * r + 1
    oh = math.ceil((H + 2 * pad - 2 * border) / s1)
    ow = math.ceil((W + 2 * pad - 2 * border) / s1)
    return d * d, oh, ow


def correlation(x1: torch.Tensor, x2: torch.Tensor, pad_size: int, kernel_size: int, max_displacement: int,
                stride1: int, stride2: int, corr_multiply: int = 1) -> torch.Tensor:
    """General-parameter correlation; differentiable, so autograd gives the backward oracle.

    out[n,(tj+r)*D+(ti+r),by,bx] = 1/(k*k*C) * sum_{j,i,c} P1[n,c,by*s1+md+j,bx*s1+md+i]
                                                     * P2[n,c,by*s1+md+tj*s2+j,bx*s1+md+ti*s2+i]
    with P = zero-padded input (taps beyond the padded buffer also read zero).  corr_multiply is
    accepted and ignored, like the reference (correlation_cuda_kernel.cu:246-247).
    """
    B, C, H, W = x1.shape
    p, k, md, s1, s2 = pad_size, kernel_size, max_displacement, stride1, stride2
    d2, oh, ow = corr_out_dims(H, W, p, k, md, s1, s2)
    if oh <= 0 or ow <= 0:
        raise ValueError("correlation: empty output")
    kr = (k - 1) // 2
    r = md // s2
    extra = kr + r * s2  # guard band so every slice below stays non-negative / in range
    P1 = F.pad(x1, [p + extra] * 4)
    P2 = F.pad(x2, [p + extra] * 4)
    span_y, span_x = (oh - 1) * s1 + 1, (ow - 1) * s1 + 1
    planes = []
    for tj in range(-r, r + 1):
        for ti in range(-r, r + 1):
            acc = None
            for j in range(-kr, kr + 1):
                for i in range(-kr, kr + 1):
                    ya, xa = extra + md + j, extra + md + i
                    yb, xb = ya + tj * s2, xa + ti * s2
                    a = P1[:, :, ya:ya + span_y:s1, xa:xa + span_x:s1]
                    b = P2[:, :, yb:yb + span_y:s1, xb:xb + span_x:s1]
                    term = (a * b).sum(dim=1, keepdim=True)
                    acc = term if acc is None else acc + term
            planes.append(acc / float(k * k * C))
    out = torch.cat(planes, dim=1)
    assert out.shape == (B, d2, oh, ow)
    return out


def pixel_grid(batch: int, height: int, width: int, like: torch.Tensor) -> torch.Tensor:
    """(B,2,H,W) tensor of pixel coordinates, channel 0 = x, 1 = y (UnFlowLoss.py:11-20)."""
    xs = torch.arange(width, device=like.device, dtype=like.dtype).view(1, 1, 1, width).expand(batch, 1, height, width)
    ys = torch.arange(height, device=like.device, dtype=like.dtype).view(1, 1, height, 1).expand(batch, 1, height, width)
    return torch.cat([xs, ys], dim=1)


def flow_warp(image: torch.Tensor, flow12: torch.Tensor, mode: int = WARP_TORCH) -> torch.Tensor:
    """Bilinear border warp of ``image`` by ``flow12`` (pixels; channel 0 = x).

    mode WARP_TORCH: grid 2*(x+u)/(W-1)-1 fed to grid_sample(align_corners=False, border)
                     exactly as UnFlowLoss.py:83-94 does.
    mode WARP_TRT:   same grid, un-normalised as ((g+1)*(size-1))/2 like
                     trt_plugins/grid_sampler.cu:55-58, then clipped and sampled bilinearly.
    """
    B, _, H, W = image.shape
    pos = pixel_grid(B, H, W, image) + flow12
    gx = 2.0 * pos[:, 0] / (W - 1) - 1.0
    gy = 2.0 * pos[:, 1] / (H - 1) - 1.0
    if mode in (WARP_TORCH, WARP_TORCH_CPU):
        grid = torch.stack([gx, gy], dim=-1)
        return F.grid_sample(image, grid, mode="bilinear", padding_mode="border", align_corners=False)
    ix = ((gx + 1.0) * (W - 1)) / 2.0
    iy = ((gy + 1.0) * (H - 1)) / 2.0
    return _bilinear_border(image, ix, iy)


def _bilinear_border(image: torch.Tensor, ix: torch.Tensor, iy: torch.Tensor) -> torch.Tensor:
    """Sample ``image`` (B,C,H,W) at pixel positions (ix, iy) (B,H,W), clipping to the border;
    differentiable w.r.t. image and positions with the ATen clip rule (zero position gradient
    at or beyond the border)."""
    B, C, H, W = image.shape
    ix = ix.clamp(0, W - 1)  # clamp has zero gradient outside, like clip_coordinates_set_grad
    iy = iy.clamp(0, H - 1)
    x0 = ix.detach().floor()
    y0 = iy.detach().floor()
    x1, y1 = x0 + 1, y0 + 1
    w_nw = (x1 - ix) * (y1 - iy)
    w_ne = (ix - x0) * (y1 - iy)
    w_sw = (x1 - ix) * (iy - y0)
    w_se = (ix - x0) * (iy - y0)
    flat = image.reshape(B, C, H * W)

    def tap(xx, yy, w):
        ok = ((xx >= 0) & (xx < W) & (yy >= 0) & (yy < H)).to(image.dtype)
        idx = (yy.clamp(0, H - 1) * W + xx.clamp(0, W - 1)).long().reshape(B, 1, H * W).expand(B, C, H * W)
        vals = flat.gather(2, idx).reshape(B, C, H, W)
        return vals * (w * ok).unsqueeze(1)

    return tap(x0, y0, w_nw) + tap(x1, y0, w_ne) + tap(x0, y1, w_sw) + tap(x1, y1, w_se)


def level_forward(x1: torch.Tensor, x2: torch.Tensor, flow: Optional[torch.Tensor], pad_size: int = 4,
                  kernel_size: int = 1, max_displacement: int = 4, stride1: int = 1, stride2: int = 1,
                  warp_mode: int = WARP_TORCH, leaky_slope: Optional[float] = 0.1) -> torch.Tensor:
    """One decoder level's hot path: warp -> correlation -> LeakyReLU (pwcnet_sfd.py:171-182)."""
    second = x2 if flow is None else flow_warp(x2, flow, warp_mode).type(x1.dtype)
    out = correlation(x1, second, pad_size, kernel_size, max_displacement, stride1, stride2)
    if leaky_slope is not None:
        out = F.leaky_relu(out, leaky_slope)
    return out
