"""Functional layer over the C ABI: tensors in, tensors out, current CUDA stream.

These are thin: argument normalisation, output allocation, one C call.  No math happens in
Python and nothing here falls back to PyTorch ops.
"""
from __future__ import annotations

import ctypes
import functools
from typing import Optional, Tuple

import torch

from . import _lib
from ._lib import (VARIANT_AUTO, WARP_TORCH, WARP_TORCH_CPU, WARP_TRT, CostVolumeError, check, current_stream_ptr,
                   device_guard, lib, make_params_cached as make_params, output_dims, ptr, require_cuda)

__all__ = ["warp_corr_forward", "warp_corr_forward_upflow", "warp_corr_backward", "flow_warp_forward", "flow_warp_backward", "corr_output_shape",
           "grid_sample_forward",
           "WARP_TORCH", "WARP_TRT", "WARP_TORCH_CPU"]


def _inner_contig(t: torch.Tensor) -> torch.Tensor:
    """The ABI needs W-stride 1 (N/C/H strides are free)."""
    if t.dim() != 4:
        raise CostVolumeError(f"expected a 4-D NCHW tensor, got shape {tuple(t.shape)}")
    return t if (t.stride(3) == 1 and min(t.stride()) >= 0) else t.contiguous()


def corr_output_shape(x_shape, pad_size, kernel_size, max_displacement, stride1, stride2) -> Tuple[int, int, int, int]:
    """(B, D*D, outH, outW) by the reference rule (correlation_cuda.cpp:6-14)."""
    return _corr_output_shape(tuple(x_shape), int(pad_size), int(kernel_size), int(max_displacement), int(stride1),
                              int(stride2))


@functools.lru_cache(maxsize=1024)
def _corr_output_shape(x_shape, pad_size, kernel_size, max_displacement, stride1, stride2):
    B, C, H, W = x_shape
    p = _lib.CorrParams()
    p.batch, p.channels, p.height, p.width = B, C, H, W
    p.pad_size, p.kernel_size, p.max_displacement, p.stride1, p.stride2 = (
        int(pad_size), int(kernel_size), int(max_displacement), int(stride1), int(stride2))
    oc, oh, ow = output_dims(p)
    return B, oc, oh, ow


def warp_corr_forward(x1: torch.Tensor, x2: torch.Tensor, flow: Optional[torch.Tensor] = None, pad_size: int = 4,
                      kernel_size: int = 1, max_displacement: int = 4, stride1: int = 1, stride2: int = 1,
                      corr_multiply: int = 1, warp_mode: int = WARP_TORCH, leaky_slope: Optional[float] = None,
                      out: Optional[torch.Tensor] = None, variant: int = VARIANT_AUTO, x2_roll: int = 0) -> torch.Tensor:
    """leaky_relu(correlation(x1, flow_warp(x2, flow))) in one kernel launch.

    ``x2_roll``: batch item n of x1 / flow / out is paired with item ``(n + x2_roll) % B`` of x2.  With
    ``x1 = x2 = features of cat([image1, image2])`` and ``x2_roll = B // 2`` one launch computes both flow
    directions of the reference's ``consistency=True`` forward (pwcnet.py:108-113) without copying a feature map.

    ``flow=None`` skips the warp, ``leaky_slope=None`` skips the activation; with both off this is
    the reference's ``torch.ops.cerberus.correlation`` (correlation_cuda.cpp:3-26).
    ``out`` may be a channel slice of a wider buffer (e.g. the decoder's concat tensor).
    """
    require_cuda(x1, x2, flow, out)
    if x1.shape != x2.shape:
        raise CostVolumeError(f"input shapes differ: {tuple(x1.shape)} vs {tuple(x2.shape)}")
    if x1.dtype != x2.dtype:
        raise CostVolumeError(f"input dtypes differ: {x1.dtype} vs {x2.dtype}")
    x1, x2 = _inner_contig(x1), _inner_contig(x2)
    if flow is not None:
        if flow.shape != (x1.shape[0], 2, x1.shape[2], x1.shape[3]):
            raise CostVolumeError(f"flow must be (B,2,H,W), got {tuple(flow.shape)}")
        flow = _inner_contig(flow.float())
    shape = corr_output_shape(x1.shape, pad_size, kernel_size, max_displacement, stride1, stride2)
    if out is None:
        out = torch.empty(shape, dtype=x1.dtype, device=x1.device)
    elif tuple(out.shape) != shape or out.dtype != x1.dtype or out.stride(3) != 1:
        raise CostVolumeError(f"out must be {shape} {x1.dtype} with unit W stride")
    p = make_params(x1, x2, flow, out, pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply,
                    warp_mode, leaky_slope, int(x2_roll))
    with device_guard(x1.device):
        rc = lib().cerb_warp_corr_forward_variant(ctypes.byref(p), ptr(x1), ptr(x2), ptr(flow), ptr(out), int(variant),
                                                  ctypes.c_void_p(current_stream_ptr(x1.device)))
    check(rc, "cerb_warp_corr_forward")
    return out


def warp_corr_forward_upflow(x1: torch.Tensor, x2: torch.Tensor, flow_coarse: torch.Tensor, pad_size: int = 4,
                             kernel_size: int = 1, max_displacement: int = 4, stride1: int = 1, stride2: int = 1,
                             corr_multiply: int = 1, warp_mode: int = WARP_TORCH, leaky_slope: Optional[float] = None,
                             out: Optional[torch.Tensor] = None, flow_up: Optional[torch.Tensor] = None):
    """The decoder's ``flow = F.interpolate(flow_coarse * 2, scale_factor=2, mode='bilinear',
    align_corners=True)`` (pwcnet_sfd.py:176) fused into :func:`warp_corr_forward`.

    ``flow_coarse`` is (B,2,H/2,W/2); returns ``(out, flow_up)`` where ``flow_up`` (B,2,H,W) is the
    up-sampled flow the decoder goes on to use -- pass a channel slice of the concat buffer to
    have it written in place.  Needs pad_size == max_displacement >= 4, kernel_size 1, strides 1.
    """
    require_cuda(x1, x2, flow_coarse, out, flow_up)
    if x1.shape != x2.shape or x1.dtype != x2.dtype:
        raise CostVolumeError("inputs must have the same shape and dtype")
    B, C, H, W = x1.shape
    if flow_coarse.shape != (B, 2, H // 2, W // 2) or H % 2 or W % 2:
        raise CostVolumeError(f"flow_coarse must be (B,2,H/2,W/2) with even H, W; got {tuple(flow_coarse.shape)}")
    x1, x2 = _inner_contig(x1), _inner_contig(x2)
    flow_coarse = _inner_contig(flow_coarse.float())
    shape = corr_output_shape(x1.shape, pad_size, kernel_size, max_displacement, stride1, stride2)
    if out is None:
        out = torch.empty(shape, dtype=x1.dtype, device=x1.device)
    elif tuple(out.shape) != shape or out.dtype != x1.dtype or out.stride(3) != 1:
        raise CostVolumeError(f"out must be {shape} {x1.dtype} with unit W stride")
    if flow_up is None:
        flow_up = torch.empty(B, 2, H, W, dtype=torch.float32, device=x1.device)
    elif tuple(flow_up.shape) != (B, 2, H, W) or flow_up.dtype != torch.float32 or flow_up.stride(3) != 1:
        raise CostVolumeError("flow_up must be (B,2,H,W) float32 with unit W stride")
    p = make_params(x1, x2, None, out, pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply,
                    warp_mode, leaky_slope)
    cs = (ctypes.c_int64 * 4)(*flow_coarse.stride())
    us = (ctypes.c_int64 * 4)(*flow_up.stride())
    with device_guard(x1.device):
        rc = lib().cerb_warp_corr_forward_upflow(ctypes.byref(p), ptr(x1), ptr(x2), ptr(flow_coarse), cs, ptr(flow_up), us,
                                                 ptr(out), ctypes.c_void_p(current_stream_ptr(x1.device)))
    check(rc, "cerb_warp_corr_forward_upflow")
    return out, flow_up


def warp_corr_backward(x1: torch.Tensor, x2: torch.Tensor, flow: Optional[torch.Tensor], out: Optional[torch.Tensor],
                       grad_out: torch.Tensor, pad_size: int = 4, kernel_size: int = 1, max_displacement: int = 4,
                       stride1: int = 1, stride2: int = 1, corr_multiply: int = 1, warp_mode: int = WARP_TORCH,
                       leaky_slope: Optional[float] = None, x2_roll: int = 0):
    """Gradients of :func:`warp_corr_forward`: ``(grad_x1, grad_x2, grad_flow or None)``.  With ``x2_roll`` the
    second gradient is laid out like x2 itself (item ``(n + x2_roll) % B`` receives what item n's output sent back)."""
    require_cuda(x1, x2, flow, out, grad_out)
    x1, x2 = _inner_contig(x1), _inner_contig(x2)
    if flow is not None:
        flow = _inner_contig(flow.float())
    grad_out = grad_out.contiguous()
    if out is not None:
        out = out.contiguous()
    if leaky_slope is not None and out is None:
        raise CostVolumeError("the activated forward output is needed for the LeakyReLU backward")
    p = make_params(x1, x2, flow, grad_out, pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply,
                    warp_mode, leaky_slope, int(x2_roll))
    g1 = torch.empty(x1.shape, dtype=x1.dtype, device=x1.device)
    g2 = torch.empty(x2.shape, dtype=x2.dtype, device=x2.device)
    gflow = torch.empty(flow.shape, dtype=torch.float32, device=x1.device) if flow is not None else None
    need = lib().cerb_warp_corr_backward_workspace(ctypes.byref(p), 1 if flow is not None else 0)
    ws = torch.empty(need, dtype=torch.uint8, device=x1.device) if need else None
    with device_guard(x1.device):
        rc = lib().cerb_warp_corr_backward(ctypes.byref(p), ptr(x1), ptr(x2), ptr(flow), ptr(out), ptr(grad_out),
                                           ptr(g1), ptr(g2), ptr(gflow), ptr(ws), need,
                                           ctypes.c_void_p(current_stream_ptr(x1.device)))
    check(rc, "cerb_warp_corr_backward")
    return g1, g2, gflow


def flow_warp_forward(image: torch.Tensor, flow: torch.Tensor, warp_mode: int = WARP_TORCH) -> torch.Tensor:
    require_cuda(image, flow)
    image = image.contiguous()
    flow = flow.float().contiguous()
    B, C, H, W = image.shape
    if flow.shape != (B, 2, H, W):
        raise CostVolumeError(f"flow must be (B,2,H,W), got {tuple(flow.shape)}")
    out = torch.empty_like(image)
    with device_guard(image.device):
        rc = lib().cerb_flow_warp_forward(ptr(image), ptr(flow), ptr(out), B, C, H, W, _lib.dtype_code(image),
                                          int(warp_mode), ctypes.c_void_p(current_stream_ptr(image.device)))
    check(rc, "cerb_flow_warp_forward")
    return out


def flow_warp_backward(image: torch.Tensor, flow: torch.Tensor, grad_out: torch.Tensor, warp_mode: int = WARP_TORCH):
    require_cuda(image, flow, grad_out)
    image = image.contiguous()
    flow = flow.float().contiguous()
    grad_out = grad_out.contiguous()
    B, C, H, W = image.shape
    gimg = torch.empty_like(image)
    gflow = torch.empty_like(flow)
    with device_guard(image.device):
        rc = lib().cerb_flow_warp_backward(ptr(image), ptr(flow), ptr(grad_out), ptr(gimg), ptr(gflow), B, C, H, W,
                                           _lib.dtype_code(image), int(warp_mode),
                                           ctypes.c_void_p(current_stream_ptr(image.device)))
    check(rc, "cerb_flow_warp_backward")
    return gimg, gflow


_GRID_MODES = {"bilinear": 0, "nearest": 1}
_GRID_PADS = {"zeros": 0, "border": 1, "reflection": 2}
_GRID_CONVS = {"trt": 0, "aten": 1}


def grid_sample_forward(input: torch.Tensor, grid: torch.Tensor, mode="bilinear", padding_mode="border",
                        align_corners: bool = False, convention: str = "trt") -> torch.Tensor:
    """Grid sampler with a normalised-coordinate grid (N,oH,oW,2), every mode of the reference's TensorRT plugin
    (trt_plugins/grid_sampler.cu:146-271).  ``mode`` / ``padding_mode`` take PyTorch's strings or the plugin's
    integers (grid_sampler.hpp:14-15).  ``convention='trt'`` un-normalises like the plugin
    (``((g+1)*(size-1))/2`` for align_corners=False, ``roundf`` for nearest), ``'aten'`` like ``F.grid_sample``."""
    require_cuda(input, grid)
    if input.dim() != 4 or grid.dim() != 4 or grid.shape[0] != input.shape[0] or grid.shape[3] != 2:
        raise CostVolumeError(f"expected input (N,C,H,W) and grid (N,oH,oW,2), got {tuple(input.shape)} / {tuple(grid.shape)}")
    input = input.contiguous()
    grid = grid.to(input.dtype).contiguous()
    m = _GRID_MODES[mode] if isinstance(mode, str) else int(mode)
    pm = _GRID_PADS[padding_mode] if isinstance(padding_mode, str) else int(padding_mode)
    N, C, H, W = input.shape
    oH, oW = grid.shape[1:3]
    out = torch.empty(N, C, oH, oW, dtype=input.dtype, device=input.device)
    with device_guard(input.device):
        rc = lib().cerb_grid_sample_forward(ptr(input), ptr(grid), ptr(out), N, C, H, W, oH, oW, _lib.dtype_code(input), m, pm,
                                            1 if align_corners else 0, _GRID_CONVS[convention],
                                            ctypes.c_void_p(current_stream_ptr(input.device)))
    check(rc, "cerb_grid_sample_forward")
    return out
