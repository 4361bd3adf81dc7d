"""Photometric term of the reference's ``unFlowLoss`` on one fused CUDA kernel pair.

    loss_photometric(im_orig, flow_warp(im_src, flow), ones)        UnFlowLoss.py:225-241, 279-283, 299

becomes ``photometric_loss(im_orig, im_src, flow, l1_weight, ssim_weight)``: flow-warp, L1, SSIM (ReflectionPad2d(1) +
3x3 average pools, loss_functions.py:47-78) and the mean in one launch, the gradient with respect to the flow in another.
``patch_unflow_loss()`` (install.py) redirects the reference loss module's ``flow_warp`` name to the CUDA warp; this
function is the further, fused step for callers that adopt it.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch

from . import _lib
from ._lib import WARP_TORCH, CostVolumeError, check, current_stream_ptr, device_guard, lib, ptr, require_cuda

__all__ = ["photometric_loss", "PhotometricLossFunction"]


def _check(im_orig, im_src, flow):
    require_cuda(im_orig, im_src, flow)
    if im_orig.shape != im_src.shape or im_orig.dim() != 4:
        raise CostVolumeError(f"images must share one (N,C,H,W) shape, got {tuple(im_orig.shape)} / {tuple(im_src.shape)}")
    N, C, H, W = im_orig.shape
    if tuple(flow.shape) != (N, 2, H, W):
        raise CostVolumeError(f"flow must be (N,2,H,W), got {tuple(flow.shape)}")
    return N, C, H, W


class PhotometricLossFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, im_orig, im_src, flow, l1_weight, ssim_weight, warp_mode):
        N, C, H, W = _check(im_orig, im_src, flow)
        a, b, f = im_orig.float().contiguous(), im_src.float().contiguous(), flow.float().contiguous()
        L = lib()
        ws_bytes = L.cerb_photometric_workspace(N, H, W)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=a.device)
        loss = torch.empty((), dtype=torch.float32, device=a.device)
        with device_guard(a.device):
            rc = L.cerb_photometric_forward(ptr(a), ptr(b), ptr(f), ptr(loss), ptr(ws), ws_bytes, N, C, H, W, float(l1_weight),
                                            float(ssim_weight), int(warp_mode), ctypes.c_void_p(current_stream_ptr(a.device)))
        check(rc, "cerb_photometric_forward")
        ctx.save_for_backward(a, b, f)
        ctx.cfg = (float(l1_weight), float(ssim_weight), int(warp_mode), flow.dtype)
        return loss

    @staticmethod
    def backward(ctx, grad_loss):
        a, b, f = ctx.saved_tensors
        l1_w, ssim_w, mode, fdtype = ctx.cfg
        N, C, H, W = a.shape
        g = grad_loss.float().contiguous()
        gflow = torch.empty_like(f)
        with device_guard(a.device):
            rc = lib().cerb_photometric_backward(ptr(a), ptr(b), ptr(f), ptr(g), ptr(gflow), N, C, H, W, l1_w, ssim_w, mode,
                                                 ctypes.c_void_p(current_stream_ptr(a.device)))
        check(rc, "cerb_photometric_backward")
        return None, None, gflow.to(fdtype), None, None, None


def photometric_loss(im_orig: torch.Tensor, im_src: torch.Tensor, flow: torch.Tensor, l1_weight: float = 0.15,
                     ssim_weight: float = 0.85, warp_mode: int = WARP_TORCH) -> torch.Tensor:
    """``l1_weight * mean|im_orig - rec| + ssim_weight * mean SSIM(rec, im_orig)`` with ``rec = flow_warp(im_src, flow)``
    -- the value of ``unFlowLoss.loss_photometric(im_orig, rec, ones)`` (weights: MonoSF_gauss.json l1 .15 / ssim .85).
    Differentiable with respect to ``flow``."""
    return PhotometricLossFunction.apply(im_orig, im_src, flow, l1_weight, ssim_weight, warp_mode)
