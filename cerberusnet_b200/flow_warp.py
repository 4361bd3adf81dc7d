"""Drop-in for ``flow_warp`` of the reference's ``nnet_training.loss_functions.UnFlowLoss``.

``flow_warp(image, flow12, pad='border', mode='bilinear')`` keeps the reference signature
(UnFlowLoss.py:83-94) but runs one CUDA kernel: no CPU mesh grid + H2D copy, no normalised grid
tensor, no ``grid_sample`` call.  Differentiable w.r.t. both arguments.
"""
from __future__ import annotations

import torch

from . import ops
from ._lib import WARP_TORCH, WARP_TORCH_CPU, WARP_TRT

__all__ = ["flow_warp", "FlowWarpFunction", "mesh_grid", "norm_grid", "grid_sample"]


class FlowWarpFunction(torch.autograd.Function):
    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, image, flow12, warp_mode=WARP_TORCH):
        ctx.save_for_backward(image, flow12)
        ctx.warp_mode = warp_mode
        return ops.flow_warp_forward(image, flow12, warp_mode)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_out):
        image, flow12 = ctx.saved_tensors
        gimg, gflow = ops.flow_warp_backward(image, flow12, grad_out, ctx.warp_mode)
        return gimg, gflow.to(flow12.dtype), None


def flow_warp(image: torch.Tensor, flow12: torch.Tensor, pad: str = "border", mode: str = "bilinear",
              warp_mode: int = WARP_TORCH) -> torch.Tensor:
    """Warps ``image`` (B,C,H,W) by ``flow12`` (B,2,H,W; channel 0 = x, pixels).

    Like the reference, the result has the dtype of ``image`` when both are fp32; under fp16
    autocast the reference returns fp32 and its callers cast back (pwcnet_sfd.py:178) -- here the
    result already has ``image.dtype``.
    """
    if pad != "border" or mode != "bilinear":
        raise NotImplementedError(
            f"flow_warp(pad={pad!r}, mode={mode!r}): only the decoder hot-path case "
            "(pad='border', mode='bilinear') has a native kernel")
    if torch.is_grad_enabled() and (image.requires_grad or flow12.requires_grad):
        return FlowWarpFunction.apply(image, flow12, warp_mode)
    return ops.flow_warp_forward(image, flow12, warp_mode)


def mesh_grid(batch_sz: int, height: int, width: int, device=None, dtype=torch.float32) -> torch.Tensor:
    """(B,2,H,W) pixel coordinates, channel 0 = x (UnFlowLoss.py:11-20).  Kept for API parity;
    the fused kernels never materialise it.  Built on ``device`` (the reference builds it on the
    CPU and copies it across on every call)."""
    xs = torch.arange(width, device=device, dtype=dtype).view(1, 1, 1, width).expand(batch_sz, 1, height, width)
    ys = torch.arange(height, device=device, dtype=dtype).view(1, 1, height, 1).expand(batch_sz, 1, height, width)
    return torch.cat([xs, ys], dim=1)


def norm_grid(v_grid: torch.Tensor) -> torch.Tensor:
    """Scale a (B,2,H,W) pixel grid to [-1,1] and return it as (B,H,W,2) (UnFlowLoss.py:22-32)."""
    _, _, height, width = v_grid.size()
    gx = 2.0 * v_grid[:, 0] / (width - 1) - 1.0
    gy = 2.0 * v_grid[:, 1] / (height - 1) - 1.0
    return torch.stack([gx, gy], dim=-1)


def grid_sample(input: torch.Tensor, grid: torch.Tensor, mode: str = "bilinear", padding_mode: str = "zeros",
                align_corners: bool = False, convention: str = "aten") -> torch.Tensor:
    """``F.grid_sample`` signature on the CUDA kernel behind the reference's TensorRT grid-sampler plugin
    (trt_plugins/grid_sampler.cu); ``convention='trt'`` reproduces the plugin's own un-normalise.  Forward only
    (the plugin is an inference node); training code warps through :func:`flow_warp`."""
    return ops.grid_sample_forward(input, grid, mode, padding_mode, align_corners, convention)
