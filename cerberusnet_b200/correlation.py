"""Drop-in for the reference module ``nnet_training.correlation_package.correlation``.

Same names, constructor arguments, call signatures and error behaviour as the reference
(nnet_training/correlation_package/correlation.py:4-80), running on the sm_100a kernels behind
the C ABI.  Extra, opt-in surface: :class:`WarpCorrelation` / :func:`warp_correlation` fuse the
flow warp of the second map and the LeakyReLU that every flow decoder applies around the op
(nnet_models/pwcnet_sfd.py:171-182) into the same launch.
"""
from __future__ import annotations

from typing import Optional

import torch

from . import ops
from ._lib import WARP_TORCH, WARP_TORCH_CPU, WARP_TRT

__all__ = ["Correlation", "CorrelationFunction", "CorrelationTorch", "WarpCorrelation", "WarpCorrelationFunction",
           "warp_correlation", "WARP_TORCH", "WARP_TRT", "WARP_TORCH_CPU"]


def _register_ops():
    """torch.library registration under a namespace of our own, so the reference oracle's
    ``cerberus::`` ops (correlation_cuda.cpp:45-48) can be loaded next to it for A/B tests."""
    if hasattr(torch.ops, "cerberus_b200") and hasattr(torch.ops.cerberus_b200, "correlation"):
        return

    @torch.library.custom_op("cerberus_b200::correlation", mutates_args=())
    def correlation(input1: torch.Tensor, input2: torch.Tensor, pad_size: int, kernel_size: int,
                    max_displacement: int, stride1: int, stride2: int, corr_type_multiply: int) -> torch.Tensor:
        return ops.warp_corr_forward(input1, input2, None, pad_size, kernel_size, max_displacement, stride1, stride2,
                                     corr_type_multiply)

    @correlation.register_fake
    def _(input1, input2, pad_size, kernel_size, max_displacement, stride1, stride2, corr_type_multiply):
        import math
        kr = (kernel_size - 1) // 2
        border = kr + max_displacement
        d = 2 * (max_displacement // stride2) + 1
        oh = math.ceil((input1.shape[2] + 2 * pad_size - 2 * border) / stride1)
        ow = math.ceil((input1.shape[3] + 2 * pad_size - 2 * border) / stride1)
        return input1.new_empty((input1.shape[0], d * d, oh, ow))

    @torch.library.custom_op("cerberus_b200::correlation_backward", mutates_args=())
    def correlation_backward(input1: torch.Tensor, input2: torch.Tensor, gradOutput: torch.Tensor, pad_size: int,
                             kernel_size: int, max_displacement: int, stride1: int, stride2: int,
                             corr_type_multiply: int) -> list[torch.Tensor]:
        g1, g2, _ = ops.warp_corr_backward(input1, input2, None, None, gradOutput, pad_size, kernel_size,
                                           max_displacement, stride1, stride2, corr_type_multiply)
        return [g1, g2]

    @correlation_backward.register_fake
    def _(input1, input2, gradOutput, pad_size, kernel_size, max_displacement, stride1, stride2, corr_type_multiply):
        return [torch.empty_like(input1), torch.empty_like(input2)]


_register_ops()


class CorrelationFunction(torch.autograd.Function):
    """Mirror of the reference autograd Function (correlation.py:23-57): same ``apply``
    signature and defaults, 8-tuple backward."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, input1, input2, pad_size=3, kernel_size=3, max_displacement=20, stride1=1, stride2=2,
                corr_multiply=1):
        ctx.save_for_backward(input1, input2)
        ctx.pad_size = pad_size
        ctx.kernel_size = kernel_size
        ctx.max_displacement = max_displacement
        ctx.stride1 = stride1
        ctx.stride2 = stride2
        ctx.corr_multiply = corr_multiply
        return ops.warp_corr_forward(input1, input2, None, pad_size, kernel_size, max_displacement, stride1, stride2,
                                     corr_multiply)

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_outputs):
        input1, input2 = ctx.saved_tensors
        grad_input1, grad_input2, _ = ops.warp_corr_backward(
            input1, input2, None, None, grad_outputs, ctx.pad_size, ctx.kernel_size, ctx.max_displacement,
            ctx.stride1, ctx.stride2, ctx.corr_multiply)
        return grad_input1, grad_input2, None, None, None, None, None, None


class Correlation(torch.nn.Module):
    """Mirror of the reference ``Correlation`` module (correlation.py:60-80): no parameters or
    buffers (state_dicts stay interchangeable); autograd Function in training mode, raw op in
    eval mode."""

    def __init__(self, pad_size=0, kernel_size=0, max_displacement=0, stride1=1, stride2=2, corr_multiply=1):
        super().__init__()
        self.pad_size = pad_size
        self.kernel_size = kernel_size
        self.max_displacement = max_displacement
        self.stride1 = stride1
        self.stride2 = stride2
        self.corr_multiply = corr_multiply

    def forward(self, input1, input2):
        # the reference uses the autograd Function only in training mode (correlation.py:72-80), which silently cuts the
        # graph when a frozen-BN fine-tune runs the module in eval(): keep the graph whenever a gradient is wanted
        if self.training or (torch.is_grad_enabled() and (input1.requires_grad or input2.requires_grad)):
            return CorrelationFunction.apply(input1, input2, self.pad_size, self.kernel_size, self.max_displacement,
                                             self.stride1, self.stride2, self.corr_multiply)
        return ops.warp_corr_forward(input1, input2, None, self.pad_size, self.kernel_size, self.max_displacement,
                                     self.stride1, self.stride2, self.corr_multiply)

    def extra_repr(self):
        return (f"pad_size={self.pad_size}, kernel_size={self.kernel_size}, max_displacement={self.max_displacement}, "
                f"stride1={self.stride1}, stride2={self.stride2}, corr_multiply={self.corr_multiply}")


class CorrelationTorch(torch.nn.Module):
    """Name-compatible stand-in for the reference's pure-PyTorch variant (correlation.py:4-21):
    same constructor, same result (pad = max_displacement, kernel 1, strides 1), but computed by
    the CUDA op -- this package has no PyTorch-math path."""

    def __init__(self, max_displacement=4, *args, **kwargs):
        super().__init__()
        self.max_displacement = max_displacement
        self.output_dim = 2 * self.max_displacement + 1
        self.pad_size = self.max_displacement

    def forward(self, x1, x2):
        return CorrelationFunction.apply(x1, x2, self.pad_size, 1, self.max_displacement, 1, 1, 1)


class WarpCorrelationFunction(torch.autograd.Function):
    """out = leaky_relu(correlation(x1, flow_warp(x2, flow)), slope), one launch each way."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, input1, input2, flow, pad_size=4, kernel_size=1, max_displacement=4, stride1=1, stride2=1,
                corr_multiply=1, warp_mode=WARP_TORCH, leaky_slope=0.1):
        out = ops.warp_corr_forward(input1, input2, flow, pad_size, kernel_size, max_displacement, stride1, stride2,
                                    corr_multiply, warp_mode, leaky_slope)
        ctx.cfg = (pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply, warp_mode, leaky_slope)
        ctx.has_flow = flow is not None
        if flow is not None:
            ctx.save_for_backward(input1, input2, flow, out)
        else:
            ctx.save_for_backward(input1, input2, out)
        return out

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_out):
        if ctx.has_flow:
            input1, input2, flow, out = ctx.saved_tensors
        else:
            input1, input2, out = ctx.saved_tensors
            flow = None
        pad, k, md, s1, s2, mult, mode, slope = ctx.cfg
        g1, g2, gflow = ops.warp_corr_backward(input1, input2, flow, out if slope is not None else None, grad_out, pad,
                                               k, md, s1, s2, mult, mode, slope)
        if gflow is not None and flow is not None and gflow.dtype != flow.dtype:
            gflow = gflow.to(flow.dtype)
        return g1, g2, gflow, None, None, None, None, None, None, None, None


def warp_correlation(input1: torch.Tensor, input2: torch.Tensor, flow: Optional[torch.Tensor] = None, pad_size: int = 4,
                     kernel_size: int = 1, max_displacement: int = 4, stride1: int = 1, stride2: int = 1,
                     corr_multiply: int = 1, warp_mode: int = WARP_TORCH,
                     leaky_slope: Optional[float] = 0.1) -> torch.Tensor:
    """Functional form of :class:`WarpCorrelation`."""
    needs_grad = torch.is_grad_enabled() and any(
        t is not None and t.requires_grad for t in (input1, input2, flow))
    if needs_grad:
        return WarpCorrelationFunction.apply(input1, input2, flow, pad_size, kernel_size, max_displacement, stride1,
                                             stride2, corr_multiply, warp_mode, leaky_slope)
    return ops.warp_corr_forward(input1, input2, flow, pad_size, kernel_size, max_displacement, stride1, stride2,
                                 corr_multiply, warp_mode, leaky_slope)


class UpflowWarpCorrelationFunction(torch.autograd.Function):
    """(out, flow_up) with flow_up = interpolate(2 * flow_coarse, x2, bilinear, align_corners=True) and
    out = leaky_relu(correlation(x1, flow_warp(x2, flow_up)), slope): the decoder's four steps
    (pwcnet_sfd.py:176-182) in one launch.  Backward: the fused backward gives the gradient with respect
    to flow_up; the up-sampling's adjoint (ATen) carries it, plus whatever arrives on flow_up, to flow_coarse."""

    @staticmethod
    @torch.amp.custom_fwd(device_type="cuda")
    def forward(ctx, input1, input2, flow_coarse, pad_size=4, kernel_size=1, max_displacement=4, stride1=1, stride2=1,
                corr_multiply=1, warp_mode=WARP_TORCH, leaky_slope=0.1):
        out, flow_up = ops.warp_corr_forward_upflow(input1, input2, flow_coarse, pad_size, kernel_size, max_displacement,
                                                    stride1, stride2, corr_multiply, warp_mode, leaky_slope)
        ctx.cfg = (pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply, warp_mode, leaky_slope)
        ctx.coarse_shape = tuple(flow_coarse.shape)
        ctx.coarse_dtype = flow_coarse.dtype
        ctx.save_for_backward(input1, input2, flow_up, out)
        return out, flow_up

    @staticmethod
    @torch.amp.custom_bwd(device_type="cuda")
    def backward(ctx, grad_out, grad_flow_up):
        input1, input2, flow_up, out = ctx.saved_tensors
        pad, k, md, s1, s2, mult, mode, slope = ctx.cfg
        g1, g2, gflow = ops.warp_corr_backward(input1, input2, flow_up, out if slope is not None else None,
                                               grad_out.contiguous(), pad, k, md, s1, s2, mult, mode, slope)
        if grad_flow_up is not None:
            gflow = gflow + grad_flow_up.float()
        H, W = flow_up.shape[-2:]
        gcoarse = torch.ops.aten.upsample_bilinear2d_backward(gflow.contiguous(), [H, W], list(ctx.coarse_shape), True,
                                                              2.0, 2.0) * 2
        return g1, g2, gcoarse.to(ctx.coarse_dtype), None, None, None, None, None, None, None, None


def warp_correlation_upflow(input1: torch.Tensor, input2: torch.Tensor, flow_coarse: torch.Tensor, pad_size: int = 4,
                            kernel_size: int = 1, max_displacement: int = 4, stride1: int = 1, stride2: int = 1,
                            corr_multiply: int = 1, warp_mode: int = WARP_TORCH, leaky_slope: Optional[float] = 0.1,
                            out: Optional[torch.Tensor] = None, flow_up: Optional[torch.Tensor] = None):
    """``(cost_volume, flow_up)`` from the next-coarser level's flow.  ``out`` / ``flow_up``
    (e.g. slices of the decoder's concat buffer) are honoured on the no-grad path."""
    needs_grad = torch.is_grad_enabled() and any(t.requires_grad for t in (input1, input2, flow_coarse))
    if needs_grad:
        return UpflowWarpCorrelationFunction.apply(input1, input2, flow_coarse, pad_size, kernel_size, max_displacement,
                                                   stride1, stride2, corr_multiply, warp_mode, leaky_slope)
    return ops.warp_corr_forward_upflow(input1, input2, flow_coarse, pad_size, kernel_size, max_displacement, stride1,
                                        stride2, corr_multiply, warp_mode, leaky_slope, out=out, flow_up=flow_up)


class WarpCorrelation(Correlation):
    """``Correlation`` with the decoder's surrounding ops fused in.

    ``forward(input1, input2)`` is exactly the reference module.  ``forward(input1, input2, flow)``
    additionally warps ``input2`` by ``flow`` (``flow_warp`` semantics, UnFlowLoss.py:83-94) and
    applies ``LeakyReLU(leaky_slope)`` -- the three steps of pwcnet_sfd.py:178-182 in one kernel.
    """

    def __init__(self, pad_size=4, kernel_size=1, max_displacement=4, stride1=1, stride2=1, corr_multiply=1,
                 warp_mode=WARP_TORCH, leaky_slope: Optional[float] = 0.1):
        super().__init__(pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply)
        self.warp_mode = warp_mode
        self.leaky_slope = leaky_slope

    def forward(self, input1, input2, flow=None, fuse_activation: bool = False):
        if flow is None and not fuse_activation:
            return super().forward(input1, input2)
        return warp_correlation(input1, input2, flow, self.pad_size, self.kernel_size, self.max_displacement,
                                self.stride1, self.stride2, self.corr_multiply, self.warp_mode, self.leaky_slope)
