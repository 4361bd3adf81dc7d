"""ONNX symbolics for the hot path -- the exporter side of the TensorRT plugins.

The reference registers two custom symbolics before ``torch.onnx.export`` (nnet_training/utilities/onnx_export.py:18-28):
``cerberus::correlation`` (attributes ``pad_size_i ... corr_multiply_i``, what CorrelationPlugin parses) and
``torch::grid_sampler`` (``interpolation_mode_i, padding_mode_i, align_corners_i``, GridSamplerPlugin).  This module
provides the same two under the same names and attribute names, plus the fused node

    cerberus::warp_correlation(im1, im2, flow)   pad_size_i, kernel_size_i, max_displacement_i, stride1_i, stride2_i,
                                                 corr_multiply_i, warp_mode_i, leaky_slope_f

which WarpCorrelationPlugin (csrc/trt_plugin_shim.cpp, type "warp_correlation") consumes: per pyramid level it replaces
the 2x ScatterND + Transpose + grid_sampler + correlation + LeakyRelu nodes of the reference's export.
In the exported graph the warp follows the TensorRT plugin's convention (CERB_WARP_TRT) unless told otherwise, because
that is what the reference's runtime computes for this sub-graph.

    from cerberusnet_b200 import onnx_export
    onnx_export.register()          # before torch.onnx.export(..., dynamo=False)
"""
from __future__ import annotations

import torch

from ._lib import WARP_TRT

OPSET = 11   # onnx_export.py:45,57


def correlation_op(g, input1, input2, pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply):
    """onnx_export.py:18-23"""
    return g.op("cerberus::correlation", input1, input2, pad_size_i=pad_size, kernel_size_i=kernel_size,
                max_displacement_i=max_displacement, stride1_i=stride1, stride2_i=stride2, corr_multiply_i=corr_multiply)


def grid_sample_op(g, input1, input2, mode, padding_mode, align_corners):
    """onnx_export.py:25-28"""
    return g.op("torch::grid_sampler", input1, input2, interpolation_mode_i=mode, padding_mode_i=padding_mode,
                align_corners_i=int(align_corners))


def warp_correlation_op(g, input1, input2, flow, pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply,
                        warp_mode, leaky_slope):
    return g.op("cerberus::warp_correlation", input1, input2, flow, pad_size_i=pad_size, kernel_size_i=kernel_size,
                max_displacement_i=max_displacement, stride1_i=stride1, stride2_i=stride2, corr_multiply_i=corr_multiply,
                warp_mode_i=warp_mode, leaky_slope_f=float(leaky_slope))


class _ExportCorrelation(torch.autograd.Function):
    """Carries the symbolic for the plain op when a model calls ``Correlation`` in eval mode."""

    @staticmethod
    def forward(ctx, input1, input2, pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply):
        from . import ops
        return ops.warp_corr_forward(input1, input2, None, pad_size, kernel_size, max_displacement, stride1, stride2,
                                     corr_multiply)

    @staticmethod
    def symbolic(g, input1, input2, pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply):
        return correlation_op(g, input1, input2, pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply)


class _ExportWarpCorrelation(torch.autograd.Function):
    @staticmethod
    def forward(ctx, input1, input2, flow, pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply,
                warp_mode, leaky_slope):
        from . import ops
        return ops.warp_corr_forward(input1, input2, flow, pad_size, kernel_size, max_displacement, stride1, stride2,
                                     corr_multiply, warp_mode, leaky_slope)

    @staticmethod
    def symbolic(g, input1, input2, flow, pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply,
                 warp_mode, leaky_slope):
        return warp_correlation_op(g, input1, input2, flow, pad_size, kernel_size, max_displacement, stride1, stride2,
                                   corr_multiply, warp_mode, leaky_slope)


class ExportableWarpCorrelation(torch.nn.Module):
    """Drop this in place of the decoder's warp -> correlation -> LeakyReLU triple before exporting: traces to ONE
    ``cerberus::warp_correlation`` node (or ``cerberus::correlation`` when called without a flow)."""

    def __init__(self, pad_size=4, kernel_size=1, max_displacement=4, stride1=1, stride2=1, corr_multiply=1,
                 warp_mode=WARP_TRT, leaky_slope=0.1):
        super().__init__()
        self.cfg = (pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply)
        self.warp_mode, self.leaky_slope = warp_mode, leaky_slope

    def forward(self, input1, input2, flow=None):
        if flow is None:
            return _ExportCorrelation.apply(input1, input2, *self.cfg)
        return _ExportWarpCorrelation.apply(input1, input2, flow, *self.cfg, self.warp_mode, self.leaky_slope)


def register(opset: int = OPSET) -> None:
    """What the reference's export_model does before torch.onnx.export (onnx_export.py:44-45), for the legacy
    (TorchScript) exporter: the raw ``cerberus::correlation`` op and ATen's grid_sampler get the plugin node names."""
    from torch.onnx.symbolic_helper import parse_args
    torch.onnx.register_custom_op_symbolic(
        "cerberus::correlation", parse_args("v", "v", "i", "i", "i", "i", "i", "i")(correlation_op), opset)
    torch.onnx.register_custom_op_symbolic(
        "cerberus_b200::correlation", parse_args("v", "v", "i", "i", "i", "i", "i", "i")(correlation_op), opset)
    torch.onnx.register_custom_op_symbolic("::grid_sampler", parse_args("v", "v", "i", "i", "b")(grid_sample_op), opset)
