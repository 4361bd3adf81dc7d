"""PWC-style flow decoder built on the fused op -- the caller side of the hot path.

The per-level loop is the one every flow decoder of the reference repeats
(nnet_models/pwcnet.py:62-101, pwcnet_sfd.py:163-203/284-328, ocrnet_sfd.py:175-214,
detr_sfd.py:227-266): up-sample the flow x2 (values x2) -> warp the second feature map -> correlation
-> LeakyReLU(0.1) -> concat with a 1x1 projection of the first map and the flow -> flow estimator ->
context network.  Here the three hot-path ops are one launch (`warp_correlation`), and in eval
mode the activated cost volume is written straight into the concat buffer.

This module is a harness for measuring the path inside a training / inference step (the `train` block of
bench.py, tools/train_bench.py); the estimator and context stacks are plain torch convolutions, sized like the
reference's "lite" estimator, not a re-implementation of the reference models.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from ._lib import WARP_TORCH
from .correlation import warp_correlation, warp_correlation_upflow


def _conv(cin: int, cout: int, k: int = 3, stride: int = 1, dilation: int = 1, act: bool = True) -> nn.Sequential:
    layers: List[nn.Module] = [nn.Conv2d(cin, cout, k, stride, ((k - 1) * dilation) // 2, dilation)]
    if act:
        layers.append(nn.LeakyReLU(0.1, inplace=True))
    return nn.Sequential(*layers)


class PyramidEncoder(nn.Module):
    """Strided conv feature pyramid, coarsest level first (shape contract of the reference's
    FeatureExtractor: channels [16,32,64,96,128,192] at strides 2..64)."""

    def __init__(self, channels: Sequence[int] = (3, 16, 32, 64, 96, 128, 192)):
        super().__init__()
        self.stages = nn.ModuleList(
            nn.Sequential(_conv(a, b, stride=2), _conv(b, b)) for a, b in zip(channels[:-1], channels[1:]))

    def forward(self, x: torch.Tensor) -> List[torch.Tensor]:
        feats = []
        for stage in self.stages:
            x = stage(x)
            feats.append(x)
        return feats[::-1]


class FlowDecoder(nn.Module):
    """Coarse-to-fine flow decoder over two feature pyramids (coarsest first)."""

    def __init__(self, channels_in: Sequence[int], max_displacement: int = 4, proj_channels: int = 32,
                 output_level: int = 4, warp_mode: int = WARP_TORCH, leaky_slope: float = 0.1,
                 fuse_upsample: bool = True):
        super().__init__()
        self.fuse_upsample = fuse_upsample
        self.md, self.warp_mode, self.slope, self.output_level = max_displacement, warp_mode, leaky_slope, output_level
        self.n_corr = (2 * max_displacement + 1) ** 2
        # only the levels the loop visits get a projection: DDP requires every parameter to take part
        self.proj = nn.ModuleList(_conv(c, proj_channels, k=1) for c in list(channels_in)[:output_level + 1])
        cin = self.n_corr + proj_channels + 2
        self.est1, self.est2 = _conv(cin, 128), _conv(128, 128)
        self.est3, self.est4, self.est5 = _conv(256, 96), _conv(224, 64), _conv(160, 32)
        self.est_flow = _conv(96, 2, act=False)
        self.context = nn.Sequential(_conv(34, 128), _conv(128, 128, dilation=2), _conv(128, 128, dilation=4),
                                     _conv(128, 96, dilation=8), _conv(96, 64, dilation=16), _conv(64, 32),
                                     _conv(32, 2, act=False))

    def _estimate(self, x):
        a = self.est1(x)
        b = self.est2(a)
        c = self.est3(torch.cat([a, b], 1))
        d = self.est4(torch.cat([b, c], 1))
        e = self.est5(torch.cat([c, d], 1))
        return e, self.est_flow(torch.cat([d, e], 1))

    def forward(self, pyr1: List[torch.Tensor], pyr2: List[torch.Tensor]) -> List[torch.Tensor]:
        flows = []
        B, _, h, w = pyr1[0].shape
        flow = pyr1[0].new_zeros(B, 2, h, w)
        for level, (f1, f2) in enumerate(zip(pyr1, pyr2)):
            H, W = f1.shape[2:]
            # flow = interpolate(2 * flow, x2, bilinear, align_corners=True) rides along in the op's prologue
            #  whenever this level is exactly twice the previous one
            fuse_up = level > 0 and self.fuse_upsample and H == 2 * flow.shape[2] and W == 2 * flow.shape[3]
            if level > 0 and not fuse_up:
                flow = F.interpolate(flow * 2, size=(H, W), mode="bilinear", align_corners=True)
            proj = self.proj[level](f1)
            warp_flow = flow if level > 0 else None
            if torch.is_grad_enabled():
                if fuse_up:
                    cost, flow = warp_correlation_upflow(f1, f2, flow, self.md, 1, self.md, 1, 1, 1, self.warp_mode,
                                                         self.slope)
                else:
                    cost = warp_correlation(f1, f2, warp_flow, self.md, 1, self.md, 1, 1, 1, self.warp_mode, self.slope)
                x = torch.cat([cost, proj, flow], 1)
            else:  # inference: the activated cost volume (and the up-sampled flow) land in the concat buffer directly
                x = f1.new_empty(B, self.n_corr + proj.shape[1] + 2, H, W)
                if fuse_up:
                    ops.warp_corr_forward_upflow(f1, f2, flow, self.md, 1, self.md, 1, 1, 1, self.warp_mode, self.slope,
                                                 out=x[:, :self.n_corr], flow_up=x[:, -2:])
                    flow = x[:, -2:]
                else:
                    ops.warp_corr_forward(f1, f2, warp_flow, self.md, 1, self.md, 1, 1, 1, self.warp_mode, self.slope,
                                          out=x[:, :self.n_corr])
                    x[:, -2:] = flow
                x[:, self.n_corr:self.n_corr + proj.shape[1]] = proj
            feat, dflow = self._estimate(x)
            flow = flow + dflow
            flow = flow + self.context(torch.cat([feat, flow], 1))
            flows.append(flow)
            if level == self.output_level:
                break
        return flows[::-1]


class FlowNetLite(nn.Module):
    """Encoder + decoder, optional backward flow (the reference's `consistency` pass,
    nnet_models/pwcnet.py:103-113)."""

    def __init__(self, channels: Sequence[int] = (3, 16, 32, 64, 96, 128, 192), **decoder_kwargs):
        super().__init__()
        self.encoder = PyramidEncoder(channels)
        self.decoder = FlowDecoder(list(channels[1:])[::-1], **decoder_kwargs)

    def forward(self, img1: torch.Tensor, img2: torch.Tensor, consistency: bool = False):
        p1, p2 = self.encoder(img1), self.encoder(img2)
        out = {"flow": self.decoder(p1, p2)}
        if consistency:
            out["flow_b"] = self.decoder(p2, p1)
        return out


def photometric_loss(img1: torch.Tensor, img2: torch.Tensor, flows: List[torch.Tensor], warp_mode: int = WARP_TORCH):
    """Unsupervised L1 photometric + first-order smoothness loss over the flow pyramid (a compact
    stand-in for unFlowLoss, loss_functions/UnFlowLoss.py:255-322, using the CUDA `flow_warp`)."""
    from .flow_warp import flow_warp
    total = img1.new_zeros(())
    for i, fl in enumerate(flows):
        scale = img1.shape[-1] // fl.shape[-1]
        a = F.avg_pool2d(img1, scale) if scale > 1 else img1
        b = F.avg_pool2d(img2, scale) if scale > 1 else img2
        rec = flow_warp(b, fl, warp_mode=warp_mode)
        smooth = (fl[..., 1:] - fl[..., :-1]).abs().mean() + (fl[..., 1:, :] - fl[..., :-1, :]).abs().mean()
        total = total + (0.5 ** i) * ((rec - a).abs().mean() + 0.1 * smooth)
    return total
