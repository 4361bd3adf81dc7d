"""Tensor-core forward (variant 7: tcgen05 / TMEM, 3xTF32 split of the fp32 operands) against the C oracle and the
CUDA-core kernels, through the C ABI.  Tolerance: 1e-5 of max|ref| (north_star, fp32) -- the split keeps ~22 mantissa
bits per product (measured 4e-7 ... 2e-6).

Covers what the generic sweeps in test_gpu_parity.py do not: several tiles per persistent CTA, C up to 384 (48 K steps
through the 4-slot operand ring), C not a multiple of the 8-channel K step, pad != max_displacement, flows far beyond the
raw box (direct-gather fallback), widths without TMA alignment, strided output, x2_batch_roll, and the full bench size.
"""
import numpy as np
import pytest
import torch

import cerberusnet_b200 as cb
from cerberusnet_b200 import ops
from conftest import rel_err
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu
TOL = 1e-5
TC = 7


def dev():
    return torch.device("cuda:0")


def case(seed, B, C, H, W, sigma):
    rs = np.random.RandomState(seed)
    x1 = rs.standard_normal((B, C, H, W)).astype(np.float32)
    x2 = rs.standard_normal((B, C, H, W)).astype(np.float32)
    fl = (rs.standard_normal((B, 2, H, W)) * sigma).astype(np.float32) if sigma is not None else None
    return x1, x2, fl


@pytest.mark.parametrize("B,C,H,W,pad,sigma", [
    (4, 48, 64, 160, 4, 1.5),     # 320 tiles: up to three per CTA; C = 48 (HRNet level 3)
    (1, 384, 16, 32, 4, 1.5),     # 48 K steps (HRNet level 0)
    (2, 13, 21, 45, 4, 2.5),      # ragged everything: W % 4 != 0 (no TMA at all), C % 8 != 0
    (2, 36, 40, 96, 4, 12.0),     # flow far beyond the raw box: direct gather
    (1, 24, 30, 52, 6, 1.5),      # pad > md: output larger than the input
    (1, 24, 30, 52, 2, 1.5),      # pad < md: output smaller, tile origin not 16-byte aligned
    (2, 32, 24, 64, 4, None),     # no flow
    (1, 8, 5, 9, 4, 1.0),         # smaller than one tile
])
@pytest.mark.parametrize("mode", [cb.WARP_TORCH, cb.WARP_TRT])
def test_tc_forward_vs_oracle(B, C, H, W, pad, sigma, mode):
    x1, x2, fl = case(B * 1000 + C, B, C, H, W, sigma)
    ref = co.level_forward(x1, x2, fl, pad, 1, 4, 1, 1, mode, 0.1)
    t1, t2 = torch.from_numpy(x1).to(dev()), torch.from_numpy(x2).to(dev())
    tf = torch.from_numpy(fl).to(dev()) if fl is not None else None
    out = ops.warp_corr_forward(t1, t2, tf, pad, 1, 4, 1, 1, 1, mode, 0.1, variant=TC)
    assert rel_err(out.cpu().numpy(), ref) < TOL
    # channel slice of a wider buffer (the decoder's concat tensor), nothing written outside the slice
    buf = torch.full((B, ref.shape[1] + 34, ref.shape[2], ref.shape[3]), 7.0, device=dev())
    ops.warp_corr_forward(t1, t2, tf, pad, 1, 4, 1, 1, 1, mode, 0.1, out=buf[:, :ref.shape[1]], variant=TC)
    assert rel_err(buf[:, :ref.shape[1]].cpu().numpy(), ref) < TOL
    assert bool((buf[:, ref.shape[1]:] == 7.0).all())


@pytest.mark.parametrize("B,C,H,W,pad,sigma", [
    (2, 32, 48, 160, 8, 1.5),     # KITTI-shaped level (BASELINE configs[4]): 60 tiles x 2 accumulator passes per image
    (1, 20, 19, 37, 8, 2.5),      # ragged, no TMA
    (1, 8, 32, 48, 6, 1.5),       # pad < md
    (2, 40, 16, 32, 8, None),     # no flow, C % 8 != 0 handled by zero-filled K step
    (1, 16, 24, 64, 8, 15.0),     # direct gather
])
def test_tc_forward_md8_vs_oracle(B, C, H, W, pad, sigma):
    """max_displacement 8 (289 planes): the 24 x 32 halo is two accumulator passes of 12 x 32 positions."""
    x1, x2, fl = case(B * 100 + C, B, C, H, W, sigma)
    ref = co.level_forward(x1, x2, fl, pad, 1, 8, 1, 1, co.WARP_TORCH, 0.1)
    t1, t2 = torch.from_numpy(x1).to(dev()), torch.from_numpy(x2).to(dev())
    tf = torch.from_numpy(fl).to(dev()) if fl is not None else None
    out = ops.warp_corr_forward(t1, t2, tf, pad, 1, 8, 1, 1, 1, cb.WARP_TORCH, 0.1, variant=TC)
    assert out.shape[1] == 289
    assert rel_err(out.cpu().numpy(), ref) < TOL


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 6e-4), (torch.bfloat16, 5e-3)])
@pytest.mark.parametrize("B,C,H,W,pad,md,sigma", [
    (2, 32, 24, 64, 4, 4, 1.5),
    (1, 64, 40, 96, 4, 4, 1.5),     # 64 channels = one full 128-byte operand row (4 K steps of 16)
    (2, 20, 19, 37, 4, 4, 2.5),     # ragged: no TMA, partial K step
    (1, 48, 32, 64, 8, 8, 1.5),     # max_displacement 8
    (1, 24, 30, 52, 2, 4, 12.0),    # pad < md, direct gather
    (1, 96, 16, 48, 4, 4, None),    # no flow
])
def test_tc_forward_16bit_vs_oracle(dtype, tol, B, C, H, W, pad, md, sigma):
    """fp16 / bf16 inputs on the tensor cores (kind::f16; the blended x2w as hi + lo in T, two products): the only error
    against an exact evaluation on the same 16-bit inputs is the rounding of the stored result (2^-11 / 2^-8 of max|ref|)."""
    x1, x2, fl = case(B * 10 + C + md, B, C, H, W, sigma)
    t1 = torch.from_numpy(x1).to(dev()).to(dtype)
    t2 = torch.from_numpy(x2).to(dev()).to(dtype)
    tf = torch.from_numpy(fl).to(dev()) if fl is not None else None
    ref = co.level_forward(t1.float().cpu().numpy(), t2.float().cpu().numpy(), fl, pad, 1, md, 1, 1, co.WARP_TORCH, 0.1)
    out = ops.warp_corr_forward(t1, t2, tf, pad, 1, md, 1, 1, 1, cb.WARP_TORCH, 0.1, variant=TC)
    assert out.dtype == dtype
    assert rel_err(out.float().cpu().numpy(), ref) < tol


def test_tc_forward_no_activation_and_roll():
    x1, _, fl = case(5, 4, 40, 32, 80, 2.0)
    f = torch.from_numpy(x1).to(dev())
    tf = torch.from_numpy(fl).to(dev())
    both = ops.warp_corr_forward(f, f, tf, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, None, variant=TC, x2_roll=2)
    ref = co.level_forward(x1, np.roll(x1, -2, 0), fl, 4, 1, 4, 1, 1, co.WARP_TORCH, None)
    assert rel_err(both.cpu().numpy(), ref) < TOL


def test_tc_is_what_auto_runs_at_bench_size_and_matches_cuda_cores():
    """Finest PWC level, both flow directions per launch (BASELINE configs[1]): AUTO == variant 7 bit for bit, and both
    agree with the CUDA-core kernel (variant 1) to the parity bar; checksum-of-planes property at full size."""
    g = torch.Generator(device=dev()).manual_seed(3)
    f = torch.nn.functional.leaky_relu(torch.randn(2, 32, 128, 256, device=dev(), generator=g), 0.1)
    fl = (torch.randn(2, 2, 128, 256, device=dev(), generator=g) * 1.5).clamp_(-6, 6)
    auto = ops.warp_corr_forward(f, f, fl, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1, x2_roll=1)
    tc = ops.warp_corr_forward(f, f, fl, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1, x2_roll=1, variant=TC)
    cc = ops.warp_corr_forward(f, f, fl, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1, x2_roll=1, variant=1)
    assert torch.equal(auto, tc)
    scale = float(cc.abs().max())
    assert float((tc - cc).abs().max()) < TOL * scale
    assert float((tc.double().sum((2, 3)) - cc.double().sum((2, 3))).abs().max()) < TOL * scale * 128 * 256
