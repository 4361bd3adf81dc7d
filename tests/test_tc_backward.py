"""Tensor-core backward (costvolume_bwd_tc.cu: both correlation gradients as banded GEMMs in TMEM, 3xTF32) followed by the
shared-memory-window splat (costvolume_splat.cu: fixed-point reductions, one TMA reduce per 8 channels) against the C
oracle and the CUDA-core kernels, through the C ABI.

Reference sites: correlation_backward_input1/2 (correlation_cuda_kernel.cu:97-242), LeakyReLU backward of
pwcnet_sfd.py:182, grid_sample backward of UnFlowLoss.py:83-94.  Tolerance: 1e-5 of max|ref| (north_star, fp32); the
fixed-point window adds at most 2^-23 of the tile's largest contribution per term (measured: ~1e-6 in total, the same as
the CUDA-core kernel's summation-order noise).

The kernel choice is pinned with cerb_debug_set_backward_kernel (1 = tensor cores wherever supported, 0 = CUDA cores).
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


def dev():
    return torch.device("cuda:0")


@pytest.fixture
def tc_backward():
    L = cb.lib()
    L.cerb_debug_set_backward_kernel(1)
    yield L
    L.cerb_debug_set_backward_kernel(-1)


def case(seed, B, C, H, W, sigma):
    rs = np.random.RandomState(seed)
    x1 = rs.standard_normal((B, C, H, W)).astype(np.float32)
    x2 = rs.standard_normal((B, C, H, W)).astype(np.float32)
    fl = (rs.standard_normal((B, 2, H, W)) * sigma).astype(np.float32) if sigma is not None else None
    go = rs.standard_normal((B, 81, H, W)).astype(np.float32)
    return x1, x2, fl, go


def run_backward(x1, x2, fl, go, slope, mode=cb.WARP_TORCH, roll=0):
    t1, t2 = torch.from_numpy(x1).to(dev()), torch.from_numpy(x2).to(dev())
    tf = torch.from_numpy(fl).to(dev()) if fl is not None else None
    tg = torch.from_numpy(go).to(dev())
    # the saved activation comes from the oracle's forward: an output within rounding of zero may have the other sign in
    # a kernel that sums in a different order, and one flipped mask bit moves a gradient element by (1 - slope)|g x| / C
    x2o = np.roll(x2, -roll, 0) if roll else x2
    out = torch.from_numpy(co.level_forward(x1, x2o, fl, 4, 1, 4, 1, 1, mode, slope)).to(dev())
    return ops.warp_corr_backward(t1, t2, tf, out, tg, 4, 1, 4, 1, 1, 1, mode, slope, x2_roll=roll)


@pytest.mark.parametrize("B,C,H,W,sigma", [
    (2, 16, 24, 64, 2.0),      # 24 tiles, windows fit
    (1, 48, 40, 96, 1.5),      # HRNet level-3 channel count, 30 tiles: three partial accumulators per gradient
    (3, 20, 17, 36, 2.0),      # ragged: tiles cut by the image border, C % 16 != 0
    (2, 32, 16, 48, 14.0),     # flow far beyond the splat window: scattered red.global path
    (4, 32, 32, 64, None),     # no flow: both gradients stored
    (1, 64, 16, 32, 1.5),      # 64 channels: two partial accumulators
    (1, 96, 16, 32, 1.5),      # 96 channels: one accumulator per gradient
    (1, 128, 8, 16, None),     # N = 128: two band slots
    (1, 8, 5, 12, 1.0),        # smaller than one tile
])
@pytest.mark.parametrize("mode", [cb.WARP_TORCH, cb.WARP_TRT])
def test_tc_backward_vs_oracle(tc_backward, B, C, H, W, sigma, mode):
    x1, x2, fl, go = case(B * 1000 + C + W, B, C, H, W, sigma)
    r1, r2, rf = co.level_backward(x1, x2, fl, go, 4, 1, 4, 1, 1, mode, 0.1)
    n0 = tc_backward.cerb_launch_count()
    g1, g2, gf = run_backward(x1, x2, fl, go, 0.1, mode)
    # the tensor-core path is one kernel (+ warp forward, memset and the splat kernel with a flow)
    assert tc_backward.cerb_launch_count() - n0 == (4 if fl is not None else 1)
    assert rel_err(g1.cpu().numpy(), r1) < TOL
    assert rel_err(g2.cpu().numpy(), r2) < TOL
    if fl is not None:
        assert rel_err(gf.cpu().numpy(), rf) < TOL


def test_tc_backward_no_activation_and_roll(tc_backward):
    """x2_batch_roll (both flow directions from one feature tensor): the second gradient comes back in x2's own order."""
    x1, _, fl, go = case(7, 4, 24, 24, 48, 1.5)
    r1, r2, rf = co.level_backward(x1, np.roll(x1, -2, 0), fl, go, 4, 1, 4, 1, 1, co.WARP_TORCH, None)
    f = torch.from_numpy(x1).to(dev())
    g1, g2, gf = ops.warp_corr_backward(f, f, torch.from_numpy(fl).to(dev()), None, torch.from_numpy(go).to(dev()),
                                        4, 1, 4, 1, 1, 1, cb.WARP_TORCH, None, x2_roll=2)
    assert rel_err(g1.cpu().numpy(), r1) < TOL
    assert rel_err(torch.roll(g2, -2, 0).cpu().numpy(), r2) < TOL
    assert rel_err(gf.cpu().numpy(), rf) < TOL


def test_tc_backward_window_and_direct_paths_are_counted(tc_backward):
    """A smooth flow keeps every tile of the splat on the shared-memory window path; a wild one takes scattered atomics."""
    import ctypes
    x1, x2, fl, go = case(11, 2, 32, 32, 64, 1.0)
    ctr = torch.zeros(4, dtype=torch.int64, device=dev())
    tc_backward.cerb_debug_set_path_counters(ctypes.c_void_p(ctr.data_ptr()))
    try:
        run_backward(x1, x2, fl, go, 0.1)
        torch.cuda.synchronize()
        smooth = ctr.cpu().tolist()
        ctr.zero_()
        run_backward(x1, x2, (fl * 20).astype(np.float32), go, 0.1)
        torch.cuda.synchronize()
        wild = ctr.cpu().tolist()
    finally:
        tc_backward.cerb_debug_set_path_counters(None)
    tiles = 2 * 4 * 4
    assert smooth[1] == tiles and smooth[2] == 0
    assert wild[2] > 0 and wild[1] + wild[2] == tiles


@pytest.mark.parametrize("C,H,W,B", [(48, 128, 256, 8), (32, 128, 256, 2)])
def test_tc_backward_full_size_matches_cuda_cores(C, H, W, B):
    """HRNet training level at batch 8 (BASELINE configs[3]) and the finest PWC level, both flow directions: tensor-core
    and CUDA-core backward agree to the parity bar; per-plane checksums agree (size-independent property)."""
    L = cb.lib()
    g = torch.Generator(device=dev()).manual_seed(9)
    f1 = torch.nn.functional.leaky_relu(torch.randn(B, C, H, W, device=dev(), generator=g), 0.1)
    f2 = torch.nn.functional.leaky_relu(torch.randn(B, C, H, W, device=dev(), generator=g), 0.1)
    fl = (torch.randn(B, 2, H, W, device=dev(), generator=g) * 1.5).clamp_(-6, 6)
    go = torch.randn(B, 81, H, W, device=dev(), generator=g)
    out = ops.warp_corr_forward(f1, f2, fl, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1)
    res = {}
    try:
        for mode in (0, 1):
            L.cerb_debug_set_backward_kernel(mode)
            res[mode] = ops.warp_corr_backward(f1, f2, fl, out, go, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1)
    finally:
        L.cerb_debug_set_backward_kernel(-1)
    for a, b in zip(res[0], res[1]):
        scale = float(a.abs().max())
        assert float((a - b).abs().max()) < TOL * scale
        assert float((a.double().sum((2, 3)) - b.double().sum((2, 3))).abs().max()) < TOL * scale * H * W
