"""Grid sampler: the oracle is pinned on the CPU against torch.nn.functional.grid_sample (ATen is the
third-party arithmetic behind the reference's PyTorch path) for every mode; the CUDA kernel is checked against the
oracle for both un-normalise conventions -- ATen's and the TensorRT plugin's own
(runtime/cerberus_net/trt_plugins/grid_sampler.cu:48-59,181-233).

Tolerances: fp32 1e-5 of max|ref| (north_star); fp16 6e-4 (one rounding of the stored output).
"""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_err
from oracle import grid_sampler_oracle as go

MODES = {go.BILINEAR: "bilinear", go.NEAREST: "nearest"}
PADS = {go.ZEROS: "zeros", go.BORDER: "border", go.REFLECTION: "reflection"}


def make_case(seed, N=2, C=5, H=11, W=14, oH=9, oW=13, spread=0.9):
    rs = np.random.RandomState(seed)
    x = rs.standard_normal((N, C, H, W)).astype(np.float32)
    g = (rs.standard_normal((N, oH, oW, 2)) * spread).astype(np.float32)
    g[0, 0, 0] = (-1.0, -1.0); g[0, 0, 1] = (1.0, 1.0); g[0, 0, 2] = (0.0, 0.0); g[0, 0, 3] = (3.5, -2.25)   # corners, centre, far outside
    return x, g


@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("pad", list(PADS))
@pytest.mark.parametrize("align", [False, True])
def test_oracle_matches_aten_grid_sample_on_cpu(mode, pad, align):
    x, g = make_case(3)
    ref = F.grid_sample(torch.from_numpy(x), torch.from_numpy(g), mode=MODES[mode], padding_mode=PADS[pad], align_corners=align).numpy()
    assert rel_err(go.grid_sample(x, g, mode, pad, align, "aten"), ref) < 2e-6


def test_trt_convention_is_a_pixel_space_warp():
    """The plugin's un-normalise ((g+1)*(size-1))/2 (grid_sampler.cu:55-58) undoes the flow_warp grid
    2*(x+u)/(W-1) - 1 exactly: sampling position = clamp(x + u) -- SURVEY 8a-2, mode R."""
    rs = np.random.RandomState(1)
    H, W = 7, 9
    x = rs.standard_normal((1, 1, H, W)).astype(np.float32)
    ys, xs = np.meshgrid(np.arange(H, dtype=np.float32), np.arange(W, dtype=np.float32), indexing="ij")
    grid = np.stack([2 * xs / (W - 1) - 1, 2 * ys / (H - 1) - 1], -1)[None].astype(np.float32)
    out = go.grid_sample(x, grid, go.BILINEAR, go.BORDER, False, "trt")
    assert np.abs(out - x).max() < 1e-5           # identity at zero flow (ATen's convention is not: SURVEY 8a-2)
    assert np.abs(go.grid_sample(x, grid, go.BILINEAR, go.BORDER, False, "aten") - x).max() > 1e-2


@pytest.mark.gpu
@pytest.mark.parametrize("conv", ["aten", "trt"])
@pytest.mark.parametrize("mode", list(MODES))
@pytest.mark.parametrize("pad", list(PADS))
@pytest.mark.parametrize("align", [False, True])
def test_kernel_matches_oracle(conv, mode, pad, align):
    import cerberusnet_b200 as cb
    x, g = make_case(7, N=2, C=6, H=19, W=23, oH=17, oW=29, spread=0.8)
    dev = torch.device("cuda:0")
    out = cb.ops.grid_sample_forward(torch.from_numpy(x).to(dev), torch.from_numpy(g).to(dev), mode, pad, align, conv).cpu().numpy()
    ref = go.grid_sample(x, g, mode, pad, align, conv)
    if mode == go.NEAREST:
        # a coordinate within one fp32 ulp of a rounding boundary may pick the neighbour: allow a handful of pixels
        bad = np.abs(out - ref).max(axis=1) > 1e-5 * np.abs(ref).max()
        assert bad.mean() < 0.01
    else:
        assert rel_err(out, ref) < 1e-5


@pytest.mark.gpu
def test_kernel_matches_aten_cuda_and_half_precision():
    import cerberusnet_b200 as cb
    from cerberusnet_b200.flow_warp import grid_sample
    dev = torch.device("cuda:0")
    torch.manual_seed(2)
    x = torch.randn(3, 48, 40, 56, device=dev)
    g = torch.randn(3, 40, 56, 2, device=dev) * 0.7
    for pad in ("zeros", "border", "reflection"):
        ref = F.grid_sample(x, g, mode="bilinear", padding_mode=pad, align_corners=False)
        out = grid_sample(x, g, "bilinear", pad, False)
        assert rel_err(out.cpu().numpy(), ref.cpu().numpy()) < 1e-5
        out16 = grid_sample(x.half(), g.half(), "bilinear", pad, False)
        ref16 = go.grid_sample(x.half().float().cpu().numpy(), g.half().float().cpu().numpy(), go.BILINEAR, go.__dict__[pad.upper()], False, "aten")
        assert out16.dtype == torch.float16
        assert rel_err(out16.float().cpu().numpy(), ref16) < 6e-4     # one fp16 rounding of the output


@pytest.mark.gpu
def test_trt_shaped_grid_sampler_enqueue():
    import cerberusnet_b200 as cb
    from cerberusnet_b200 import _lib
    L = cb.lib()
    dev = torch.device("cuda:0")
    x, g = make_case(9, N=2, C=4, H=16, W=24, oH=16, oW=24)
    tx, tg = torch.from_numpy(x).to(dev), torch.from_numpy(g).to(dev)
    out = torch.empty(2, 4, 16, 24, device=dev)
    f = _lib.TrtGridSamplerFields()
    L.cerb_trt_grid_sampler_default_fields(ctypes.byref(f))

    def mk(dims):
        d = _lib.TrtTensorDesc()
        d.dims.nbDims = 4
        for i, v in enumerate(dims):
            d.dims.d[i] = v
        d.type, d.format, d.scale = 0, 0, 1.0
        return d

    ind = (_lib.TrtTensorDesc * 2)(mk((2, 4, 16, 24)), mk((2, 16, 24, 2)))
    outd = (_lib.TrtTensorDesc * 1)(mk((2, 4, 16, 24)))
    ins = (ctypes.c_void_p * 2)(tx.data_ptr(), tg.data_ptr())
    outs = (ctypes.c_void_p * 1)(out.data_ptr())
    assert L.cerb_trt_grid_sampler_enqueue(ctypes.byref(f), ind, outd, ins, outs, None,
                                           ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)) == 0
    torch.cuda.synchronize()
    assert rel_err(out.cpu().numpy(), go.grid_sample(x, g, go.BILINEAR, go.BORDER, False, "trt")) < 1e-5
    ind[1].type = 1   # grid kHALF with a kFLOAT input: the reference throws, we return an error code
    assert L.cerb_trt_grid_sampler_enqueue(ctypes.byref(f), ind, outd, ins, outs, None, None) < 0
