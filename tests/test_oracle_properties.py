"""Properties of the oracle that no golden vector covers: general parameters (the two independent
restatements must agree), adjointness of the backward, reference-order accumulation, output-size
rule (correlation_cuda.cpp:6-14) and the pad != max_displacement case of the reference's own
__main__ block (correlation.py:87).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import rel_err
from oracle import c_oracle as co
from oracle import torch_oracle as to

PARAMS = [(4, 1, 4, 1, 1), (3, 3, 4, 2, 2), (2, 1, 4, 1, 2), (6, 3, 4, 1, 1), (5, 1, 4, 2, 1), (8, 1, 8, 1, 1),
          (0, 1, 2, 1, 1), (20, 3, 20, 1, 2)]


def _inputs(seed, B=2, C=10, H=14, W=22, sigma=3.0):
    rs = np.random.RandomState(seed)
    return (rs.standard_normal((B, C, H, W)).astype(np.float32), rs.standard_normal((B, C, H, W)).astype(np.float32),
            (rs.standard_normal((B, 2, H, W)) * sigma).astype(np.float32))


@pytest.mark.parametrize("p,k,md,s1,s2", PARAMS)
def test_c_and_torch_oracles_agree(p, k, md, s1, s2):
    x1, x2, _ = _inputs(1)
    a = co.corr_forward(x1, x2, p, k, md, s1, s2)
    b = to.correlation(torch.from_numpy(x1), torch.from_numpy(x2), p, k, md, s1, s2).numpy()
    assert a.shape == b.shape == (2,) + co.corr_out_dims(14, 22, p, k, md, s1, s2)
    assert rel_err(a, b) < 2e-6
    assert rel_err(co.corr_forward(x1, x2, p, k, md, s1, s2, co.ACC_REFERENCE_ORDER), a) < 2e-6


@pytest.mark.parametrize("p,k,md,s1,s2", PARAMS[:6])
def test_backward_is_the_adjoint(p, k, md, s1, s2):
    """<corr(x1,x2) , g> differentiated: the backward must satisfy the dot-product test and agree
    with autograd of the independent PyTorch restatement."""
    x1, x2, _ = _inputs(2)
    t1 = torch.from_numpy(x1).requires_grad_()
    t2 = torch.from_numpy(x2).requires_grad_()
    y = to.correlation(t1, t2, p, k, md, s1, s2)
    g = np.random.RandomState(3).standard_normal(tuple(y.shape)).astype(np.float32)
    y.backward(torch.from_numpy(g))
    g1, g2 = co.corr_backward(x1, x2, g, p, k, md, s1, s2)
    assert rel_err(g1, t1.grad.numpy()) < 2e-6
    assert rel_err(g2, t2.grad.numpy()) < 2e-6
    if k == 1 and s1 == 1:
        r1, r2 = co.corr_backward_reforder(x1, x2, g, p, md, s2)
        assert rel_err(r1, g1) < 2e-6 and rel_err(r2, g2) < 2e-6
    # bilinear form: <g, corr(x1, x2)> == <g1, x1> == <g2, x2>
    lhs = float((co.corr_forward(x1, x2, p, k, md, s1, s2).astype(np.float64) * g).sum())
    assert abs(lhs - float((g1.astype(np.float64) * x1).sum())) < 1e-4 * max(1.0, abs(lhs))
    assert abs(lhs - float((g2.astype(np.float64) * x2).sum())) < 1e-4 * max(1.0, abs(lhs))


def test_reference_main_block_shape():
    """correlation.py:85-87: pad 4, md 10 -> (H-12) x (W-12), 441 channels."""
    assert co.corr_out_dims(128, 64, 4, 1, 10, 1, 1) == (441, 116, 52)
    x1, x2, _ = _inputs(4, B=1, C=5, H=16, W=20)
    out = co.corr_forward(x1, x2, 4, 1, 10, 1, 1)
    assert out.shape == (1, 441, 4, 8)
    # centre displacement (dy=dx=0) is the plain channel mean of products on the cropped window
    centre = out[:, 220]
    ref = (x1 * x2).mean(1)[:, 6:10, 6:14]
    assert rel_err(centre, ref) < 2e-6


def test_empty_output_is_an_error():
    with pytest.raises(ValueError):
        co.corr_out_dims(12, 20, 4, 1, 10, 1, 1)


@pytest.mark.parametrize("mode", [co.WARP_TORCH_CPU, co.WARP_TRT])
def test_warp_oracles_agree_and_clip(mode):
    img, _, flow = _inputs(5, C=6, sigma=9.0)  # most samples leave the image
    a = co.flow_warp_forward(img, flow, mode)
    b = to.flow_warp(torch.from_numpy(img), torch.from_numpy(flow), mode).numpy()
    assert rel_err(a, b) < 2e-6
    assert a.min() >= img.min() - 1e-6 and a.max() <= img.max() + 1e-6  # convex combination of pixels
    ti = torch.from_numpy(img).requires_grad_()
    tf = torch.from_numpy(flow).requires_grad_()
    g = np.random.RandomState(6).standard_normal(img.shape).astype(np.float32)
    to.flow_warp(ti, tf, mode).backward(torch.from_numpy(g))
    gi, gf = co.flow_warp_backward(img, flow, g, mode)
    assert rel_err(gi, ti.grad.numpy()) < 2e-6
    assert rel_err(gf, tf.grad.numpy()) < 1e-5
    # a sample clipped at the border has zero flow gradient (ATen clip_coordinates_set_grad)
    far = flow.copy()
    far[:, 0] = 1000.0
    _, gfar = co.flow_warp_backward(img, far, g, mode)
    assert np.all(gfar[:, 0] == 0)


def test_trt_warp_is_pixel_exact():
    """Mode TRT samples at clamp(x+u, 0, W-1): an integer flow is a pure shift with border clamp."""
    img, _, _ = _inputs(7, B=1, C=3, H=9, W=12)
    flow = np.zeros((1, 2, 9, 12), np.float32)
    flow[:, 0] = 2.0
    flow[:, 1] = -1.0
    out = co.flow_warp_forward(img, flow, co.WARP_TRT)
    ys = np.clip(np.arange(9) - 1, 0, 8)
    xs = np.clip(np.arange(12) + 2, 0, 11)
    assert rel_err(out, img[:, :, ys][:, :, :, xs]) < 1e-6


def test_leaky_relu():
    x = np.array([-2.0, -0.0, 0.0, 3.0], np.float32)
    np.testing.assert_allclose(co.leaky_relu(x, 0.1), [-0.2, 0.0, 0.0, 3.0], rtol=1e-7)
    y = co.leaky_relu(x, 0.1)
    np.testing.assert_allclose(co.leaky_relu_backward(y, np.ones(4, np.float32), 0.1), [0.1, 0.1, 0.1, 1.0], rtol=1e-7)


def test_flow_upsample_restatement_matches_aten():
    """the C restatement of `F.interpolate(2*flow, scale_factor=2, 'bilinear', align_corners=True)`
    (pwcnet_sfd.py:176) against ATen's CPU kernel, and its composition with the level oracle against the
    un-fused composition -- the oracle the GPU test of cerb_warp_corr_forward_upflow leans on."""
    rs = np.random.RandomState(3)
    for (B, Hc, Wc) in ((1, 2, 2), (2, 5, 9), (1, 8, 16), (3, 16, 12)):
        coarse = rs.standard_normal((B, 2, Hc, Wc)).astype(np.float32) * 2.0
        want = torch.nn.functional.interpolate(torch.from_numpy(coarse) * 2, scale_factor=2, mode="bilinear",
                                               align_corners=True).numpy()
        got = co.flow_upsample2x(coarse)
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= 2e-6 * max(1.0, np.abs(want).max())
    # exactness where it must be exact: a constant field stays constant (x2), corners are the coarse corners (x2)
    const = np.full((1, 2, 4, 6), 0.75, np.float32)
    assert np.all(co.flow_upsample2x(const) == 1.5)
    up = co.flow_upsample2x(coarse)
    assert np.array_equal(up[:, :, 0, 0], 2 * coarse[:, :, 0, 0]) and np.array_equal(up[:, :, -1, -1], 2 * coarse[:, :, -1, -1])
    # composed level: warp by the up-sampled flow
    x1 = rs.standard_normal((1, 6, 16, 24)).astype(np.float32)
    x2 = rs.standard_normal((1, 6, 16, 24)).astype(np.float32)
    cf = rs.standard_normal((1, 2, 8, 12)).astype(np.float32)
    a = co.level_forward(x1, x2, co.flow_upsample2x(cf))
    b = to.level_forward(torch.from_numpy(x1), torch.from_numpy(x2),
                         torch.nn.functional.interpolate(torch.from_numpy(cf) * 2, scale_factor=2, mode="bilinear",
                                                         align_corners=True), warp_mode=to.WARP_TORCH_CPU).numpy()
    a_cpu = co.level_forward(x1, x2, co.flow_upsample2x(cf), warp_mode=co.WARP_TORCH_CPU)
    assert rel_err(a_cpu, b) < 1e-5 and a.shape == b.shape
