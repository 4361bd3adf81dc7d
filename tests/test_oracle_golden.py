"""The oracle against the golden vectors generated from the reference's own Python
(tests/golden/make_golden.py: CorrelationTorch correlation.py:4-21, flow_warp UnFlowLoss.py:83-94,
leaky_relu(0.1) pwcnet_sfd.py:182, autograd for the gradients).  CPU only."""
import os

import numpy as np
import pytest
import torch

from conftest import rel_err
from make_golden import CASES, golden_inputs
from oracle import c_oracle as co
from oracle import torch_oracle as to

TOL = 2e-6  # fp32 summation-order noise only; the oracles accumulate in float64 / ATen order


def _load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name + ".npz"))


@pytest.mark.parametrize("name", [n for n in CASES if n.startswith("corr_")])
def test_correlation_forward_matches_reference(golden_dir, name):
    x1, x2, _ = golden_inputs(name)
    gold = _load(golden_dir, name)
    for acc in (co.ACC_DOUBLE, co.ACC_REFERENCE_ORDER):
        out = co.corr_forward(x1, x2, 4, 1, 4, 1, 1, acc)
        if name == "corr_config1":
            assert rel_err(out[:, :, ::4, ::4], gold["out_sub4"]) < TOL
            sums = out.astype(np.float64).sum(axis=(2, 3))
            assert rel_err(sums, gold["plane_sums_f64"]) < 1e-5
        else:
            assert out.shape == gold["out"].shape
            assert rel_err(out, gold["out"]) < TOL
    if name != "corr_config1":
        t = to.correlation(torch.from_numpy(x1), torch.from_numpy(x2), 4, 1, 4, 1, 1).numpy()
        assert rel_err(t, gold["out"]) < TOL


@pytest.mark.parametrize("name", [n for n in CASES if n.startswith("warp_")])
def test_flow_warp_matches_reference(golden_dir, name):
    img, _, flow = golden_inputs(name)
    gold = _load(golden_dir, name)
    out = co.flow_warp_forward(img, flow, co.WARP_TORCH_CPU)  # fixtures come from ATen's CPU kernels
    assert rel_err(out, gold["out"]) < TOL
    # the CUDA flavour (grid / (size-1) as a reciprocal multiply) is 1 ulp apart in the grid
    assert rel_err(co.flow_warp_forward(img, flow, co.WARP_TORCH), gold["out"]) < 5e-5
    g = np.random.RandomState(1000 + CASES[name][0]).standard_normal(out.shape).astype(np.float32)
    gimg, gflow = co.flow_warp_backward(img, flow, g, co.WARP_TORCH_CPU)
    assert rel_err(gimg, gold["grad_image"]) < TOL
    assert rel_err(gflow, gold["grad_flow"]) < 1e-5
    t = to.flow_warp(torch.from_numpy(img), torch.from_numpy(flow), to.WARP_TORCH).numpy()
    assert rel_err(t, gold["out"]) < TOL


def test_flow_warp_is_not_identity_at_zero_flow(golden_dir):
    """the training-path warp samples at (x*W/(W-1) - 0.5), so zero flow does
    not return the image; the TensorRT convention does."""
    img, _, flow = golden_inputs("warp_zero_flow")
    assert np.all(flow == 0)
    gold = _load(golden_dir, "warp_zero_flow")["out"]
    assert np.abs(gold - img).max() > 0.1
    assert rel_err(co.flow_warp_forward(img, flow, co.WARP_TORCH_CPU), gold) < TOL
    # TRT convention: position x+u up to the fp32 rounding of the normalise/un-normalise round trip
    assert rel_err(co.flow_warp_forward(img, flow, co.WARP_TRT), img) < 2e-6


@pytest.mark.parametrize("name", [n for n in CASES if n.startswith("level_")])
def test_decoder_level_matches_reference(golden_dir, name):
    x1, x2, flow = golden_inputs(name)
    gold = _load(golden_dir, name)
    out = co.level_forward(x1, x2, flow, 4, 1, 4, 1, 1, co.WARP_TORCH_CPU, 0.1)
    assert rel_err(out, gold["out"]) < TOL
    assert rel_err(co.level_forward(x1, x2, flow, 4, 1, 4, 1, 1, co.WARP_TORCH, 0.1), gold["out"]) < 5e-5
    g = np.random.RandomState(1000 + CASES[name][0]).standard_normal(out.shape).astype(np.float32)
    g1, g2, gf = co.level_backward(x1, x2, flow, g, 4, 1, 4, 1, 1, co.WARP_TORCH_CPU, 0.1)
    assert rel_err(g1, gold["grad_x1"]) < TOL
    assert rel_err(g2, gold["grad_x2"]) < TOL
    assert rel_err(gf, gold["grad_flow"]) < 1e-5
    t = to.level_forward(torch.from_numpy(x1), torch.from_numpy(x2), torch.from_numpy(flow)).numpy()
    assert rel_err(t, gold["out"]) < TOL
