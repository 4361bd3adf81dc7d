"""Round-2 parity additions ("close the parity holes"):

  * the training shapes themselves -- HRNetV2-W48 levels at batch 8 (BASELINE configs[3]: C=384 16x32 ... C=48 128x256)
    forward and backward against the reference's own CUDA op (oracle/_ref/correlation_ref.so, fresh process) and,
    fused with a flow, against the C oracle;
  * BASELINE configs[4] shapes -- KITTI 1248x384 HRNet pyramid incl. W = 39 and 78 (rows not 16-byte aligned: the
    non-TMA staging path), max_displacement 8, forward and backward against the C oracle;
  * x2_batch_roll (both flow directions in one launch) against two launches with swapped inputs.

Tolerance: 1e-5 of max|ref| (north_star), written at each assert.
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import cerberusnet_b200 as cb
from cerberusnet_b200 import ops
from conftest import ROOT, rel_err
from oracle import c_oracle as co

pytestmark = pytest.mark.gpu
TOL = 1e-5
REF_SO = os.path.join(ROOT, "oracle", "_ref", "correlation_ref.so")
HRNET_B8 = [(8, 384, 16, 32), (8, 192, 32, 64), (8, 96, 64, 128), (8, 48, 128, 256)]
KITTI_HRNET = [(2, 384, 12, 39), (2, 192, 24, 78), (2, 96, 48, 156), (2, 48, 96, 312)]


def dev():
    return torch.device("cuda:0")


WORKER = r"""
import json, sys
import numpy as np, torch
sys.path.insert(0, %(root)r)
torch.ops.load_library(%(so)r)
import cerberusnet_b200 as cb
from cerberusnet_b200 import ops
ref_ops = torch.ops.cerberus
def rel(a, b):
    a = a.double(); b = b.double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))
res = {}
for shape in %(shapes)r:
    torch.manual_seed(sum(shape))
    x1 = torch.randn(*shape, device="cuda"); x2 = torch.randn(*shape, device="cuda")
    want = ref_ops.correlation(x1, x2, 4, 1, 4, 1, 1, 1)
    got = ops.warp_corr_forward(x1, x2, None, 4, 1, 4, 1, 1)
    g = torch.randn_like(want)
    w1, w2 = ref_ops.correlation_backward(x1, x2, g, 4, 1, 4, 1, 1, 1)
    g1, g2, _ = ops.warp_corr_backward(x1, x2, None, None, g, 4, 1, 4, 1, 1)
    res[str(shape)] = {"fwd": rel(got, want), "g1": rel(g1, w1), "g2": rel(g2, w2)}
print("RESULT " + json.dumps(res))
"""


def test_hrnet_training_shapes_batch8_vs_reference_cuda_op():
    if not os.path.exists(REF_SO):
        pytest.skip("oracle/_ref/correlation_ref.so not built (needs the reference checkout at build time)")
    r = subprocess.run([sys.executable, "-c", WORKER % {"root": ROOT, "so": REF_SO, "shapes": HRNET_B8}],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-3000:]
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1][len("RESULT "):])
    assert len(res) == 4
    for name, e in res.items():
        assert e["fwd"] < TOL and e["g1"] < TOL and e["g2"] < TOL, (name, e)


@pytest.mark.parametrize("shape", HRNET_B8)
def test_hrnet_training_shapes_batch8_fused_vs_oracle(shape):
    """Fused warp + correlation + LeakyReLU forward and backward at the training shapes; the C oracle checks batch
    items 0 and 7 (first and last: the persistent tile loop crosses every item in between)."""
    B, C, H, W = shape
    rs = np.random.RandomState(C)
    x1 = rs.standard_normal(shape).astype(np.float32)
    x2 = rs.standard_normal(shape).astype(np.float32)
    fl = (rs.standard_normal((B, 2, H, W)) * 2.0).astype(np.float32)
    t1, t2, tf = (torch.from_numpy(a).to(dev()) for a in (x1, x2, fl))
    out = ops.warp_corr_forward(t1, t2, tf, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1)
    g = rs.standard_normal(out.shape).astype(np.float32)
    g[np.abs(out.cpu().numpy()) < 1e-6 * float(out.abs().max())] = 0.0   # outputs within rounding of zero: LeakyReLU sign is not defined
    g1, g2, gf = ops.warp_corr_backward(t1, t2, tf, out, torch.from_numpy(g).to(dev()), 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1)
    for n in (0, B - 1):
        s = slice(n, n + 1)
        ref = co.level_forward(x1[s], x2[s], fl[s], 4, 1, 4, 1, 1, co.WARP_TORCH, 0.1)
        assert rel_err(out[s].cpu().numpy(), ref) < TOL
        r1, r2, rf = co.level_backward(x1[s], x2[s], fl[s], g[s], 4, 1, 4, 1, 1, co.WARP_TORCH, 0.1)
        assert rel_err(g1[s].cpu().numpy(), r1) < TOL
        assert rel_err(g2[s].cpu().numpy(), r2) < TOL
        assert rel_err(gf[s].cpu().numpy(), rf) < TOL


@pytest.mark.parametrize("shape", KITTI_HRNET)
@pytest.mark.parametrize("with_flow", [False, True])
def test_kitti_md8_shapes_incl_unaligned_widths(shape, with_flow):
    """BASELINE configs[4]: 1248x384 HRNet pyramid, max_displacement 8 (289 planes).  W = 39 and 78 break the 16-byte
    row alignment TMA needs, so those levels run the LDG staging path; all four are checked against the oracle."""
    B, C, H, W = shape
    rs = np.random.RandomState(W)
    x1 = rs.standard_normal(shape).astype(np.float32)
    x2 = rs.standard_normal(shape).astype(np.float32)
    fl = (rs.standard_normal((B, 2, H, W)) * 2.5).astype(np.float32) if with_flow else None
    t1, t2 = torch.from_numpy(x1).to(dev()), torch.from_numpy(x2).to(dev())
    tf = torch.from_numpy(fl).to(dev()) if with_flow else None
    out = ops.warp_corr_forward(t1, t2, tf, 8, 1, 8, 1, 1, 1, cb.WARP_TORCH, 0.1)
    assert tuple(out.shape) == (B, 289, H, W)
    ref = co.level_forward(x1, x2, fl, 8, 1, 8, 1, 1, co.WARP_TORCH, 0.1)
    assert rel_err(out.cpu().numpy(), ref) < TOL
    g = rs.standard_normal(ref.shape).astype(np.float32)
    # the oracle masks the LeakyReLU with its OWN forward: where that is within rounding of zero (one output in ~17 M
    # here) the two sign patterns may differ legitimately, so no gradient is sent through those outputs
    g[np.abs(ref) < 1e-6 * np.abs(ref).max()] = 0.0
    g1, g2, gf = ops.warp_corr_backward(t1, t2, tf, out, torch.from_numpy(g).to(dev()), 8, 1, 8, 1, 1, 1, cb.WARP_TORCH, 0.1)
    r1, r2, rf = co.level_backward(x1, x2, fl, g, 8, 1, 8, 1, 1, co.WARP_TORCH, 0.1)
    assert rel_err(g1.cpu().numpy(), r1) < TOL and rel_err(g2.cpu().numpy(), r2) < TOL
    if with_flow:
        assert rel_err(gf.cpu().numpy(), rf) < TOL


@pytest.mark.parametrize("shape", [(2, 32, 128, 256), (2, 64, 64, 128), (4, 96, 32, 64), (2, 128, 16, 32), (6, 20, 24, 40), (2, 192, 8, 16)])
@pytest.mark.parametrize("with_flow", [False, True])
def test_x2_batch_roll_both_flow_directions_in_one_launch(shape, with_flow):
    """x1 = x2 = features of [image 1; image 2] with x2_batch_roll = B/2 (cerb_corr_params.x2_batch_roll): the
    reference's consistency=True forward (pwcnet.py:108-113) in one launch per level.  Must equal -- bit for bit in
    the forward, to 1e-6 in the backward (atomics) -- two launches with explicitly swapped inputs; every kernel
    variant that can take the shape, and the generic kernel."""
    B, C, H, W = shape
    torch.manual_seed(B * C + W)
    f = torch.randn(B, C, H, W, device=dev())
    fl = torch.randn(B, 2, H, W, device=dev()) * 2.0 if with_flow else None
    roll = B // 2
    f_rolled = torch.roll(f, -roll, 0).contiguous()
    ref = ops.warp_corr_forward(f, f_rolled, fl, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1)
    for variant in (0, 1, 3, 5, 6):
        out = ops.warp_corr_forward(f, f, fl, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1, x2_roll=roll, variant=variant)
        ref_v = ops.warp_corr_forward(f, f_rolled, fl, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1, variant=variant)
        assert torch.equal(out, ref_v), variant
        assert rel_err(out.cpu().numpy(), ref.cpu().numpy()) < TOL
    g = torch.randn_like(ref)
    out = ops.warp_corr_forward(f, f, fl, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1, x2_roll=roll)
    g1, g2, gf = ops.warp_corr_backward(f, f, fl, out, g, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1, x2_roll=roll)
    k1, k2, kf = ops.warp_corr_backward(f, f_rolled, fl, ref, g, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1)
    assert rel_err(g1.cpu().numpy(), k1.cpu().numpy()) < 1e-6
    assert rel_err(g2.cpu().numpy(), torch.roll(k2, roll, 0).cpu().numpy()) < 1e-6   # grad_x2 comes back in x2's own batch order
    if with_flow:
        assert rel_err(gf.cpu().numpy(), kf.cpu().numpy()) < 1e-6


def test_x2_batch_roll_is_validated():
    f = torch.randn(2, 8, 16, 32, device=dev())
    with pytest.raises(cb.CostVolumeError):
        ops.warp_corr_forward(f, f, None, 4, 1, 4, 1, 1, x2_roll=2)
    with pytest.raises(cb.CostVolumeError):
        ops.warp_corr_forward(f, f, None, 4, 1, 4, 1, 1, x2_roll=-1)


def test_large_flow_variation_uses_the_large_raw_box():
    """A x2 up-sampled decoder-like flow (sigma 3 px and more) no longer drops 8x32 tiles to the direct-gather
    fallback: they take the large raw box.  Parity on those tiles, and the path mix."""
    import ctypes
    torch.manual_seed(12)
    B, C, H, W = 1, 32, 128, 256
    x1, x2 = torch.randn(B, C, H, W, device=dev()), torch.randn(B, C, H, W, device=dev())
    coarse = torch.randn(B, 2, H // 2, W // 2, device=dev()) * 2.5
    fl = torch.nn.functional.interpolate(coarse * 2, scale_factor=2, mode="bilinear", align_corners=True)
    L = cb.lib()

    def run(variant):
        ctr = torch.zeros(4, dtype=torch.int64, device=dev())
        L.cerb_debug_set_path_counters(ctypes.c_void_p(ctr.data_ptr()))
        try:
            o = ops.warp_corr_forward(x1, x2, fl, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1, variant=variant)
            torch.cuda.synchronize()
        finally:
            L.cerb_debug_set_path_counters(None)
        return o, int(ctr[1]), int(ctr[2]), int(ctr[3])

    out, small, direct, large = run(1)   # CUDA-core kernel, 8x32 tiles
    assert small + direct + large == 128
    assert large > 0 and direct <= 2, (small, large, direct)
    out_tc, raw_tc, direct_tc, _ = run(7)   # tensor-core kernel, 8x16 tiles: one box size, the rest gathers from global memory
    assert raw_tc + direct_tc == 256 and raw_tc > 0
    ref = ops.warp_corr_forward(x1, x2, fl, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1, variant=5)   # generic kernel (pinned to the oracle elsewhere)
    assert rel_err(out.cpu().numpy(), ref.cpu().numpy()) < TOL
    assert rel_err(out_tc.cpu().numpy(), ref.cpu().numpy()) < TOL
    ref_o = co.level_forward(x1.cpu().numpy(), x2.cpu().numpy(), fl.cpu().numpy(), 4, 1, 4, 1, 1, co.WARP_TORCH, 0.1)
    assert rel_err(out.cpu().numpy(), ref_o) < TOL
    assert rel_err(out_tc.cpu().numpy(), ref_o) < TOL


@pytest.mark.parametrize("dtype,tol", [(torch.float16, 6e-4), (torch.bfloat16, 5e-3)])
@pytest.mark.parametrize("with_flow", [False, True])
def test_half_precision_backward_rounds_once(dtype, tol, with_flow):
    """16-bit gradients at max_displacement 4: fp32 accumulation (the warp's splat too: fp32 workspace, narrowed once),
    so the only error against an exact evaluation on the same 16-bit inputs is the rounding of the stored result:
    2^-11 = 4.9e-4 (fp16) / 2^-8 = 3.9e-3 (bf16) of max|ref|."""
    rs = np.random.RandomState(17)
    B, C, H, W = 2, 24, 24, 64
    t1 = torch.from_numpy(rs.standard_normal((B, C, H, W)).astype(np.float32)).to(dev()).to(dtype)
    t2 = torch.from_numpy(rs.standard_normal((B, C, H, W)).astype(np.float32)).to(dev()).to(dtype)
    fl = (rs.standard_normal((B, 2, H, W)) * 2.0).astype(np.float32) if with_flow else None
    tf = torch.from_numpy(fl).to(dev()) if with_flow else None
    out = ops.warp_corr_forward(t1, t2, tf, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1)
    tg = torch.from_numpy(rs.standard_normal(tuple(out.shape)).astype(np.float32)).to(dev()).to(dtype)
    g1, g2, gf = ops.warp_corr_backward(t1, t2, tf, out, tg, 4, 1, 4, 1, 1, 1, cb.WARP_TORCH, 0.1)
    x1, x2, g = (t.float().cpu().numpy() for t in (t1, t2, tg))
    r1, r2, rf = co.level_backward(x1, x2, fl, g, 4, 1, 4, 1, 1, co.WARP_TORCH, 0.1)
    assert g1.dtype == dtype and g2.dtype == dtype
    assert rel_err(g1.float().cpu().numpy(), r1) < tol
    assert rel_err(g2.float().cpu().numpy(), r2) < tol
    if with_flow:
        # the gradient with respect to the warped map passes through a 16-bit workspace between the correlation backward
        # and the warp backward, so the flow gradient carries that one rounding too
        assert gf.dtype == torch.float32 and rel_err(gf.cpu().numpy(), rf) < tol
