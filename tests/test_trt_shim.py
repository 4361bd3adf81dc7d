"""The TensorRT plugin shim is real code: cerberusnet_b200/csrc/trt_plugin_shim.cpp is compiled against
tests/trt_stub/NvInfer.h (a minimal stand-in for the public plugin API -- TensorRT itself is not in the image) and
its three plugin classes are driven through their creators by tests/trt_stub/harness.cpp, the way the reference's
runtime uses them (runtime/cerberus_net/trt_plugins/correlation.hpp:10-108, grid_sampler.hpp:21-112):

  CPU  creator lookup by (type, version); field names and defaults; getOutputDimensions; 24 / 32 / 9-byte
       serialisation; serialize -> deserializePlugin -> serialize; clone; supportsFormatCombination; zero workspace
  GPU  enqueue through the plugin == the C-ABI call, bit for bit; wrong output extent -> error code, not abort
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

CSRC = os.path.join(ROOT, "cerberusnet_b200", "csrc")
STUB = os.path.join(ROOT, "tests", "trt_stub")
PKG = os.path.join(ROOT, "cerberusnet_b200")


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    import cerberusnet_b200 as cb
    cb.lib()   # builds libcerberus_costvolume.so if needed
    exe = str(tmp_path_factory.mktemp("trt") / "trt_harness")
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Werror=return-type", f"-I{STUB}", f"-I{cuda}/include", "-o", exe,
           os.path.join(STUB, "harness.cpp"), os.path.join(CSRC, "trt_plugin_shim.cpp"), f"-L{PKG}", "-lcerberus_costvolume",
           f"-L{cuda}/lib64", "-lcudart", f"-Wl,-rpath,{PKG}", f"-Wl,-rpath,{cuda}/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-4000:]
    return exe


def test_shim_compiles_and_plugins_behave_like_the_reference_on_cpu(harness):
    r = subprocess.run([harness], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout[-3000:] + r.stderr[-2000:]


def test_shim_is_empty_without_tensorrt_headers(tmp_path):
    """Without <NvInfer.h> on the include path the translation unit must compile to nothing (the product build
    never needs TensorRT)."""
    obj = str(tmp_path / "shim.o")
    r = subprocess.run(["g++", "-std=c++17", "-c", os.path.join(CSRC, "trt_plugin_shim.cpp"), "-o", obj],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    syms = subprocess.run(["nm", "--defined-only", obj], capture_output=True, text=True).stdout
    assert "Plugin" not in syms


def test_new_trt_field_blocks_serialise_like_the_reference():
    """grid_sampler: bool + int + int = 9 bytes (grid_sampler.cpp:57-74), defaults false/Bilinear/Border (:40-42);
    warp_correlation: the six correlation ints + warp_mode + leaky_slope = 32 bytes."""
    import cerberusnet_b200 as cb
    from cerberusnet_b200 import _lib
    L = cb.lib()
    g = _lib.TrtGridSamplerFields()
    L.cerb_trt_grid_sampler_default_fields(ctypes.byref(g))
    assert (g.align_corners, g.interpolation_mode, g.padding_mode) == (0, 0, 1)
    assert L.cerb_trt_grid_sampler_serialize(ctypes.byref(g), None) == 9
    g.align_corners, g.interpolation_mode, g.padding_mode = 1, 1, 2
    buf = ctypes.create_string_buffer(9)
    assert L.cerb_trt_grid_sampler_serialize(ctypes.byref(g), buf) == 9
    assert buf.raw == bytes([1]) + np.array([1, 2], np.int32).tobytes()
    g2 = _lib.TrtGridSamplerFields()
    assert L.cerb_trt_grid_sampler_deserialize(buf, 9, ctypes.byref(g2)) == 0
    assert (g2.align_corners, g2.interpolation_mode, g2.padding_mode) == (1, 1, 2)
    assert L.cerb_trt_grid_sampler_deserialize(buf, 8, ctypes.byref(g2)) != 0
    w = _lib.TrtWarpCorrFields()
    L.cerb_trt_warp_corr_default_fields(ctypes.byref(w))
    assert (w.corr.pad_size, w.corr.kernel_size, w.corr.max_displacement, w.warp_mode) == (4, 1, 4, cb.WARP_TRT)
    assert abs(w.leaky_slope - 0.1) < 1e-7
    assert L.cerb_trt_warp_corr_serialize(ctypes.byref(w), None) == 32
    buf = ctypes.create_string_buffer(32)
    L.cerb_trt_warp_corr_serialize(ctypes.byref(w), buf)
    assert buf.raw[:24] == np.array([4, 1, 4, 1, 1, 1], np.int32).tobytes()
    w2 = _lib.TrtWarpCorrFields()
    assert L.cerb_trt_warp_corr_deserialize(buf, 32, ctypes.byref(w2)) == 0 and w2.warp_mode == w.warp_mode


@pytest.mark.gpu
def test_plugin_enqueue_matches_the_c_abi_on_the_gpu(harness):
    r = subprocess.run([harness, "--gpu"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and r.stdout.strip().endswith("OK"), r.stdout[-3000:] + r.stderr[-2000:]


@pytest.mark.gpu
def test_enqueue_i64_descriptor_variant():
    """TensorRT >= 10 widens Dims::d to int64: cerb_trt_corr_enqueue_i64 must give the same bits as the int32 entry."""
    import torch
    import cerberusnet_b200 as cb
    from cerberusnet_b200 import _lib
    L = cb.lib()
    dev = torch.device("cuda:0")
    torch.manual_seed(5)
    N, C, H, W = 2, 40, 24, 64
    x1, x2 = torch.randn(N, C, H, W, device=dev), torch.randn(N, C, H, W, device=dev)
    f = _lib.TrtCorrFields()
    L.cerb_trt_corr_default_fields(ctypes.byref(f))

    def mk(cls, dims):
        d = cls()
        d.dims.nbDims = 4
        for i, v in enumerate(dims):
            d.dims.d[i] = v
        d.type, d.format, d.scale = 0, 0, 1.0
        return d

    outs = []
    for cls, fn in ((_lib.TrtTensorDesc, L.cerb_trt_corr_enqueue), (_lib.TrtTensorDesc64, L.cerb_trt_corr_enqueue_i64)):
        ind = (cls * 2)(mk(cls, (N, C, H, W)), mk(cls, (N, C, H, W)))
        outd = (cls * 1)(mk(cls, (N, 81, H, W)))
        out = torch.empty(N, 81, H, W, device=dev)
        ins = (ctypes.c_void_p * 2)(x1.data_ptr(), x2.data_ptr())
        os_ = (ctypes.c_void_p * 1)(out.data_ptr())
        rc = fn(ctypes.byref(f), ind, outd, ins, os_, None, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert rc == 0
        torch.cuda.synchronize()
        outs.append(out)
        bad = (cls * 1)(mk(cls, (N, 81, H, W + 1)))
        assert fn(ctypes.byref(f), ind, bad, ins, os_, None, None) < 0   # CERB_ESHAPE, not an abort
    assert torch.equal(outs[0], outs[1])
    ref = cb.ops.warp_corr_forward(x1, x2, None, 4, 1, 4, 1, 1)
    assert torch.equal(outs[0], ref)
