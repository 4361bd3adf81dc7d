"""The drop-in boundary without a GPU: the C-ABI library loads and exports every symbol the
headers declare, argument errors come back as codes (never aborts), the TensorRT-shaped helpers
behave like the reference plugin's host side, the Python surfaces mirror the reference module, and
the product never touches the oracle."""
import ctypes
import math
import os
import re
import sys

import pytest
import torch

import cerberusnet_b200 as cb
from cerberusnet_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    names = set()
    for h in ("cerberus_costvolume.h", "cerberus_trt_plugin.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        names |= set(re.findall(r"CERB_API\s+[\w\s\*]+?\b(cerb_\w+)\s*\(", text))
    return names


def test_library_exports_every_declared_symbol():
    lib = cb.lib()
    declared = _declared_symbols()
    assert len(declared) >= 20
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert getattr(lib, name) is not None
    assert lib.cerb_abi_version() == 1


def _params(B=1, C=8, H=16, W=32, pad=4, k=1, md=4, s1=1, s2=1, dtype=0):
    p = _lib.CorrParams()
    p.batch, p.channels, p.height, p.width = B, C, H, W
    p.pad_size, p.kernel_size, p.max_displacement, p.stride1, p.stride2, p.corr_multiply = pad, k, md, s1, s2, 1
    p.dtype = dtype
    p.leaky_slope = math.nan
    return p


@pytest.mark.parametrize("shape,params,expect", [
    ((64, 128), (4, 1, 4, 1, 1), (81, 64, 128)),     # production config, BASELINE configs[0]
    ((128, 64), (4, 1, 10, 1, 1), (441, 116, 52)),   # reference __main__ block, correlation.py:85-87
    ((20, 28), (3, 3, 4, 2, 2), (25, 8, 12)),
    ((375, 1242), (8, 1, 8, 1, 1), (289, 375, 1242)),  # KITTI-shaped, configs[4]
])
def test_output_dims_rule(shape, params, expect):
    H, W = shape
    p = _params(H=H, W=W, pad=params[0], k=params[1], md=params[2], s1=params[3], s2=params[4])
    assert _lib.output_dims(p) == expect


def test_argument_errors_are_codes_not_aborts():
    lib = cb.lib()
    oc = ctypes.c_int32()
    p = _params(H=12, W=20, md=10)  # empty output
    assert lib.cerb_corr_output_dims(ctypes.byref(p), ctypes.byref(oc), None, None) == -2
    p = _params(C=0)
    assert lib.cerb_corr_output_dims(ctypes.byref(p), None, None, None) == -1
    p = _params(dtype=7)
    assert lib.cerb_corr_output_dims(ctypes.byref(p), None, None, None) == -1
    p = _params()
    p.x1_stride = (ctypes.c_int64 * 4)(4096, 512, 32, 2)  # innermost stride must be 1
    assert lib.cerb_corr_output_dims(ctypes.byref(p), None, None, None) == -3
    assert lib.cerb_corr_output_dims(None, None, None, None) == -1
    # null tensors are rejected before anything is launched
    p = _params()
    assert lib.cerb_warp_corr_forward(ctypes.byref(p), None, None, None, None, None) == -1
    assert lib.cerb_warp_corr_backward(ctypes.byref(p), None, None, None, None, None, None, None, None, None, 0, None) == -1
    assert lib.cerb_flow_warp_forward(None, None, None, 1, 1, 8, 8, 0, 0, None) == -1
    for code in (0, -1, -2, -3, -4, -5, 1, 700):
        assert len(lib.cerb_error_string(code)) > 0
    # with a flow: the re-materialised warped map + the gradient wrt it; without: nothing
    assert lib.cerb_warp_corr_backward_workspace(ctypes.byref(_params()), 0) == 0
    assert lib.cerb_warp_corr_backward_workspace(ctypes.byref(_params(k=3, pad=5)), 0) == 0
    assert lib.cerb_warp_corr_backward_workspace(ctypes.byref(_params()), 1) == 2 * 8 * 16 * 32 * 4
    assert lib.cerb_warp_corr_backward_workspace(ctypes.byref(_params(dtype=2)), 1) == (2 * 2 + 4) * 8 * 16 * 32
    assert lib.cerb_warp_corr_backward_workspace(ctypes.byref(_params(k=3, pad=5)), 1) == 2 * 8 * 16 * 32 * 4


def test_trt_host_side_matches_reference_plugin():
    """trt_plugins/correlation.cpp: defaults 4,1,4,1,1,1 (:54-62); 24-byte serialisation of six
    ints in field order (:65-105); output dims (:178-205); format support (:144-176); and a zero
    workspace instead of 2*N*C*(H+2p)*(W+2p)*sizeof(T) (:114-136)."""
    lib = cb.lib()
    f = _lib.TrtCorrFields()
    lib.cerb_trt_corr_default_fields(ctypes.byref(f))
    assert [getattr(f, n) for n, _ in f._fields_] == [4, 1, 4, 1, 1, 1]
    assert lib.cerb_trt_corr_serialization_size() == 24
    f.max_displacement, f.pad_size = 8, 6
    buf = (ctypes.c_char * 24)()
    assert lib.cerb_trt_corr_serialize(ctypes.byref(f), buf) == 24
    import struct
    assert struct.unpack("<6i", bytes(buf)) == (6, 1, 8, 1, 1, 1)
    g = _lib.TrtCorrFields()
    assert lib.cerb_trt_corr_deserialize(buf, 24, ctypes.byref(g)) == 0
    assert (g.pad_size, g.max_displacement) == (6, 8)
    assert lib.cerb_trt_corr_deserialize(buf, 23, ctypes.byref(g)) == -1

    lib.cerb_trt_corr_default_fields(ctypes.byref(f))
    din, dout = _lib.TrtDims(), _lib.TrtDims()
    din.nbDims = 4
    for i, v in enumerate((2, 48, 128, 256)):
        din.d[i] = v
    assert lib.cerb_trt_corr_output_dims(ctypes.byref(f), ctypes.byref(din), ctypes.byref(dout)) == 0
    assert (dout.nbDims, list(dout.d[:4])) == (4, [2, 81, 128, 256])

    descs = (_lib.TrtTensorDesc * 3)()
    for i, dims in enumerate(((2, 48, 128, 256), (2, 48, 128, 256), (2, 81, 128, 256))):
        descs[i].dims.nbDims = 4
        for j, v in enumerate(dims):
            descs[i].dims.d[j] = v
        descs[i].type, descs[i].format = 0, 0
    assert all(lib.cerb_trt_corr_supports_format(pos, descs, 2, 1) == 1 for pos in range(3))
    descs[1].type = 1  # mixed float / half inputs
    assert lib.cerb_trt_corr_supports_format(1, descs, 2, 1) == 0
    descs[1].type = 0
    descs[2].format = 1  # not kLINEAR
    assert lib.cerb_trt_corr_supports_format(2, descs, 2, 1) == 0
    assert lib.cerb_trt_corr_workspace_size(ctypes.byref(f), descs, 2, descs, 1) == 0
    # enqueue validates descriptors before touching the device
    ins = (ctypes.c_void_p * 2)(None, None)
    outs = (ctypes.c_void_p * 1)(None)
    descs[2].format = 0
    descs[1].dims.d[3] = 255
    assert lib.cerb_trt_corr_enqueue(ctypes.byref(f), descs, ctypes.byref(descs[2]), ins, outs, None, None) == -1


def test_python_surface_mirrors_reference_module():
    """correlation.py:23-80: constructor arguments and defaults, attribute names, no parameters or
    buffers, Function.apply arity."""
    import inspect
    m = cb.Correlation()
    assert (m.pad_size, m.kernel_size, m.max_displacement, m.stride1, m.stride2, m.corr_multiply) == (0, 0, 0, 1, 2, 1)
    m = cb.Correlation(pad_size=4, kernel_size=1, max_displacement=4, stride1=1, stride2=1, corr_multiply=1)
    assert list(m.parameters()) == [] and list(m.buffers()) == [] and m.state_dict() == {}
    sig = inspect.signature(cb.CorrelationFunction.forward)
    assert list(sig.parameters)[1:] == ["input1", "input2", "pad_size", "kernel_size", "max_displacement", "stride1",
                                        "stride2", "corr_multiply"]
    assert [p.default for p in list(sig.parameters.values())[3:]] == [3, 3, 20, 1, 2, 1]
    assert list(inspect.signature(cb.flow_warp).parameters)[:4] == ["image", "flow12", "pad", "mode"]
    t = cb.CorrelationTorch(max_displacement=4)
    assert (t.output_dim, t.pad_size) == (9, 4)


def test_no_cpu_fallback():
    x = torch.zeros(1, 4, 8, 8)
    with pytest.raises(cb.CostVolumeError):
        cb.Correlation(4, 1, 4, 1, 1, 1)(x, x)
    with pytest.raises(cb.CostVolumeError):
        cb.flow_warp(x, torch.zeros(1, 2, 8, 8))
    with pytest.raises(NotImplementedError):
        cb.flow_warp(x, torch.zeros(1, 2, 8, 8), pad="zeros")


def test_install_registers_reference_import_path():
    mod = cb.install()
    import importlib
    ref = importlib.import_module("nnet_training.correlation_package.correlation")
    assert ref is mod and ref.Correlation is cb.Correlation
    from nnet_training.correlation_package.correlation import Correlation, CorrelationFunction  # noqa: F401
    assert torch.ops.cerberus_b200.correlation is not None
    assert torch.ops.cerberus.correlation is not None  # schema of correlation_cuda.cpp:45-48
    # fake / meta implementation gives the reference output shape
    meta = torch.ops.cerberus_b200.correlation(torch.empty(2, 16, 24, 40, device="meta"),
                                               torch.empty(2, 16, 24, 40, device="meta"), 4, 1, 4, 1, 1, 1)
    assert tuple(meta.shape) == (2, 81, 24, 40)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "cerberusnet_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                text = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), fn
                assert "libcostvolume_oracle" not in text, fn
    assert "oracle" not in {m.split(".")[0] for m in sys.modules if m.startswith("cerberusnet_b200")}
