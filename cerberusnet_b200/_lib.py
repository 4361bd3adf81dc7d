"""ctypes binding of libcerberus_costvolume.so (the C ABI in include/cerberus_costvolume.h).

The library is the product: there is no CPU or PyTorch fallback.  If it is missing it is built
in-tree with nvcc; if that fails, or a tensor is not on a CUDA device, the call raises.
"""
from __future__ import annotations

import ctypes
import math
import os
from typing import Optional

import torch

from . import build as _build

CERB_F32, CERB_F16, CERB_BF16 = 0, 1, 2
WARP_TORCH, WARP_TRT, WARP_TORCH_CPU = 0, 1, 2
VARIANT_AUTO, VARIANT_FAST, VARIANT_FAST_NOTMA, VARIANT_SMALL, VARIANT_SMALL_NOTMA, VARIANT_GENERIC, VARIANT_MID, VARIANT_TC = range(8)

_DTYPES = {torch.float32: CERB_F32, torch.float16: CERB_F16, torch.bfloat16: CERB_BF16}


class CorrParams(ctypes.Structure):
    """struct cerb_corr_params (include/cerberus_costvolume.h)."""
    _fields_ = [
        ("batch", ctypes.c_int32), ("channels", ctypes.c_int32), ("height", ctypes.c_int32), ("width", ctypes.c_int32),
        ("pad_size", ctypes.c_int32), ("kernel_size", ctypes.c_int32), ("max_displacement", ctypes.c_int32),
        ("stride1", ctypes.c_int32), ("stride2", ctypes.c_int32), ("corr_multiply", ctypes.c_int32),
        ("dtype", ctypes.c_int32), ("warp_mode", ctypes.c_int32), ("leaky_slope", ctypes.c_float),
        ("x2_batch_roll", ctypes.c_int32),
        ("x1_stride", ctypes.c_int64 * 4), ("x2_stride", ctypes.c_int64 * 4),
        ("flow_stride", ctypes.c_int64 * 4), ("out_stride", ctypes.c_int64 * 4),
    ]


class TrtDims(ctypes.Structure):
    _fields_ = [("nbDims", ctypes.c_int32), ("d", ctypes.c_int32 * 8)]


class TrtTensorDesc(ctypes.Structure):
    """Layout of nvinfer1::PluginTensorDesc (TensorRT 7/8)."""
    _fields_ = [("dims", TrtDims), ("type", ctypes.c_int32), ("format", ctypes.c_int32), ("scale", ctypes.c_float)]


class TrtDims64(ctypes.Structure):
    """nvinfer1::Dims of TensorRT >= 10 (int64 extents)."""
    _fields_ = [("nbDims", ctypes.c_int32), ("d", ctypes.c_int64 * 8)]


class TrtTensorDesc64(ctypes.Structure):
    """Layout of nvinfer1::PluginTensorDesc with TensorRT >= 10 dims."""
    _fields_ = [("dims", TrtDims64), ("type", ctypes.c_int32), ("format", ctypes.c_int32), ("scale", ctypes.c_float)]


class TrtCorrFields(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in
                ("pad_size", "kernel_size", "max_displacement", "stride1", "stride2", "corr_multiply")]


class TrtGridSamplerFields(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int32) for n in ("align_corners", "interpolation_mode", "padding_mode")]


class TrtWarpCorrFields(ctypes.Structure):
    _fields_ = [("corr", TrtCorrFields), ("warp_mode", ctypes.c_int32), ("leaky_slope", ctypes.c_float)]


EXPORTS = [
    "cerb_abi_version", "cerb_error_string", "cerb_corr_output_dims", "cerb_warp_corr_forward",
    "cerb_warp_corr_forward_variant", "cerb_warp_corr_forward_upflow", "cerb_warp_corr_backward_workspace",
    "cerb_warp_corr_backward",
    "cerb_flow_warp_forward", "cerb_flow_warp_backward", "cerb_warp_corr_forward_host_workspace",
    "cerb_warp_corr_forward_host", "cerb_launch_count",
    "cerb_trt_corr_default_fields", "cerb_trt_corr_serialization_size", "cerb_trt_corr_serialize",
    "cerb_trt_corr_deserialize", "cerb_trt_corr_output_dims", "cerb_trt_corr_supports_format",
    "cerb_trt_corr_workspace_size", "cerb_trt_corr_enqueue", "cerb_trt_corr_enqueue_i64",
    "cerb_trt_warp_corr_enqueue", "cerb_measure_fma_peak", "cerb_grid_sample_forward",
    "cerb_trt_warp_corr_default_fields", "cerb_trt_warp_corr_serialize", "cerb_trt_warp_corr_deserialize",
    "cerb_trt_grid_sampler_default_fields", "cerb_trt_grid_sampler_serialize", "cerb_trt_grid_sampler_deserialize",
    "cerb_trt_grid_sampler_enqueue",
    "cerb_photometric_workspace", "cerb_photometric_forward", "cerb_photometric_backward",
]

_lib: Optional[ctypes.CDLL] = None


def lib() -> ctypes.CDLL:
    """Load (building first if needed) the shared library.  Raises if it cannot be had."""
    global _lib
    if _lib is not None:
        return _lib
    path = _build.LIB_PATH
    override = os.environ.get("CERB_LIB_OVERRIDE")  # kernel A/B experiments (tools/ab_variants.py): an alternative build
    if override:
        path = override
    else:
        path = _build.build_library()   # no-op unless a source or header is newer than the library (or it is missing)
    L = ctypes.CDLL(path)
    vp, i32, f32p = ctypes.c_void_p, ctypes.c_int32, ctypes.c_void_p
    pp = ctypes.POINTER(CorrParams)
    L.cerb_abi_version.restype = ctypes.c_int
    L.cerb_error_string.restype = ctypes.c_char_p
    L.cerb_error_string.argtypes = [ctypes.c_int]
    L.cerb_corr_output_dims.argtypes = [pp] + [ctypes.POINTER(i32)] * 3
    L.cerb_warp_corr_forward.argtypes = [pp, vp, vp, f32p, vp, vp]
    i64x4 = ctypes.POINTER(ctypes.c_int64)
    L.cerb_warp_corr_forward_upflow.argtypes = [pp, vp, vp, f32p, i64x4, f32p, i64x4, vp, vp]
    L.cerb_warp_corr_forward_variant.argtypes = [pp, vp, vp, f32p, vp, ctypes.c_int, vp]
    L.cerb_warp_corr_backward_workspace.argtypes = [pp, ctypes.c_int]
    L.cerb_warp_corr_backward_workspace.restype = ctypes.c_size_t
    L.cerb_warp_corr_backward.argtypes = [pp, vp, vp, f32p, vp, vp, vp, vp, f32p, vp, ctypes.c_size_t, vp]
    L.cerb_flow_warp_forward.argtypes = [vp, f32p, vp, i32, i32, i32, i32, i32, i32, vp]
    L.cerb_flow_warp_backward.argtypes = [vp, f32p, vp, vp, f32p, i32, i32, i32, i32, i32, i32, vp]
    L.cerb_warp_corr_forward_host_workspace.argtypes = [pp, ctypes.c_int]
    L.cerb_warp_corr_forward_host_workspace.restype = ctypes.c_size_t
    L.cerb_warp_corr_forward_host.argtypes = [pp, vp, vp, f32p, vp, vp, ctypes.c_size_t, vp]
    L.cerb_launch_count.restype = ctypes.c_uint64
    fp = ctypes.POINTER(TrtCorrFields)
    dp = ctypes.POINTER(TrtTensorDesc)
    L.cerb_trt_corr_default_fields.argtypes = [fp]
    L.cerb_trt_corr_default_fields.restype = None
    L.cerb_trt_corr_serialization_size.restype = ctypes.c_size_t
    L.cerb_trt_corr_serialize.argtypes = [fp, vp]
    L.cerb_trt_corr_serialize.restype = ctypes.c_size_t
    L.cerb_trt_corr_deserialize.argtypes = [vp, ctypes.c_size_t, fp]
    L.cerb_trt_corr_output_dims.argtypes = [fp, ctypes.POINTER(TrtDims), ctypes.POINTER(TrtDims)]
    L.cerb_trt_corr_supports_format.argtypes = [ctypes.c_int, dp, ctypes.c_int, ctypes.c_int]
    L.cerb_trt_corr_workspace_size.argtypes = [fp, dp, ctypes.c_int, dp, ctypes.c_int]
    L.cerb_trt_corr_workspace_size.restype = ctypes.c_size_t
    L.cerb_trt_corr_enqueue.argtypes = [fp, dp, dp, ctypes.POINTER(vp), ctypes.POINTER(vp), vp, vp]
    L.cerb_trt_warp_corr_enqueue.argtypes = [fp, i32, ctypes.c_float, dp, dp, ctypes.POINTER(vp), ctypes.POINTER(vp),
                                             vp, vp]
    dp64 = ctypes.POINTER(TrtTensorDesc64)
    L.cerb_trt_corr_enqueue_i64.argtypes = [fp, dp64, dp64, ctypes.POINTER(vp), ctypes.POINTER(vp), vp, vp]
    L.cerb_measure_fma_peak.argtypes = [ctypes.POINTER(ctypes.c_double), vp]
    L.cerb_grid_sample_forward.argtypes = [vp, vp, vp] + [i32] * 11 + [vp]
    L.cerb_photometric_workspace.argtypes = [i32, i32, i32]
    L.cerb_photometric_workspace.restype = ctypes.c_size_t
    L.cerb_photometric_forward.argtypes = [vp, vp, vp, vp, vp, ctypes.c_size_t, i32, i32, i32, i32, ctypes.c_float, ctypes.c_float, i32, vp]
    L.cerb_photometric_backward.argtypes = [vp, vp, vp, vp, vp, i32, i32, i32, i32, ctypes.c_float, ctypes.c_float, i32, vp]
    gfp = ctypes.POINTER(TrtGridSamplerFields)
    L.cerb_trt_grid_sampler_default_fields.argtypes = [gfp]
    L.cerb_trt_grid_sampler_default_fields.restype = None
    L.cerb_trt_grid_sampler_serialize.argtypes = [gfp, vp]
    L.cerb_trt_grid_sampler_serialize.restype = ctypes.c_size_t
    L.cerb_trt_grid_sampler_deserialize.argtypes = [vp, ctypes.c_size_t, gfp]
    L.cerb_trt_grid_sampler_enqueue.argtypes = [gfp, dp, dp, ctypes.POINTER(vp), ctypes.POINTER(vp), vp, vp]
    wfp = ctypes.POINTER(TrtWarpCorrFields)
    L.cerb_trt_warp_corr_default_fields.argtypes = [wfp]
    L.cerb_trt_warp_corr_default_fields.restype = None
    L.cerb_trt_warp_corr_serialize.argtypes = [wfp, vp]
    L.cerb_trt_warp_corr_serialize.restype = ctypes.c_size_t
    L.cerb_trt_warp_corr_deserialize.argtypes = [vp, ctypes.c_size_t, wfp]
    L.cerb_debug_set_path_counters.argtypes = [vp]
    L.cerb_debug_set_path_counters.restype = None
    L.cerb_debug_set_backward_kernel.argtypes = [ctypes.c_int]
    L.cerb_debug_set_backward_kernel.restype = None
    if L.cerb_abi_version() != 1:
        raise RuntimeError("libcerberus_costvolume.so: ABI version mismatch")
    _lib = L
    return L


class CostVolumeError(RuntimeError):
    """Raised when the C ABI returns non-zero (the reference raises RuntimeError through
    AT_ERROR, correlation_cuda.cpp:22-23)."""


def check(code: int, what: str):
    if code != 0:
        msg = lib().cerb_error_string(code).decode()
        raise CostVolumeError(f"{what} failed: {msg} (code {code})")


def dtype_code(t: torch.Tensor) -> int:
    try:
        return _DTYPES[t.dtype]
    except KeyError:
        raise CostVolumeError(f"unsupported dtype {t.dtype}: float32, float16 and bfloat16 are implemented") from None


def require_cuda(*tensors: torch.Tensor):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise CostVolumeError("cerberusnet_b200 has no CPU path: tensors must live on a CUDA device "
                                  "(reference op is CUDA-only too, correlation_cuda.cpp:45-48)")


def _strides(t: Optional[torch.Tensor]):
    if t is None:
        return (ctypes.c_int64 * 4)(0, 0, 0, 0)
    return (ctypes.c_int64 * 4)(*t.stride())


def make_params(x1: torch.Tensor, x2: torch.Tensor, flow: Optional[torch.Tensor], out: Optional[torch.Tensor],
                pad_size: int, kernel_size: int, max_displacement: int, stride1: int, stride2: int,
                corr_multiply: int, warp_mode: int, leaky_slope: Optional[float], x2_roll: int = 0) -> CorrParams:
    B, C, H, W = x1.shape
    p = CorrParams()
    p.batch, p.channels, p.height, p.width = B, C, H, W
    p.pad_size, p.kernel_size, p.max_displacement = int(pad_size), int(kernel_size), int(max_displacement)
    p.stride1, p.stride2, p.corr_multiply = int(stride1), int(stride2), int(corr_multiply)
    p.dtype = dtype_code(x1)
    p.warp_mode = int(warp_mode)
    p.leaky_slope = math.nan if leaky_slope is None else float(leaky_slope)
    p.x2_batch_roll = int(x2_roll)
    p.x1_stride = _strides(x1)
    p.x2_stride = _strides(x2)
    p.flow_stride = _strides(flow)
    p.out_stride = _strides(out)
    return p


_PARAM_CACHE: dict = {}


def make_params_cached(x1, x2, flow, out, pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply,
                       warp_mode, leaky_slope, x2_roll=0) -> CorrParams:
    """`make_params` memoised on everything the block depends on (shapes, dtype, strides, configuration): building
    the ctypes structure costs more host time than the launch itself.  The blocks are read-only after creation."""
    key = (x1.shape, x1.dtype, x1.stride(), x2.stride(), None if flow is None else flow.stride(),
           None if out is None else out.stride(), pad_size, kernel_size, max_displacement, stride1, stride2,
           corr_multiply, warp_mode, leaky_slope, x2_roll)
    p = _PARAM_CACHE.get(key)
    if p is None:
        if len(_PARAM_CACHE) > 4096:
            _PARAM_CACHE.clear()
        p = make_params(x1, x2, flow, out, pad_size, kernel_size, max_displacement, stride1, stride2, corr_multiply,
                        warp_mode, leaky_slope, x2_roll)
        _PARAM_CACHE[key] = p
    return p


class _NullGuard:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NULL_GUARD = _NullGuard()


def device_guard(device):
    """Make `device` current for the call, without paying for a context switch when it already is."""
    if device.index is None or device.index == torch.cuda.current_device():
        return _NULL_GUARD
    return torch.cuda.device(device)


def output_dims(p: CorrParams):
    oc, oh, ow = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int32()
    check(lib().cerb_corr_output_dims(ctypes.byref(p), ctypes.byref(oc), ctypes.byref(oh), ctypes.byref(ow)),
          "cerb_corr_output_dims")
    return oc.value, oh.value, ow.value


def current_stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def ptr(t: Optional[torch.Tensor]):
    return None if t is None else ctypes.c_void_p(t.data_ptr())
