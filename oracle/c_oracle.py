"""ctypes binding of oracle/costvolume_oracle.c (numpy float32 in / out).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libcostvolume_oracle.so")

ACC_REFERENCE_ORDER = 0
ACC_DOUBLE = 1
WARP_TORCH = 0      # ATen CUDA flavour (grid / (size-1) as reciprocal multiply): the reference's training path
WARP_TRT = 1
WARP_TORCH_CPU = 2  # ATen CPU flavour (true division): what the golden fixtures were generated with

_f32p = ctypes.POINTER(ctypes.c_float)
_lib = None


def build(force: bool = False) -> str:
    """Compile the C oracle in-tree (gcc, seconds)."""
    src = os.path.join(_HERE, "costvolume_oracle.c")
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "libcostvolume_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        i = ctypes.c_int
        ip = ctypes.POINTER(ctypes.c_int)
        L.cvo_corr_out_dims.argtypes = [i] * 7 + [ip, ip, ip]
        L.cvo_corr_forward.argtypes = [_f32p, _f32p, _f32p] + [i] * 10
        L.cvo_corr_backward.argtypes = [_f32p] * 5 + [i] * 9
        L.cvo_corr_backward_reforder.argtypes = [_f32p] * 5 + [i] * 7
        L.cvo_flow_warp_forward.argtypes = [_f32p] * 3 + [i] * 5
        L.cvo_flow_warp_backward.argtypes = [_f32p] * 5 + [i] * 5
        L.cvo_leaky_relu.argtypes = [_f32p, ctypes.c_size_t, ctypes.c_float]
        L.cvo_leaky_relu.restype = None
        L.cvo_leaky_relu_backward.argtypes = [_f32p, _f32p, ctypes.c_size_t, ctypes.c_float]
        L.cvo_leaky_relu_backward.restype = None
        L.cvo_level_forward.argtypes = [_f32p] * 5 + [i] * 10 + [ctypes.c_float, i]
        L.cvo_flow_upsample2x.argtypes = [_f32p, _f32p, i, i, i]
        _lib = L
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.float32 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(_f32p)


def _c(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float32))


def _check(rc: int, what: str):
    if rc != 0:
        raise ValueError(f"oracle {what}: invalid arguments (rc={rc})")


def corr_out_dims(H, W, pad, k, md, s1, s2):
    d2, oh, ow = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    rc = lib().cvo_corr_out_dims(H, W, pad, k, md, s1, s2, ctypes.byref(d2), ctypes.byref(oh), ctypes.byref(ow))
    _check(rc, "corr_out_dims")
    return d2.value, oh.value, ow.value


def corr_forward(x1, x2, pad, k, md, s1, s2, acc=ACC_DOUBLE) -> np.ndarray:
    x1, x2 = _c(x1), _c(x2)
    B, C, H, W = x1.shape
    d2, oh, ow = corr_out_dims(H, W, pad, k, md, s1, s2)
    out = np.empty((B, d2, oh, ow), np.float32)
    _check(lib().cvo_corr_forward(_p(x1), _p(x2), _p(out), B, C, H, W, pad, k, md, s1, s2, acc), "corr_forward")
    return out


def corr_backward(x1, x2, gout, pad, k, md, s1, s2):
    x1, x2, gout = _c(x1), _c(x2), _c(gout)
    B, C, H, W = x1.shape
    g1, g2 = np.empty_like(x1), np.empty_like(x2)
    _check(lib().cvo_corr_backward(_p(x1), _p(x2), _p(gout), _p(g1), _p(g2), B, C, H, W, pad, k, md, s1, s2),
           "corr_backward")
    return g1, g2


def corr_backward_reforder(x1, x2, gout, pad, md, s2=1):
    x1, x2, gout = _c(x1), _c(x2), _c(gout)
    B, C, H, W = x1.shape
    g1, g2 = np.empty_like(x1), np.empty_like(x2)
    _check(lib().cvo_corr_backward_reforder(_p(x1), _p(x2), _p(gout), _p(g1), _p(g2), B, C, H, W, pad, md, s2),
           "corr_backward_reforder")
    return g1, g2


def flow_warp_forward(img, flow, mode=WARP_TORCH) -> np.ndarray:
    img, flow = _c(img), _c(flow)
    B, C, H, W = img.shape
    assert flow.shape == (B, 2, H, W)
    out = np.empty_like(img)
    _check(lib().cvo_flow_warp_forward(_p(img), _p(flow), _p(out), B, C, H, W, mode), "flow_warp_forward")
    return out


def flow_warp_backward(img, flow, gout, mode=WARP_TORCH):
    img, flow, gout = _c(img), _c(flow), _c(gout)
    B, C, H, W = img.shape
    gimg, gflow = np.empty_like(img), np.empty_like(flow)
    _check(lib().cvo_flow_warp_backward(_p(img), _p(flow), _p(gout), _p(gimg), _p(gflow), B, C, H, W, mode),
           "flow_warp_backward")
    return gimg, gflow


def flow_upsample2x(coarse) -> np.ndarray:
    """F.interpolate(2 * coarse, scale_factor=2, mode='bilinear', align_corners=True) for a (B,2,Hc,Wc) flow."""
    coarse = _c(coarse)
    B, two, Hc, Wc = coarse.shape
    assert two == 2
    up = np.empty((B, 2, 2 * Hc, 2 * Wc), np.float32)
    _check(lib().cvo_flow_upsample2x(_p(coarse), _p(up), B, Hc, Wc), "flow_upsample2x")
    return up


def leaky_relu(x, slope=0.1) -> np.ndarray:
    y = _c(x).copy()
    lib().cvo_leaky_relu(_p(y), y.size, slope)
    return y


def leaky_relu_backward(y, g, slope=0.1) -> np.ndarray:
    y, g = _c(y), _c(g).copy()
    lib().cvo_leaky_relu_backward(_p(y), _p(g), g.size, slope)
    return g


def level_forward(x1, x2, flow, pad=4, k=1, md=4, s1=1, s2=1, warp_mode=WARP_TORCH, slope=0.1,
                  acc=ACC_DOUBLE) -> np.ndarray:
    """warp (if flow is not None) -> correlation -> LeakyReLU (if slope is not None)."""
    x1, x2 = _c(x1), _c(x2)
    B, C, H, W = x1.shape
    d2, oh, ow = corr_out_dims(H, W, pad, k, md, s1, s2)
    out = np.empty((B, d2, oh, ow), np.float32)
    null = ctypes.cast(None, _f32p)
    if flow is not None:
        flow = _c(flow)
        scratch = np.empty_like(x2)
        fp, sp = _p(flow), _p(scratch)
    else:
        fp, sp = null, null
    rc = lib().cvo_level_forward(_p(x1), _p(x2), fp, sp, _p(out), B, C, H, W, pad, k, md, s1, s2, warp_mode,
                                 -1.0 if slope is None else float(slope), acc)
    _check(rc, "level_forward")
    return out


def level_backward(x1, x2, flow, gout, pad=4, k=1, md=4, s1=1, s2=1, warp_mode=WARP_TORCH, slope=0.1):
    """Adjoint of level_forward: returns (g_x1, g_x2, g_flow or None)."""
    x1, x2, gout = _c(x1), _c(x2), _c(gout)
    second = flow_warp_forward(x2, flow, warp_mode) if flow is not None else x2
    if slope is not None:
        y = leaky_relu(corr_forward(x1, second, pad, k, md, s1, s2), slope)
        gout = leaky_relu_backward(y, gout, slope)
    g1, g2w = corr_backward(x1, second, gout, pad, k, md, s1, s2)
    if flow is None:
        return g1, g2w, None
    g2, gflow = flow_warp_backward(x2, flow, g2w, warp_mode)
    return g1, g2, gflow
