"""Make unmodified CerberusNet model code run on this package.

The reference models bind the hot path by module path and by name:

    from nnet_training.correlation_package.correlation import Correlation      (pwcnet.py:8, ...)
    from nnet_training.loss_functions.UnFlowLoss import flow_warp              (pwcnet.py:7, ...)

``install()`` registers :mod:`cerberusnet_b200.correlation` in ``sys.modules`` under the
reference's dotted name *before* the models are imported (the reference module itself cannot be
imported: its line 2 loads a py3.8 .so by a cwd-relative path), and ``patch_flow_warp()`` swaps
the ``flow_warp`` name inside already-imported model modules -- the reference's loss module
(``nnet_training.loss_functions.UnFlowLoss``, whose ``unFlowLoss.forward`` warps the image pyramid at four scales in
both directions, UnFlowLoss.py:279-283) included.  The fused forms (``WarpCorrelation``, ``decoder.FlowDecoder``,
``photometric.photometric_loss``) are opt-in replacements for callers that adopt them.
"""
from __future__ import annotations

import sys
import types
from typing import Iterable

import torch

from . import correlation as _correlation
from .flow_warp import flow_warp as _flow_warp_fn   # (the package re-exports the function under the sub-module's name)

REFERENCE_CORRELATION_MODULE = "nnet_training.correlation_package.correlation"
REFERENCE_MODEL_MODULES = (
    "nnet_training.nnet_models.pwcnet",
    "nnet_training.nnet_models.pwcnet_sfd",
    "nnet_training.nnet_models.ocrnet_sfd",
    "nnet_training.nnet_models.detr_sfd",
)
# the loss side binds flow_warp inside its own module (UnFlowLoss.py:83, used at :279-283): patching the module's
# global redirects every call unFlowLoss.forward makes
REFERENCE_LOSS_MODULES = ("nnet_training.loss_functions.UnFlowLoss",)


def install(register_cerberus_ops: bool = True) -> types.ModuleType:
    """Register the drop-in module under the reference's import path.  Idempotent."""
    parts = REFERENCE_CORRELATION_MODULE.split(".")
    for i in range(1, len(parts)):
        name = ".".join(parts[:i])
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                pkg = types.ModuleType(name)
                pkg.__path__ = []  # namespace-like placeholder
                sys.modules[name] = pkg
    sys.modules[REFERENCE_CORRELATION_MODULE] = _correlation
    parent = sys.modules.get(".".join(parts[:-1]))
    if parent is not None:
        setattr(parent, parts[-1], _correlation)
    if register_cerberus_ops:
        _alias_cerberus_namespace()
    return _correlation


def _alias_cerberus_namespace():
    """If the reference's own op library is not loaded, answer ``torch.ops.cerberus.correlation``
    / ``correlation_backward`` (the schema of correlation_cuda.cpp:45-48) with our kernels, so
    code that calls the raw op (correlation.py:78-80, the ONNX symbolic onnx_export.py:21) works."""
    try:
        torch.ops.cerberus.correlation  # noqa: B018  (raises if undefined)
        return
    except (AttributeError, RuntimeError):
        pass
    lib = torch.library.Library("cerberus", "DEF")
    lib.define("correlation(Tensor input1, Tensor input2, int pad_size, int kernel_size, int max_displacement, "
               "int stride1, int stride2, int corr_type_multiply) -> Tensor")
    lib.define("correlation_backward(Tensor input1, Tensor input2, Tensor gradOutput, int pad_size, int kernel_size, "
               "int max_displacement, int stride1, int stride2, int corr_type_multiply) -> Tensor[]")
    lib.impl("correlation", lambda *a: torch.ops.cerberus_b200.correlation(*a), "CUDA")
    lib.impl("correlation_backward", lambda *a: torch.ops.cerberus_b200.correlation_backward(*a), "CUDA")
    install._cerberus_lib = lib  # keep the registration alive


def patch_flow_warp(modules: Iterable[str] = REFERENCE_MODEL_MODULES + REFERENCE_LOSS_MODULES) -> int:
    """Point the name ``flow_warp`` of every already-imported reference model AND loss module at the CUDA
    kernel (differentiable with respect to image and flow).  Returns how many modules were patched."""
    n = 0
    for name in modules:
        mod = sys.modules.get(name)
        if mod is not None and hasattr(mod, "flow_warp"):
            mod.flow_warp = _flow_warp_fn
            n += 1
    return n
