"""cerberusnet_b200 -- B200-native (sm_100a) cost-volume hot path of CerberusNet.

Correlation (forward + backward) fused with the bilinear flow warp of the second feature map and
the LeakyReLU(0.1) that follows it, behind the reference's own Python surfaces
(``Correlation`` / ``CorrelationFunction`` / ``flow_warp``) and a C ABI
(``include/cerberus_costvolume.h``) shaped for the TensorRT correlation plugin's ``enqueue``.
"""
from ._lib import WARP_TORCH, WARP_TORCH_CPU, WARP_TRT, CostVolumeError, lib  # noqa: F401
from .correlation import (Correlation, CorrelationFunction, CorrelationTorch, UpflowWarpCorrelationFunction,  # noqa: F401
                          WarpCorrelation, WarpCorrelationFunction, warp_correlation, warp_correlation_upflow)
from .flow_warp import FlowWarpFunction, flow_warp, grid_sample, mesh_grid, norm_grid  # noqa: F401
from .photometric import PhotometricLossFunction, photometric_loss  # noqa: F401
from .install import install, patch_flow_warp  # noqa: F401
from .host_pipeline import HostPipeline  # noqa: F401
from . import ops  # noqa: F401

__version__ = "0.1.0"
