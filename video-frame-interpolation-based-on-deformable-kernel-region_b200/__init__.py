"""vfidkr_b200 -- B200-native (sm_100a) per-pixel sampling / warping operators of VFIDKR.

The directory name carries the full reference name; import it as `vfidkr_b200` (a shim package at the
repository root maps that name onto this directory).  Everything here is a thin Python front-end over the
C ABI of libvfidkr_b200.so (include/vfidkr_b200.h); there is no CPU or framework fallback.
"""
from . import _lib
from ._lib import (VfidkrError, abi_version, debug_force_correlation_path, debug_force_forward_path, launch_count,
                   trim_scratch)
from .correlation import Correlation, CorrelationFunction, correlation_output_shape, correlation_pair
from .filter_interpolation import (FilterInterpolationBlendLayer, filter_interpolate_blend, filter_interpolate_into,
                                   FilterInterpolationLayer, FilterInterpolationLayerDeforConv,
                                   FilterInterpolationLayerDKR, FilterInterpolationLayerNoFilterWithDeforConv,
                                   FilterInterpolationModule)
from .flow_projection import (DepthFlowProjectionLayer, DepthFlowProjectionModule, FlowProjectionLayer,
                              FlowProjectionModule, minDepthFlowProjectionLayer, minDepthFlowProjectionModule,
                              flow_project_lowres, flow_upsample4)
from .interpolation import InterpolationChLayer, InterpolationChModule, InterpolationLayer, InterpolationModule
from .separable_conv import (SeparableConvFlowLayer, SeparableConvFlowModule, SeparableConvLayer,
                             SeparableConvModule)
from .compat import install_reference_aliases
from .host_stream import PairStream, bind_to_gpu_numa_node
from .pwc_warp import PWCWarpLayer, pwc_warp
from .frame_io import frame_padding, frames_to_padded, padded_to_frames

__version__ = "0.1.0"
