"""Drop-in aliases: make the reference's own import lines resolve to this package.

The reference's networks import the operators as (networks/DAIN.py:11-13, PWCNet/PWCNet.py:15)
    from my_package.FilterInterpolation import FilterInterpolationModule
    from my_package.FlowProjection import FlowProjectionModule
    from my_package.DepthFlowProjection import DepthFlowProjectionModule
    from PWCNet.correlation_package_pytorch1_0.correlation import Correlation
`install_reference_aliases()` registers modules of those names in sys.modules that export this package's
classes, so an unmodified caller picks up the B200 kernels.  Nothing is registered implicitly.
"""
from __future__ import annotations

import sys
import types

_MY_PACKAGE = {
    "FilterInterpolation": ("filter_interpolation", ["FilterInterpolationModule", "FilterInterpolationLayer"]),
    "FlowProjection": ("flow_projection", ["FlowProjectionModule", "FlowProjectionLayer"]),
    "DepthFlowProjection": ("flow_projection", ["DepthFlowProjectionModule", "DepthFlowProjectionLayer"]),
    "MinDepthFlowProjection": ("flow_projection", ["minDepthFlowProjectionModule", "minDepthFlowProjectionLayer"]),
    "Interpolation": ("interpolation", ["InterpolationModule", "InterpolationLayer"]),
    "InterpolationCh": ("interpolation", ["InterpolationChModule", "InterpolationChLayer"]),
    "SeparableConv": ("separable_conv", ["SeparableConvModule", "SeparableConvLayer"]),
    "SeparableConvFlow": ("separable_conv", ["SeparableConvFlowModule", "SeparableConvFlowLayer"]),
}


def install_reference_aliases(overwrite: bool = False) -> list[str]:
    """Register `my_package.*` and `PWCNet.correlation_package_pytorch1_0.correlation` aliases.
    Returns the list of module names that were registered."""
    import importlib

    pkg = __name__.rsplit(".", 1)[0]
    done = []

    def register(name, module):
        if name in sys.modules and not overwrite:
            return
        sys.modules[name] = module
        done.append(name)

    root = types.ModuleType("my_package")
    root.__path__ = []   # mark as package
    register("my_package", root)
    for sub, (impl, names) in _MY_PACKAGE.items():
        src = importlib.import_module(f"{pkg}.{impl}")
        m = types.ModuleType(f"my_package.{sub}")
        for n in names:
            setattr(m, n, getattr(src, n))
        m.__all__ = list(names)
        register(f"my_package.{sub}", m)
        setattr(sys.modules["my_package"], sub, m)

    corr = importlib.import_module(f"{pkg}.correlation")
    if "PWCNet" not in sys.modules or overwrite:
        p = types.ModuleType("PWCNet")
        p.__path__ = []
        register("PWCNet", p)
    cp = types.ModuleType("PWCNet.correlation_package_pytorch1_0")
    cp.__path__ = []
    register("PWCNet.correlation_package_pytorch1_0", cp)
    cm = types.ModuleType("PWCNet.correlation_package_pytorch1_0.correlation")
    cm.Correlation, cm.CorrelationFunction = corr.Correlation, corr.CorrelationFunction
    register("PWCNet.correlation_package_pytorch1_0.correlation", cm)
    return done
