"""Drop-in aliases: make the reference's own import lines resolve to this package.

The reference's networks import the operators as (networks/DAIN.py:11-13, PWCNet/PWCNet.py:15)
    from my_package.FilterInterpolation import FilterInterpolationModule
    from my_package.FlowProjection import FlowProjectionModule
    from my_package.DepthFlowProjection import DepthFlowProjectionModule
    from PWCNet.correlation_package_pytorch1_0.correlation import Correlation
and, in the same files, everything else from the reference tree itself:
    from PWCNet.PWCNet import conv, deconv            (networks/DAIN.py:9-10)
    import PWCNet ... PWCNet.__dict__['pwc_dc_net']    (networks/DAIN.py:17,63-65; PWCNet/__init__.py:1)
`install_reference_aliases()` therefore puts ONE finder in front of `sys.meta_path` that answers exactly the
operator module names -- `my_package`, `my_package.<Op>[.<Op>Module|.<Op>Layer]`,
`PWCNet.correlation_package_pytorch1_0[.correlation]` -- and nothing else: `PWCNet`, `PWCNet.PWCNet`, `networks`,
`S2D_models`, ... keep coming from the reference tree on `sys.path`, so its unmodified networks run on the B200
kernels.  Only if no `PWCNet` package can be found anywhere does the finder supply an empty stand-in, so that
the import line of PWCNet/PWCNet.py:15 still resolves for callers that vendor nothing but that line.
Nothing is registered implicitly; `remove_reference_aliases()` undoes it.
"""
from __future__ import annotations

import importlib
import importlib.abc
import importlib.machinery
import sys

_PKG = __name__.rsplit(".", 1)[0]

# reference sub-package of my_package -> (implementing module here, [classes it exports])
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
_CORR_PKG = "PWCNet.correlation_package_pytorch1_0"


def _table() -> dict:
    """module name -> (is_package, implementing module here or None, [exported names])."""
    t = {"my_package": (True, None, [])}
    for sub, (impl, names) in _MY_PACKAGE.items():
        t[f"my_package.{sub}"] = (True, impl, names)          # `from my_package.X import XModule`
        for n in names:                                        # `from my_package.X.XModule import XModule`
            t[f"my_package.{sub}.{n}"] = (False, impl, names)
    t[_CORR_PKG] = (True, None, [])
    t[_CORR_PKG + ".correlation"] = (False, "correlation", ["Correlation", "CorrelationFunction"])
    return t


class _OperatorAliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Answers the operator module names above; every other name falls through to the normal finders."""

    def __init__(self):
        self.table = _table()

    # -- finder
    def find_spec(self, fullname, path=None, target=None):
        if fullname in self.table:
            return importlib.machinery.ModuleSpec(fullname, self, is_package=self.table[fullname][0])
        if fullname == "PWCNet":
            # the reference's real package wins whenever it is importable; an empty stand-in otherwise
            for finder in sys.meta_path:
                if finder is self or not hasattr(finder, "find_spec"):
                    continue
                try:
                    spec = finder.find_spec(fullname, path, target)
                except Exception:
                    spec = None
                if spec is not None:
                    return spec
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    # -- loader
    def create_module(self, spec):
        return None

    def exec_module(self, module):
        name = module.__name__
        module.__vfidkr_b200_alias__ = True
        if name == "PWCNet":
            module.__path__ = []
            return
        is_pkg, impl, names = self.table[name]
        if is_pkg:
            module.__path__ = []
        if impl is not None:
            src = importlib.import_module(f"{_PKG}.{impl}")
            for n in names:
                setattr(module, n, getattr(src, n))
            module.__all__ = list(names)


def _installed():
    return [f for f in sys.meta_path if isinstance(f, _OperatorAliasFinder)]


def _is_operator_module(name: str, table) -> bool:
    return name in table or name.startswith("my_package.")


def install_reference_aliases(overwrite: bool = False) -> list[str]:
    """Route the reference's operator imports (`my_package.*`, `PWCNet.correlation_package_pytorch1_0.correlation`)
    to this package.  With `overwrite`, modules of those names that are already imported (the reference's own
    Python layers, say) are dropped from `sys.modules` so that the next import resolves here; the reference's
    `PWCNet` package itself is never touched.  Returns the module names this call made resolvable."""
    finders = _installed()
    finder = finders[0] if finders else _OperatorAliasFinder()
    if not finders:
        sys.meta_path.insert(0, finder)
    done = []
    for name in finder.table:
        loaded = sys.modules.get(name)
        if loaded is not None and not getattr(loaded, "__vfidkr_b200_alias__", False):
            if not overwrite:
                continue
            for k in [k for k in sys.modules if k == name or k.startswith(name + ".")]:
                if not getattr(sys.modules[k], "__vfidkr_b200_alias__", False):
                    del sys.modules[k]
            parent, _, leaf = name.rpartition(".")
            if parent in sys.modules and hasattr(sys.modules[parent], leaf):
                delattr(sys.modules[parent], leaf)
        done.append(name)
    importlib.invalidate_caches()
    return done


def remove_reference_aliases() -> None:
    """Take the finder out again and forget the alias modules it created."""
    for f in _installed():
        sys.meta_path.remove(f)
    for k in [k for k, m in list(sys.modules.items()) if getattr(m, "__vfidkr_b200_alias__", False)]:
        del sys.modules[k]
    importlib.invalidate_caches()
