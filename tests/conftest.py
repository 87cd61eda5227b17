import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    # GPU tests are selected with `-m gpu`; if someone runs the whole suite on a CPU-only box, skip them
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def lib():
    """The product library, built if missing (nvcc cross-compiles without a GPU)."""
    import vfidkr_b200.build as B
    B.build_library()
    import vfidkr_b200
    return vfidkr_b200


@pytest.fixture(autouse=True)
def _forward_path_is_automatic_again():
    """A test that pins a FilterInterpolation forward implementation must not leak the setting into the next one."""
    yield
    mod = sys.modules.get("vfidkr_b200")
    if mod is not None and getattr(mod._lib, "_lib", None) is not None:
        mod.debug_force_forward_path(None)
        mod.debug_force_correlation_path(None)


def pytest_terminal_summary(terminalreporter):
    """Which comparisons needed the absolute floor of the per-element criterion (tests/util.py)."""
    try:
        import util as U
    except Exception:
        return
    if U.NEEDS_FLOOR:
        terminalreporter.write_line(f"[parity] {len(U.NEEDS_FLOOR)} comparisons hold |d| <= rtol*|ref| + floor only thanks to the "
                                    "absolute floor (elements tiny against the tensor's scale); worst pure-relative factors:")
        for k, v in sorted(U.NEEDS_FLOOR.items(), key=lambda kv: -kv[1])[:12]:
            terminalreporter.write_line(f"[parity]   {k}: x{v:.1f}")
