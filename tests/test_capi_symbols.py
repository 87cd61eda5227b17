"""The C-ABI library loads without a GPU and exports every symbol include/vfidkr_b200.h declares.
No compute entry point is called here."""
import ctypes
import re
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    text = (ROOT / "include" / "vfidkr_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vfidkr_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_families():
    names = _declared()
    for needle in ("filterinterpolation_forward_ori", "filterinterpolation_backward_dkr",
                   "filterinterpolation_forward_deforconv", "filterinterpolation_backward_nofilterwithdeforconv",
                   "flowprojection_forward", "depthflowprojection_backward", "interpolation_forward",
                   "separableconv_backward", "separableconvflow_forward", "correlation_forward",
                   "correlation_backward"):
        assert any(needle in n for n in names), needle


def test_library_exports_every_declared_symbol(lib):
    dll = ctypes.CDLL(str(lib._lib.LIB_PATH))
    missing = [n for n in _declared() if not hasattr(dll, n)]
    assert not missing, f"declared in the header but not exported: {missing}"


def test_python_binding_covers_every_declared_symbol(lib):
    assert sorted(lib._lib.EXPORTED_SYMBOLS) == _declared()


def test_info_entry_points(lib):
    assert lib.abi_version() == 100
    assert lib.launch_count() >= 0
    oc, oh, ow = lib.correlation_output_shape(64, 96, 4, 1, 4, 1, 1)   # pure host arithmetic
    assert (oc, oh, ow) == (81, 64, 96)
    assert lib.correlation_output_shape(64, 96, 3, 3, 20, 1, 2) == (441, 28, 60)


def test_only_sm100a_code_is_embedded(lib):
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        import pytest
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", str(lib._lib.LIB_PATH)], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
