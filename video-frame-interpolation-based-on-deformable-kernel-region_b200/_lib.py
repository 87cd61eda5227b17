"""ctypes binding of libvfidkr_b200.so (the C ABI declared in include/vfidkr_b200.h).

There is NO fallback: if the shared library is missing or a call fails, an exception is raised.
PyTorch is used by the callers only for device memory and streams; no torch type crosses this boundary.
"""
from __future__ import annotations

import ctypes
from ctypes import c_float, c_int, c_longlong, c_void_p
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libvfidkr_b200.so"

VFIDKR_OK, VFIDKR_ERR_ARG, VFIDKR_ERR_CUDA = 0, 1, 2


class VfidkrError(RuntimeError):
    """A C-ABI call returned non-zero (the reference only prints the code, e.g.
    FilterInterpolationLayer.py:76-77; here it is an exception)."""


_P = c_void_p
_I = c_int

# name -> argtypes (restype is always int unless listed in _SPECIAL)
_SIGNATURES = {
    "vfidkr_filterinterpolation_forward_ori": [_P, _P, _P, _P, _I, _I, _I, _I, _I, _P],
    "vfidkr_filterinterpolation_forward_ori_blend": [_P, _P, _P, _P, _I, _I, _I, _I, _I, c_float, _I, c_longlong, _P],
    "vfidkr_filterinterpolation_backward_ori": [_P] * 7 + [_I] * 5 + [_P],
    "vfidkr_filterinterpolation_forward_dkr": [_P] * 5 + [_I] * 5 + [_P],
    "vfidkr_filterinterpolation_backward_dkr": [_P] * 9 + [_I] * 5 + [_P],
    "vfidkr_filterinterpolation_forward_deforconv": [_P] * 5 + [_I] * 5 + [_P],
    "vfidkr_filterinterpolation_backward_deforconv": [_P] * 9 + [_I] * 5 + [_P],
    "vfidkr_filterinterpolation_forward_nofilterwithdeforconv": [_P] * 4 + [_I] * 5 + [_P],
    "vfidkr_filterinterpolation_backward_nofilterwithdeforconv": [_P] * 7 + [_I] * 5 + [_P],
    "vfidkr_flowprojection_forward_lowres": [_P, c_float, c_float, _P, _P, _P, _I, _I, _I, _I, _P],
    "vfidkr_flow_upsample4": [_P, c_float, c_float, _P, _I, _I, _I, _P],
    "vfidkr_flow_upsample4_backward": [_P, c_float, c_float, _P, _I, _I, _I, _P],
    "vfidkr_mindepthflowprojection_forward": [_P] * 4 + [_I] * 4 + [_P],
    "vfidkr_mindepthflowprojection_backward": [_P] * 6 + [_I] * 3 + [_P],
    "vfidkr_frame_padding": [_I, ctypes.POINTER(_I)],
    "vfidkr_frames_u8_to_padded_f32": [_P, _P, _I, _I, _I, _P],
    "vfidkr_padded_f32_to_frames_u8": [_P, _P, _I, _I, _I, _P],
    "vfidkr_pwcwarp_forward": [_P] * 3 + [_I] * 5 + [_P],
    "vfidkr_pwcwarp_backward": [_P] * 5 + [_I] * 5 + [_P],
    "vfidkr_flowprojection_forward": [_P] * 3 + [_I] * 4 + [_P],
    "vfidkr_flowprojection_backward": [_P] * 4 + [_I] * 3 + [_P],
    "vfidkr_depthflowprojection_forward": [_P] * 4 + [_I] * 4 + [_P],
    "vfidkr_depthflowprojection_backward": [_P] * 7 + [_I] * 3 + [_P],
    "vfidkr_interpolation_forward": [_P] * 3 + [_I] * 5 + [_P],
    "vfidkr_interpolation_backward": [_P] * 5 + [_I] * 5 + [_P],
    "vfidkr_separableconv_forward": [_P] * 4 + [_I] * 5 + [_P],
    "vfidkr_separableconv_backward": [_P] * 7 + [_I] * 5 + [_P],
    "vfidkr_separableconvflow_forward": [_P] * 3 + [_I] * 4 + [_P],
    "vfidkr_separableconvflow_backward": [_P] * 5 + [_I] * 4 + [_P],
    "vfidkr_correlation_outshape": [_I] * 7 + [ctypes.POINTER(_I)] * 3,
    "vfidkr_correlation_forward": [_P] * 3 + [_I] * 10 + [_P],
    "vfidkr_correlation_forward_pair": [_P] * 4 + [_I] * 10 + [_P],
    "vfidkr_correlation_backward": [_P] * 5 + [_I] * 10 + [_P],
    "vfidkr_abi_version": [],
    "vfidkr_trim_scratch": [],
    "vfidkr_debug_force_forward_path": [_I],
    "vfidkr_debug_force_correlation_path": [_I],
}
EXPORTED_SYMBOLS = sorted(list(_SIGNATURES) + ["vfidkr_launch_count", "vfidkr_last_error"])

_lib = None


def load() -> ctypes.CDLL:
    """Load the library (once).  Raises ImportError with build instructions if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built. Run "
            "`python -c 'import __graft_entry__ as g; g.build()'` (or `python -m vfidkr_b200.build`). "
            "There is no CPU fallback.")
    lib = ctypes.CDLL(str(LIB_PATH))
    for name, argtypes in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = c_int
    lib.vfidkr_launch_count.argtypes = []
    lib.vfidkr_launch_count.restype = ctypes.c_ulonglong
    lib.vfidkr_last_error.argtypes = []
    lib.vfidkr_last_error.restype = ctypes.c_char_p
    _lib = lib
    return lib


def call(name: str, *args) -> None:
    """Invoke a C-ABI entry point and turn a non-zero status into an exception."""
    lib = load()
    err = getattr(lib, name)(*args)
    if err != VFIDKR_OK:
        detail = lib.vfidkr_last_error().decode(errors="replace") if err == VFIDKR_ERR_CUDA else "argument rejected"
        raise VfidkrError(f"{name} failed with status {err}: {detail}")


_PATHS = {None: 0, "auto": 0, "strip": 1, "tile": 2, "direct": 3}


def debug_force_forward_path(path) -> int:
    """TEST HOOK: pin the FilterInterpolation forward implementation (None/"auto", "strip", "tile", "direct")."""
    prev = load().vfidkr_debug_force_forward_path(_PATHS[path])
    if prev < 0:
        raise VfidkrError("vfidkr_debug_force_forward_path rejected the value")
    return prev


def debug_force_correlation_path(path) -> int:
    """TEST / MEASUREMENT HOOK: None/"auto", "simt" or "tensor" (tcgen05 kind::tf32, 3 x TF32) for the correlation forward."""
    prev = load().vfidkr_debug_force_correlation_path({None: 0, "auto": 0, "simt": 1, "tensor": 2}[path])
    if prev < 0:
        raise VfidkrError("vfidkr_debug_force_correlation_path rejected the value")
    return prev


def trim_scratch() -> None:
    """Give the scratch memory cached by the library's private pool back to the device."""
    call("vfidkr_trim_scratch")


def launch_count() -> int:
    return int(load().vfidkr_launch_count())


def abi_version() -> int:
    return int(load().vfidkr_abi_version())
