"""ctypes front-end of the CPU oracle (oracle/vfidkr_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.

Every function takes float32 numpy arrays (NCHW, C-contiguous) and returns float64 arrays.
Parity status: unpinned by reference fixtures (the reference has none); pinned by the
known-answer identities in tests/test_oracle_kat.py.  See the header of vfidkr_oracle.c.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "libvfidkr_oracle.so"

FI_ORI, FI_DKR, FI_DEFORCONV, FI_NOFILTER = 0, 1, 2, 3
FI_VARIANTS = {"ori": FI_ORI, "dkr": FI_DKR, "deforconv": FI_DEFORCONV, "nofilterwithdeforconv": FI_NOFILTER}


def build(force: bool = False) -> Path:
    """Compile the C restatement with the committed Makefile (gcc, OpenMP)."""
    src = _HERE / "vfidkr_oracle.c"
    if force or not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE)] + (["-B"] if force else []), check=True,
                       stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(str(_SO))
    return _lib


def _f32(a) -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(a), dtype=np.float32)
    return a


def _p(a):
    if a is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(a.ctypes.data)


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def set_num_threads(n: int) -> None:
    """Must be called before the first oracle call of the process to take effect reliably."""
    os.environ["OMP_NUM_THREADS"] = str(int(n))


# ---------------------------------------------------------------- FilterInterpolation
def fi_forward(variant, input1, input2, input3, input4=None):
    v = FI_VARIANTS[variant] if isinstance(variant, str) else int(variant)
    i1, i2, i3 = _f32(input1), _f32(input2), _f32(input3)
    i4 = _f32(input4) if input4 is not None else None
    B, C, H, W = i1.shape
    T2 = i3.shape[1] // 2 if v == FI_NOFILTER else i3.shape[1]
    F = int(np.sqrt(float(T2)))  # filterinterpolation_cuda.cc:556-557, :395
    assert i2.shape == (B, 2, H, W)
    out = np.empty((B, C, H, W), dtype=np.float64)
    err = lib().oracle_fi_forward(v, _p(i1), _p(i2), _p(i3), _p(i4), _p(out), B, C, H, W, F)
    if err:
        raise RuntimeError(f"oracle_fi_forward returned {err}")
    return out


def fi_backward(variant, input1, input2, input3, input4, gradoutput):
    v = FI_VARIANTS[variant] if isinstance(variant, str) else int(variant)
    i1, i2, i3, g = _f32(input1), _f32(input2), _f32(input3), _f32(gradoutput)
    i4 = _f32(input4) if input4 is not None else None
    B, C, H, W = i1.shape
    T2 = i3.shape[1] // 2 if v == FI_NOFILTER else i3.shape[1]
    F = int(np.sqrt(float(T2)))
    gi1 = np.empty(i1.shape, np.float64)
    gi2 = np.empty(i2.shape, np.float64)
    gi3 = np.empty(i3.shape, np.float64)
    gi4 = np.empty(i4.shape, np.float64) if i4 is not None else None
    err = lib().oracle_fi_backward(v, _p(i1), _p(i2), _p(i3), _p(i4), _p(g), _p(gi1), _p(gi2), _p(gi3),
                                   _p(gi4), B, C, H, W, F)
    if err:
        raise RuntimeError(f"oracle_fi_backward returned {err}")
    return gi1, gi2, gi3, gi4


# ---------------------------------------------------------------- (Depth)FlowProjection
def flowprojection_forward(input1, depth=None, fillhole=0):
    i1 = _f32(input1)
    d = _f32(depth) if depth is not None else None
    B, two, H, W = i1.shape
    assert two == 2
    count = np.empty((B, 1, H, W), np.float64)
    out = np.empty((B, 2, H, W), np.float64)
    lib().oracle_flowprojection_forward(_p(i1), _p(d), _p(count), _p(out), B, H, W, int(fillhole))
    return out, count


def flowprojection_backward(input1, depth, count, output, gradoutput):
    i1, g, cn = _f32(input1), _f32(gradoutput), _f32(count)
    d = _f32(depth) if depth is not None else None
    o = _f32(output) if output is not None else None
    B, _, H, W = i1.shape
    gi1 = np.empty(i1.shape, np.float64)
    gi2 = np.empty((B, 1, H, W), np.float64) if d is not None else None
    lib().oracle_flowprojection_backward(_p(i1), _p(d), _p(cn), _p(o), _p(g), _p(gi1), _p(gi2), B, H, W)
    return gi1, gi2


# ---------------------------------------------------------------- Interpolation(Ch)
def interpolation_forward(input1, input2):
    i1, i2 = _f32(input1), _f32(input2)
    B, C, H, W = i1.shape
    out = np.empty(i1.shape, np.float64)
    lib().oracle_interpolation_forward(_p(i1), _p(i2), _p(out), B, C, H, W)
    return out


def interpolation_backward(input1, input2, gradoutput):
    i1, i2, g = _f32(input1), _f32(input2), _f32(gradoutput)
    B, C, H, W = i1.shape
    gi1 = np.empty(i1.shape, np.float64)
    gi2 = np.empty(i2.shape, np.float64)
    lib().oracle_interpolation_backward(_p(i1), _p(i2), _p(g), _p(gi1), _p(gi2), B, C, H, W)
    return gi1, gi2


# ---------------------------------------------------------------- SeparableConv(Flow)
def sepconv_forward(input1, input2, input3):
    i1, i2, i3 = _f32(input1), _f32(input2), _f32(input3)
    B, C, H, W = i1.shape
    F = i2.shape[1]
    out = np.empty((B, C, H - F + 1, W - F + 1), np.float64)
    err = lib().oracle_sepconv_forward(_p(i1), _p(i2), _p(i3), _p(out), B, C, H, W, F)
    if err:
        raise RuntimeError("oracle_sepconv_forward: bad shape")
    return out


def sepconv_backward(input1, input2, input3, gradoutput):
    i1, i2, i3, g = _f32(input1), _f32(input2), _f32(input3), _f32(gradoutput)
    B, C, H, W = i1.shape
    F = i2.shape[1]
    gi1, gi2, gi3 = (np.empty(a.shape, np.float64) for a in (i1, i2, i3))
    err = lib().oracle_sepconv_backward(_p(i1), _p(i2), _p(i3), _p(g), _p(gi1), _p(gi2), _p(gi3), B, C, H, W, F)
    if err:
        raise RuntimeError("oracle_sepconv_backward: bad shape")
    return gi1, gi2, gi3


def sepconvflow_forward(input2, input3):
    i2, i3 = _f32(input2), _f32(input3)
    B, F, Ho, Wo = i2.shape
    flow = np.empty((B, 2, Ho, Wo), np.float64)
    lib().oracle_sepconvflow_forward(_p(i2), _p(i3), _p(flow), B, Ho, Wo, F)
    return flow


def sepconvflow_backward(input2, input3, gradflow):
    i2, i3, g = _f32(input2), _f32(input3), _f32(gradflow)
    B, F, Ho, Wo = i2.shape
    gi2, gi3 = np.empty(i2.shape, np.float64), np.empty(i3.shape, np.float64)
    lib().oracle_sepconvflow_backward(_p(i2), _p(i3), _p(g), _p(gi2), _p(gi3), B, Ho, Wo, F)
    return gi2, gi3


# ---------------------------------------------------------------- Correlation
def correlation_outshape(H, W, pad, k, md, s1, s2):
    oc, oh, ow = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    lib().oracle_correlation_outshape(H, W, pad, k, md, s1, s2, ctypes.byref(oc), ctypes.byref(oh), ctypes.byref(ow))
    return oc.value, oh.value, ow.value


def correlation_forward(input1, input2, pad=4, k=1, md=4, s1=1, s2=1):
    i1, i2 = _f32(input1), _f32(input2)
    B, C, H, W = i1.shape
    oc, oh, ow = correlation_outshape(H, W, pad, k, md, s1, s2)
    out = np.empty((B, oc, oh, ow), np.float64)
    err = lib().oracle_correlation_forward(_p(i1), _p(i2), _p(out), B, C, H, W, pad, k, md, s1, s2)
    if err:
        raise RuntimeError("oracle_correlation_forward: empty output")
    return out


def correlation_backward(input1, input2, gradoutput, pad=4, k=1, md=4, s1=1, s2=1):
    i1, i2, g = _f32(input1), _f32(input2), _f32(gradoutput)
    B, C, H, W = i1.shape
    gi1, gi2 = np.empty(i1.shape, np.float64), np.empty(i2.shape, np.float64)
    err = lib().oracle_correlation_backward(_p(i1), _p(i2), _p(g), _p(gi1), _p(gi2), B, C, H, W, pad, k, md, s1, s2)
    if err:
        raise RuntimeError("oracle_correlation_backward: empty output")
    return gi1, gi2
