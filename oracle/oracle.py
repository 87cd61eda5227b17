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


# ---------------------------------------------------------------------------------------------------------------
# PWCDCNet.warp (PWCNet/PWCNet.py:159-199).  The sampling itself is torch.nn.functional.grid_sample, a third-party
# dependency of the reference, pinned at torch 1.0.1 (environment.yaml:88 `pytorch=1.0.1`, :104 `torch==1.0.1.post2`):
# bilinear / zeros and -- there being no align_corners argument before torch 1.3 -- the align_corners=True rule
# ix = ((nx + 1) / 2) * (W - 1).  Its published algorithm (aten/src/ATen/native/GridSampler.h, cuda/GridSampler.cu) is
# restated here in numpy, both unnormalisation rules (align_corners=False is what the same line computes on torch >= 1.3):
# coordinate arithmetic in float32 in the reference's operation order, value accumulation in float64.
# Pinned against torch's own CPU grid_sample, both modes, in tests/test_oracle_kat.py.
# ---------------------------------------------------------------------------------------------------------------
def _pwc_geometry(flo, H, W, align_corners=True):
    f32 = np.float32
    xs = np.arange(W, dtype=f32)[None, None, :]
    ys = np.arange(H, dtype=f32)[None, :, None]
    vx = (xs + flo[:, 0]).astype(f32)
    vy = (ys + flo[:, 1]).astype(f32)
    nx = ((f32(2.0) * vx).astype(f32) / f32(max(W - 1, 1))).astype(f32) - f32(1.0)       # PWCNet.py:178
    ny = ((f32(2.0) * vy).astype(f32) / f32(max(H - 1, 1))).astype(f32) - f32(1.0)       # :179
    if align_corners:
        ix = (((nx + f32(1.0)).astype(f32) / f32(2.0)).astype(f32) * f32(W - 1)).astype(f32)
        iy = (((ny + f32(1.0)).astype(f32) / f32(2.0)).astype(f32) * f32(H - 1)).astype(f32)
    else:
        ix = ((((nx + f32(1.0)).astype(f32) * f32(W)).astype(f32) - f32(1.0)).astype(f32) / f32(2.0)).astype(f32)
        iy = ((((ny + f32(1.0)).astype(f32) * f32(H)).astype(f32) - f32(1.0)).astype(f32) / f32(2.0)).astype(f32)
    x0f, y0f = np.floor(ix), np.floor(iy)
    x1f, y1f = (x0f + f32(1.0)).astype(f32), (y0f + f32(1.0)).astype(f32)
    wts = [((x1f - ix).astype(f32) * (y1f - iy).astype(f32)).astype(f32), ((ix - x0f).astype(f32) * (y1f - iy).astype(f32)).astype(f32),
           ((x1f - ix).astype(f32) * (iy - y0f).astype(f32)).astype(f32), ((ix - x0f).astype(f32) * (iy - y0f).astype(f32)).astype(f32)]
    x0 = np.clip(x0f, -1e9, 1e9).astype(np.int64)
    y0 = np.clip(y0f, -1e9, 1e9).astype(np.int64)
    corners = [(y0, x0), (y0, x0 + 1), (y0 + 1, x0), (y0 + 1, x0 + 1)]
    inb = [(cy >= 0) & (cy < H) & (cx >= 0) & (cx < W) for cy, cx in corners]
    m = np.zeros_like(ix, dtype=f32)
    for w_, ok in zip(wts, inb):
        m = np.where(ok, (m + w_).astype(f32), m)
    mask = np.where(m < f32(0.9999), f32(0.0), np.where(m > 0, f32(1.0), m)).astype(np.float64)      # :193-194
    tx, ty = (ix - x0f).astype(f32).astype(np.float64), (iy - y0f).astype(f32).astype(np.float64)
    return corners, inb, [w_.astype(np.float64) for w_ in wts], mask, tx, ty


def pwc_warp_forward(x, flo, align_corners=True):
    x, flo = _f32(x), _f32(flo)
    B, C, H, W = x.shape
    corners, inb, wts, mask, _, _ = _pwc_geometry(flo, H, W, align_corners)
    out = np.zeros((B, C, H, W), np.float64)
    bi = np.arange(B)[:, None, None]
    for (cy, cx), ok, w_ in zip(corners, inb, wts):
        cyc, cxc = np.clip(cy, 0, H - 1), np.clip(cx, 0, W - 1)
        for c in range(C):
            out[:, c] += np.where(ok, x[bi, c, cyc, cxc].astype(np.float64) * w_, 0.0)
    return out * mask[:, None]


def pwc_warp_backward(x, flo, gradoutput, align_corners=True):
    x, flo, g = _f32(x), _f32(flo), _f32(gradoutput).astype(np.float64)
    B, C, H, W = x.shape
    corners, inb, wts, mask, tx, ty = _pwc_geometry(flo, H, W, align_corners)
    gm = g * mask[:, None]
    gx = np.zeros((B, C, H, W), np.float64)
    bi = np.broadcast_to(np.arange(B)[:, None, None], (B, H, W))
    vals = []
    for (cy, cx), ok, w_ in zip(corners, inb, wts):
        cyc, cxc = np.clip(cy, 0, H - 1), np.clip(cx, 0, W - 1)
        v = np.zeros((B, C, H, W), np.float64)
        for c in range(C):
            np.add.at(gx[:, c], (bi[ok], cyc[ok], cxc[ok]), (gm[:, c] * w_)[ok])
            v[:, c] = np.where(ok, x[bi, c, cyc, cxc].astype(np.float64), 0.0)
        vals.append(v)
    v0, v1, v2, v3 = vals
    gix = (gm * ((v1 - v0) * (1 - ty)[:, None] + (v3 - v2) * ty[:, None])).sum(1)
    giy = (gm * ((v2 - v0) * (1 - tx)[:, None] + (v3 - v1) * tx[:, None])).sum(1)
    sx, sy = ((W - 1), (H - 1)) if align_corners else (W, H)
    gflo = np.stack([gix * (sx / max(W - 1, 1)), giy * (sy / max(H - 1, 1))], 1)
    return gx, gflo


# ---------------------------------------------------------------------------------------------------------------
# MinDepthFlowProjection (my_package/MinDepthFlowProjection/mindepthflowprojection_cuda_kernel.cu:29-312), numpy.
# The reference's forward is a racy read-compare-write (:79-84); this restates the rule it implements when the race does
# not strike -- per cell (top-left corner only), the in-range source with the largest input2 > 0 wins; among equal
# input2 the lowest pixel index (raster order) -- which is what the product's atomicMax computes.  Parity is by this
# property, not by reference fixtures.
# ---------------------------------------------------------------------------------------------------------------
def _mindepth_corners(flo, H, W):
    f32 = np.float32
    xs = np.arange(W, dtype=f32)[None, None, :]
    ys = np.arange(H, dtype=f32)[None, :, None]
    x2 = (xs + flo[:, 0]).astype(f32)
    y2 = (ys + flo[:, 1]).astype(f32)
    ok = (x2 >= 0) & (y2 >= 0) & (x2 <= f32(W - 1)) & (y2 <= f32(H - 1))
    L = np.where(ok, x2, 0).astype(np.int64)
    T = np.where(ok, y2, 0).astype(np.int64)
    return ok, L, T, np.minimum(L + 1, W - 1), np.minimum(T + 1, H - 1)


def mindepth_forward(input1, input2, fillhole=0):
    flo, dep = _f32(input1), _f32(input2)
    B, _, H, W = flo.shape
    ok, L, T, _, _ = _mindepth_corners(flo, H, W)
    out = np.zeros((B, 2, H, W), np.float64)
    cnt = np.zeros((B, 1, H, W), np.float64)
    for b in range(B):
        sel = ok[b] & (dep[b, 0] > 0)
        src = np.flatnonzero(sel.ravel())
        cell = (T[b].ravel() * W + L[b].ravel())[src]
        d = dep[b, 0].ravel()[src]
        order = np.lexsort((src, -d.astype(np.float64), cell))     # by cell, then largest depth, then lowest index
        cell_s, src_s, d_s = cell[order], src[order], d[order]
        first = np.ones(len(order), bool)
        first[1:] = cell_s[1:] != cell_s[:-1]
        wc, ws = cell_s[first], src_s[first]
        cnt[b, 0].ravel()[wc] = d_s[first]
        out[b, 0].ravel()[wc] = -flo[b, 0].ravel()[ws].astype(np.float64)
        out[b, 1].ravel()[wc] = -flo[b, 1].ravel()[ws].astype(np.float64)
    if fillhole:
        filled = out.copy()
        for b in range(B):
            c = cnt[b, 0]
            for y, x in zip(*np.nonzero(c <= 0)):
                vals = []
                for dy, dx in ((0, -1), (0, 1), (-1, 0), (1, 0)):
                    yy, xx = y + dy, x + dx
                    while 0 <= yy < H and 0 <= xx < W and c[yy, xx] == 0:
                        yy, xx = yy + dy, xx + dx
                    if 0 <= yy < H and 0 <= xx < W and c[yy, xx] > 0:
                        vals.append(out[b, :, yy, xx])
                if vals:
                    filled[b, :, y, x] = np.sum(vals, 0) / len(vals)
        out = filled
    return out, cnt


def mindepth_backward(input1, input2, count, gradoutput):
    flo, dep, cnt, g = _f32(input1), _f32(input2), _f32(count), _f32(gradoutput).astype(np.float64)
    B, _, H, W = flo.shape
    ok, L, T, R, Bm = _mindepth_corners(flo, H, W)
    gi1 = np.zeros((B, 2, H, W), np.float64)
    bi = np.arange(B)[:, None, None]
    for cy, cx in ((T, L), (T, R), (Bm, L), (Bm, R)):
        hit = ok & (dep[:, 0] == cnt[bi, 0, cy, cx])
        for ch in range(2):
            gi1[:, ch] += np.where(hit, -g[bi, ch, cy, cx], 0.0)
    return gi1, np.zeros((B, 1, H, W), np.float64)
