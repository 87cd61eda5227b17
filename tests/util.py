"""Seeded synthetic inputs and the tolerance rule used by the parity tests."""
from __future__ import annotations

import numpy as np

# Tolerances stated by BASELINE.json north_star:
RTOL_FWD = 1e-5      # fp32 forward outputs vs the float64 oracle
RTOL_ATOMIC = 1e-4   # atomically accumulated results (scatter gradients, projection sums)


def rng(seed):
    return np.random.default_rng(seed)


def image(r, B, C, H, W, kind="uniform"):
    if kind == "uniform":   # test_module.py:917-919
        return r.random((B, C, H, W), dtype=np.float32)
    return r.standard_normal((B, C, H, W)).astype(np.float32)


def flow(r, B, H, W, kind="gauss"):
    """kind: gauss = N(0,4^2) clipped to +-20 px; unit = U(-1,1) (test_module.py:1018);
    stress = gauss with 2 % of the pixels thrown far out of range; smooth = low-frequency field."""
    if kind == "unit":
        f = r.random((B, 2, H, W), dtype=np.float32) * 2 - 1
    elif kind == "smooth":
        yy, xx = np.meshgrid(np.linspace(0, 3, H), np.linspace(0, 3, W), indexing="ij")
        f = np.stack([4 * np.sin(xx + 0.3) + 2 * np.cos(yy), 3 * np.cos(xx * 0.7) - 2 * np.sin(yy + 1)], 0)
        f = np.broadcast_to(f[None], (B, 2, H, W)).astype(np.float32) + \
            0.05 * r.standard_normal((B, 2, H, W)).astype(np.float32)
    else:
        f = np.clip(r.standard_normal((B, 2, H, W)) * 4.0, -20, 20).astype(np.float32)
        if kind == "stress":
            m = r.random((B, 1, H, W)) < 0.02
            f = np.where(m, f + np.float32(3 * max(H, W)) * np.sign(f + 1e-3), f).astype(np.float32)
            # exact-integer and border landings
            f[..., 0, :] = 0.0
            f[:, 0, :, -1] = 0.0
            f[:, 1, -1, :] = 0.0
    return np.ascontiguousarray(f, dtype=np.float32)


def filt(r, B, F, H, W, kind="softmax"):
    if kind == "uniform":
        return r.random((B, F * F, H, W), dtype=np.float32)
    z = r.standard_normal((B, F * F, H, W))
    e = np.exp(z - z.max(axis=1, keepdims=True))
    return (e / e.sum(axis=1, keepdims=True)).astype(np.float32)


def offsets(r, B, F, H, W, amp=0.45):
    """DKR offset field; amp 0.45 keeps every deformed tap inside the in-contract domain for interior taps."""
    return ((r.random((B, 2 * F * F, H, W), dtype=np.float32) * 2 - 1) * np.float32(amp)).astype(np.float32)


def depth_inv(r, B, H, W):
    return (0.1 + 0.9 * r.random((B, 1, H, W), dtype=np.float32)).astype(np.float32)   # test_module.py:1019


def max_err(got, ref):
    """max over elements of |got-ref| / (|ref| + scale), scale = max|ref| (so exact zeros are judged
    against the magnitude of the tensor, not against 0)."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    if ref.size == 0:
        return 0.0
    assert np.isfinite(got).all(), "non-finite values in result"
    scale = float(np.abs(ref).max())
    if scale == 0.0:
        return float(np.abs(got).max())
    return float((np.abs(got - ref) / (np.abs(ref) + scale)).max())


# The second, per-element criterion: |got - ref| <= rtol * |ref| + FLOOR * rtol * max|ref|.
# With FLOOR = 0.1 this is north_star's "1e-5 relative" for every element that is not tiny compared with the
# tensor's own scale (there the absolute floor 1e-6 * max|ref| -- 1e-5 * max|ref| for atomically accumulated
# tensors -- takes over: an fp32 sum of O(1) terms that cancels to 1e-4 cannot be relatively exact).
FLOOR = 0.1


def rel_err(got, ref, rtol):
    """max over elements of |got-ref| / (rtol*|ref| + FLOOR*rtol*max|ref|): <= 1 means the per-element criterion holds."""
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    if ref.size == 0:
        return 0.0
    scale = float(np.abs(ref).max())
    if scale == 0.0:
        return 0.0 if not np.abs(got).max() else float("inf")
    return float((np.abs(got - ref) / (rtol * np.abs(ref) + FLOOR * rtol * scale)).max())


NEEDS_FLOOR = {}     # what -> worst ratio of an element that passed only thanks to the absolute floor (reported by conftest)


def assert_close(got, ref, rtol, what=""):
    """Both criteria: the normalised error of max_err (against |ref| + max|ref|) AND the per-element relative one."""
    e = max_err(got, ref)
    assert e <= rtol, f"{what}: normalised max error {e:.3e} > {rtol:.1e}"
    r = rel_err(got, ref, rtol)
    assert r <= 1.0, f"{what}: per-element criterion |d| <= {rtol:.0e}*|ref| + {FLOOR * rtol:.0e}*max|ref| violated by a factor {r:.2f}"
    # bookkeeping: would the purely relative rule (no floor) have held?
    g64, r64 = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    nz = np.abs(r64) > 0
    if nz.any():
        pure = float((np.abs(g64 - r64)[nz] / (rtol * np.abs(r64)[nz])).max())
        if pure > 1.0:
            NEEDS_FLOOR[what or "?"] = max(NEEDS_FLOOR.get(what or "?", 0.0), pure)
    return e
