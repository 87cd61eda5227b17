"""Pins the CPU oracle: known-answer identities derived from the reference kernels (SURVEY.md section 8c,
items 1-9), an independent pure-Python restatement on tiny cases, and finite-difference gradient checks of
the float64 forward.  The reference ships no fixtures for this path, so these are the oracle's anchors."""
import numpy as np
import pytest

import util as U


# ----------------------------------------------------------------------------- pure-Python second opinion
def _py_fi_ori(I, flow, filt):
    """Independent restatement of filterinterpolation_cuda_kernel.cu:2726-2820, written quadrant by
    quadrant exactly like the CUDA text (the C oracle loops taps once and classifies them instead)."""
    B, C, H, W = I.shape
    F = int(np.sqrt(filt.shape[1]))
    out = np.zeros(I.shape, np.float64)
    f32 = np.float32
    for b in range(B):
        for h in range(H):
            for w in range(W):
                fx, fy = flow[b, 0, h, w], flow[b, 1, h, w]
                x2, y2 = f32(f32(w) + fx), f32(f32(h) + fy)
                ok = (x2 >= 0 and y2 >= 0 and x2 <= f32(W - 1) and y2 <= f32(H - 1)
                      and abs(fx) < f32(W) / f32(2) and abs(fy) < f32(H) / f32(2))
                if not ok:
                    out[b, :, h, w] = I[b, :, h, w]
                    continue
                ix, iy = int(x2), int(y2)
                L, T = ix + 1 - F // 2, iy + 1 - F // 2
                R, Bm = L + F, T + F
                a, be = float(f32(x2 - f32(ix))), float(f32(y2 - f32(iy)))

                def tap(c, j, i):
                    return float(I[b, c, min(max(0, j), H - 1), min(max(0, i), W - 1)]) * \
                        float(filt[b, (j - T) * F + (i - L), h, w])
                for c in range(C):
                    TL = sum(tap(c, j, i) for j in range(T, iy + 1) for i in range(L, ix + 1))
                    TR = sum(tap(c, j, i) for j in range(T, iy + 1) for i in range(ix + 1, R))
                    BL = sum(tap(c, j, i) for j in range(iy + 1, Bm) for i in range(L, ix + 1))
                    BR = sum(tap(c, j, i) for j in range(iy + 1, Bm) for i in range(ix + 1, R))
                    out[b, c, h, w] = (1 - a) * (1 - be) * TL + a * (1 - be) * TR + (1 - a) * be * BL + a * be * BR
    return out


@pytest.mark.parametrize("F", [2, 4, 5, 6])
def test_fi_ori_matches_pure_python(oracle, F):
    r = U.rng(10 + F)
    B, C, H, W = 1, 2, 9, 11
    I = U.image(r, B, C, H, W)
    fl = U.flow(r, B, H, W, "stress")
    ft = U.filt(r, B, F, H, W, "uniform")
    got = oracle.fi_forward("ori", I, fl, ft)
    ref = _py_fi_ori(I, fl, ft)
    assert np.abs(got - ref).max() < 1e-12


# ----------------------------------------------------------------------------- KAT 1-4: FilterInterpolation
def test_kat1_identity_filter(oracle):
    r = U.rng(1)
    I = U.image(r, 2, 3, 17, 23)
    fl = np.zeros((2, 2, 17, 23), np.float32)
    ft = np.zeros((2, 16, 17, 23), np.float32)
    ft[:, 5] = 1     # tap (row 1, col 1) is the pixel itself at zero flow
    assert np.array_equal(oracle.fi_forward("ori", I, fl, ft), I.astype(np.float64))


def test_kat2_bilinear_filter_equals_interpolation(oracle):
    r = U.rng(2)
    B, C, H, W = 2, 3, 20, 24
    I, fl = U.image(r, B, C, H, W), U.flow(r, B, H, W, "gauss")
    ft = np.zeros((B, 16, H, W), np.float32)
    ft[:, [5, 6, 9, 10]] = 1
    a, b = oracle.fi_forward("ori", I, fl, ft), oracle.interpolation_forward(I, fl)
    x2 = np.arange(W, dtype=np.float32)[None, None, :] + fl[:, 0]
    y2 = np.arange(H, dtype=np.float32)[None, :, None] + fl[:, 1]
    both = (x2 >= 0) & (y2 >= 0) & (x2 <= W - 1) & (y2 <= H - 1) & (np.abs(fl[:, 0]) < W / 2) & (np.abs(fl[:, 1]) < H / 2)
    m = np.broadcast_to(both[:, None], a.shape)
    assert m.sum() > 100 and np.abs(a - b)[m].max() < 1e-12


def test_kat3_out_of_range(oracle):
    r = U.rng(3)
    B, C, H, W = 1, 3, 12, 16
    I = U.image(r, B, C, H, W)
    fl = np.zeros((B, 2, H, W), np.float32)
    fl[:, 0] = W / 2          # |fx| >= W/2 -> rejected although x2 may be inside
    ft = U.filt(r, B, 4, H, W)
    assert np.array_equal(oracle.fi_forward("ori", I, fl, ft), I.astype(np.float64))
    fl[:, 0] = 3 * W          # really outside
    assert np.array_equal(oracle.interpolation_forward(I, fl), np.zeros(I.shape))
    g = U.image(r, B, C, H, W, "normal")
    gi1, gi2, gi3, _ = oracle.fi_backward("ori", I, fl, ft, None, g)
    assert not gi1.any() and not gi2.any() and not gi3.any()
    gi1, gi2 = oracle.interpolation_backward(I, fl, g)
    assert not gi1.any() and not gi2.any()


@pytest.mark.parametrize("variant", ["dkr", "deforconv"])
def test_kat4_zero_offsets_reduce_to_ori(oracle, variant):
    r = U.rng(4)
    B, C, H, W = 2, 3, 18, 21
    I, ft = U.image(r, B, C, H, W), U.filt(r, B, 4, H, W, "uniform")
    fl = U.flow(r, B, H, W, "gauss")
    # keep away from x2 == W-1 / y2 == H-1 exactly, where _deforconv's <= test re-classifies a clamped tap
    off = np.zeros((B, 32, H, W), np.float32)
    a = oracle.fi_forward("ori", I, fl, ft)
    b = oracle.fi_forward(variant, I, fl, ft, off)
    assert np.abs(a - b).max() < 1e-12
    g = U.image(r, B, C, H, W, "normal")
    ga, gb = oracle.fi_backward("ori", I, fl, ft, None, g), oracle.fi_backward(variant, I, fl, ft, off, g)
    for x, y in zip(ga[:3], gb[:3]):
        assert np.abs(x - y).max() < 1e-12


def test_dkr_forward_gate_on_filter_size(oracle):
    r = U.rng(5)
    B, C, H, W, F = 1, 2, 8, 9, 5
    out = oracle.fi_forward("dkr", U.image(r, B, C, H, W), U.flow(r, B, H, W, "unit"), U.filt(r, B, F, H, W),
                            U.offsets(r, B, F, H, W))
    assert not out.any()      # filterinterpolation_cuda_kernel.cu:68 -- only F in {4,6} computes
    out = oracle.fi_forward("deforconv", U.image(r, B, C, H, W), U.flow(r, B, H, W, "unit"), U.filt(r, B, F, H, W),
                            U.offsets(r, B, F, H, W))
    assert out.any()


def test_nofilter_equals_deforconv_with_unit_filter(oracle):
    r = U.rng(6)
    B, C, H, W = 1, 3, 14, 15
    I, fl, off = U.image(r, B, C, H, W), U.flow(r, B, H, W, "unit"), U.offsets(r, B, 4, H, W)
    ones = np.ones((B, 16, H, W), np.float32)
    a = oracle.fi_forward("nofilterwithdeforconv", I, fl, off)
    b = oracle.fi_forward("deforconv", I, fl, ones, off)
    assert np.abs(a - b).max() < 1e-12
    g = U.image(r, B, C, H, W, "normal")
    ga = oracle.fi_backward("nofilterwithdeforconv", I, fl, off, None, g)
    gb = oracle.fi_backward("deforconv", I, fl, ones, off, g)
    assert np.abs(ga[0] - gb[0]).max() < 1e-12 and np.abs(ga[1] - gb[1]).max() < 1e-12
    assert np.abs(ga[2] - gb[3]).max() < 1e-12      # offset gradient: gradinput3 here, gradinput4 there


# ----------------------------------------------------------------------------- KAT 5-6: projection
def test_kat5_depth_one_equals_flowprojection(oracle):
    r = U.rng(7)
    fl = U.flow(r, 2, 19, 22, "unit")
    for fh in (0, 1):
        o1, c1 = oracle.flowprojection_forward(fl, None, fh)
        o2, c2 = oracle.flowprojection_forward(fl, np.ones((2, 1, 19, 22), np.float32), fh)
        assert np.array_equal(o1, o2) and np.array_equal(c1, c2)


def test_kat6_zero_and_constant_flow(oracle):
    H, W = 6, 7
    z = np.zeros((1, 2, H, W), np.float32)
    out, cnt = oracle.flowprojection_forward(z)
    assert not out.any()
    assert (cnt[0, 0, 1:-1, 1:-1] == 4).all() and cnt[0, 0, 0, 0] == 1 and cnt[0, 0, -1, -1] == 9
    z[:, 0] = 1
    out, cnt = oracle.flowprojection_forward(z, None, 0)
    assert (cnt[0, 0, :, 0] == 0).all() and (out[0, 0, :, 0] == 0).all()      # column 0 is a hole
    out, cnt = oracle.flowprojection_forward(z, None, 1)
    assert (out[0, 0] == -1).all() and (out[0, 1] == 0).all()                 # filled from the right neighbour


# ----------------------------------------------------------------------------- KAT 7-8: correlation, sepconvflow
def test_kat7_correlation_of_ones(oracle):
    f = np.ones((1, 8, 10, 12), np.float32)
    out = oracle.correlation_forward(f, f)
    assert out.shape == (1, 81, 10, 12)
    assert (out[0, 40] == 1).all()
    for tj in range(-4, 5):
        for ti in range(-4, 5):
            ch = out[0, (tj + 4) * 9 + ti + 4]
            yy, xx = np.mgrid[0:10, 0:12]
            inside = (yy + tj >= 0) & (yy + tj < 10) & (xx + ti >= 0) & (xx + ti < 12)
            assert np.array_equal(ch, inside.astype(np.float64))


@pytest.mark.parametrize("pad,k,md,s1,s2", [(4, 1, 4, 1, 1), (3, 3, 20, 1, 2), (20, 1, 20, 2, 2), (2, 3, 2, 1, 1)])
def test_correlation_shapes(oracle, pad, k, md, s1, s2):
    import math
    H, W = 24, 40
    oc, oh, ow = oracle.correlation_outshape(H, W, pad, k, md, s1, s2)
    kr = (k - 1) // 2
    assert oc == (2 * (md // s2) + 1) ** 2
    assert oh == math.ceil((H + 2 * pad - 2 * (kr + md)) / s1) and ow == math.ceil((W + 2 * pad - 2 * (kr + md)) / s1)


def test_kat8_sepconvflow(oracle):
    ones, zeros = np.ones((1, 5, 3, 4), np.float32), np.zeros((1, 5, 3, 4), np.float32)
    fl = oracle.sepconvflow_forward(ones, zeros)
    assert (fl[0, 1] == 0).all() and (fl[0, 0] == -2000).all()


# ----------------------------------------------------------------------------- KAT 9: finite differences
def _fd(fun, x, idx, eps=2e-3):
    xp, xm = x.copy(), x.copy()
    xp[idx] += np.float32(eps)
    xm[idx] -= np.float32(eps)
    return (fun(xp) - fun(xm)) / (float(xp[idx]) - float(xm[idx]))


def _pick(r, shape, n=6):
    return [tuple(int(r.integers(0, s)) for s in shape) for _ in range(n)]


@pytest.mark.parametrize("variant", ["ori", "dkr", "deforconv", "nofilterwithdeforconv"])
def test_kat9_fi_gradients_match_finite_differences(oracle, variant):
    r = U.rng(20)
    B, C, H, W, F = 1, 2, 12, 13, 4
    I = U.image(r, B, C, H, W)
    # flows with fractional parts away from 0/1 so the integer cell does not flip under +-eps
    fl = (np.floor(r.random((B, 2, H, W)) * 4 - 2) + 0.25 + 0.5 * r.random((B, 2, H, W))).astype(np.float32)
    ft = U.filt(r, B, F, H, W, "uniform")
    off = (0.15 + 0.3 * r.random((B, 2 * F * F, H, W))).astype(np.float32)   # fractional part safely inside (0,1)
    g = U.image(r, B, C, H, W, "normal")

    def fwd(i1=I, i2=fl, i3=None, i4=None):
        if variant == "ori":
            return float((oracle.fi_forward(variant, i1, i2, ft if i3 is None else i3) * g).sum())
        if variant == "nofilterwithdeforconv":
            return float((oracle.fi_forward(variant, i1, i2, off if i3 is None else i3) * g).sum())
        return float((oracle.fi_forward(variant, i1, i2, ft if i3 is None else i3, off if i4 is None else i4) * g).sum())

    if variant == "ori":
        gi1, gi2, gi3, _ = oracle.fi_backward(variant, I, fl, ft, None, g)
    elif variant == "nofilterwithdeforconv":
        gi1, gi2, gi3, _ = oracle.fi_backward(variant, I, fl, off, None, g)
    else:
        gi1, gi2, gi3, gi4 = oracle.fi_backward(variant, I, fl, ft, off, g)

    # interior pixels only: at the border the clamp makes taps coincide, which is still differentiable,
    # but the flow derivative of the reference ignores the clamp
    inner = [(0, k, int(h), int(w)) for k in (0, 1) for h, w in zip(r.integers(3, H - 3, 4), r.integers(3, W - 3, 4))]
    tol = 2e-3
    # flow gradient: valid for every family except the data-dependent-quadrant ones, where the quadrant
    # membership itself depends on the flow (the reference ignores that term) -- checked for ori and dkr
    if variant in ("ori", "dkr"):
        for idx in inner:
            fd = _fd(lambda x: fwd(i2=x), fl, idx)
            assert abs(fd - gi2[idx]) <= tol * (1 + abs(fd)), (variant, "gi2", idx, fd, gi2[idx])
    if variant == "ori":       # the image gradient is the true gradient only without deformation ...
        # ... and only where the flow is in range: forward COPIES input1 at out-of-range pixels
        # (filterinterpolation_cuda_kernel.cu:2814-2819) but backward gives them no gradient (:2863).
        x2 = np.arange(W, dtype=np.float32)[None, :] + fl[0, 0]
        y2 = np.arange(H, dtype=np.float32)[:, None] + fl[0, 1]
        oor = ~((x2 >= 0) & (y2 >= 0) & (x2 <= W - 1) & (y2 <= H - 1) & (np.abs(fl[0, 0]) < W / 2) & (np.abs(fl[0, 1]) < H / 2))
        assert oor.any()
        for idx in _pick(r, I.shape) + [(0, 0, 0, 0), (0, 1, H - 1, W - 1)]:
            fd = _fd(lambda x: fwd(i1=x), I, idx)
            expect = gi1[idx] + (float(g[idx]) if oor[idx[2], idx[3]] else 0.0)
            assert abs(fd - expect) <= tol * (1 + abs(fd)), (variant, "gi1", idx, fd, gi1[idx])
    if variant != "nofilterwithdeforconv":
        for idx in _pick(r, ft.shape):
            fd = _fd(lambda x: fwd(i3=x), ft, idx)
            assert abs(fd - gi3[idx]) <= tol * (1 + abs(fd)), (variant, "gi3", idx, fd, gi3[idx])
    if variant in ("dkr", "deforconv"):
        for idx in [(0, k, h, w) for (_, k, h, w) in _pick(r, (1, 2 * F * F, H - 6, W - 6))]:
            idx = (0, idx[1], idx[2] + 3, idx[3] + 3)
            fd = _fd(lambda x: fwd(i4=x), off, idx, eps=1e-3)
            assert abs(fd - gi4[idx]) <= tol * (1 + abs(fd)), (variant, "gi4", idx, fd, gi4[idx])
    if variant == "nofilterwithdeforconv":
        for idx in [(0, k, h, w) for (_, k, h, w) in _pick(r, (1, 2 * F * F, H - 6, W - 6))]:
            idx = (0, idx[1], idx[2] + 3, idx[3] + 3)
            fd = _fd(lambda x: fwd(i3=x), off, idx, eps=1e-3)
            assert abs(fd - gi3[idx]) <= tol * (1 + abs(fd)), (variant, "goff", idx, fd, gi3[idx])


def test_interpolation_gradients_match_finite_differences(oracle):
    r = U.rng(21)
    B, C, H, W = 1, 3, 10, 11
    I = U.image(r, B, C, H, W)
    fl = (np.floor(r.random((B, 2, H, W)) * 4 - 2) + 0.25 + 0.5 * r.random((B, 2, H, W))).astype(np.float32)
    g = U.image(r, B, C, H, W, "normal")
    gi1, gi2 = oracle.interpolation_backward(I, fl, g)
    for idx in _pick(r, I.shape):
        fd = _fd(lambda x: float((oracle.interpolation_forward(x, fl) * g).sum()), I, idx)
        assert abs(fd - gi1[idx]) <= 2e-3 * (1 + abs(fd))
    for idx in [(0, k, int(h), int(w)) for k in (0, 1) for h, w in zip(r.integers(3, H - 3, 4), r.integers(3, W - 3, 4))]:
        fd = _fd(lambda x: float((oracle.interpolation_forward(I, x) * g).sum()), fl, idx)
        assert abs(fd - gi2[idx]) <= 2e-3 * (1 + abs(fd))


@pytest.mark.parametrize("with_depth", [False, True])
def test_projection_gradients_match_finite_differences(oracle, with_depth):
    r = U.rng(22)
    B, H, W = 1, 9, 10
    fl = (np.floor(r.random((B, 2, H, W)) * 4 - 2) + 0.25 + 0.5 * r.random((B, 2, H, W))).astype(np.float32)
    d = U.depth_inv(r, B, H, W) if with_depth else None
    g = r.standard_normal((B, 2, H, W)).astype(np.float32)

    def fwd(f=fl, dd=d):
        return float((oracle.flowprojection_forward(f, dd, 0)[0] * g).sum())
    out, cnt = oracle.flowprojection_forward(fl, d, 0)
    gi1, gi2 = oracle.flowprojection_backward(fl, d, cnt, out, g)
    # the splat position is piecewise constant in the flow, so d/dflow only sees the splatted VALUE (-flow)
    for idx in _pick(r, fl.shape, 8):
        fd = _fd(lambda x: fwd(f=x), fl, idx, eps=1e-3)
        assert abs(fd - gi1[idx]) <= 2e-3 * (1 + abs(fd)), (idx, fd, gi1[idx])
    if with_depth:
        # Reference quirk, restated on purpose: the depth gradient uses (f - output)
        # (depthflowprojection_cuda_kernel.cu:311-335) although output = -avg(f), so the true derivative
        # is -(g/count)*(f + output).  Finite differences must agree with the TRUE formula, and the oracle
        # must differ from it exactly by the sign of the `output` term.
        true = np.zeros_like(gi2)
        for h in range(H):
            for w in range(W):
                fx, fy = fl[0, 0, h, w], fl[0, 1, h, w]
                x2, y2 = np.float32(w) + fx, np.float32(h) + fy
                if not (x2 >= 0 and y2 >= 0 and x2 <= W - 1 and y2 <= H - 1):
                    continue
                L, T = int(x2), int(y2)
                for yy, xx in ((T, L), (T, min(L + 1, W - 1)), (min(T + 1, H - 1), L), (min(T + 1, H - 1), min(L + 1, W - 1))):
                    for ch, f in ((0, fx), (1, fy)):
                        true[0, 0, h, w] += -g[0, ch, yy, xx] / cnt[0, 0, yy, xx] * (float(f) + out[0, ch, yy, xx])
        for idx in _pick(r, d.shape, 8):
            fd = _fd(lambda x: fwd(dd=x), d, idx, eps=1e-3)
            assert abs(fd - true[idx]) <= 5e-3 * (1 + abs(fd)), (idx, fd, true[idx])
        # oracle (= reference) value: same expression with (f - output)
        ref = np.zeros_like(gi2)
        for h in range(H):
            for w in range(W):
                fx, fy = fl[0, 0, h, w], fl[0, 1, h, w]
                x2, y2 = np.float32(w) + fx, np.float32(h) + fy
                if not (x2 >= 0 and y2 >= 0 and x2 <= W - 1 and y2 <= H - 1):
                    continue
                L, T = int(x2), int(y2)
                for yy, xx in ((T, L), (T, min(L + 1, W - 1)), (min(T + 1, H - 1), L), (min(T + 1, H - 1), min(L + 1, W - 1))):
                    for ch, f in ((0, fx), (1, fy)):
                        ref[0, 0, h, w] += -g[0, ch, yy, xx] / np.float32(cnt[0, 0, yy, xx]) * (float(f) - np.float32(out[0, ch, yy, xx]))
        assert np.abs(ref - gi2).max() < 1e-5


def test_sepconv_gradients_match_finite_differences(oracle):
    r = U.rng(23)
    B, C, H, W, F = 1, 3, 9, 10, 3
    I = U.image(r, B, C, H, W)
    v = r.random((B, F, H - F + 1, W - F + 1), dtype=np.float32)
    hz = r.random((B, F, H - F + 1, W - F + 1), dtype=np.float32)
    g = r.standard_normal((B, C, H - F + 1, W - F + 1)).astype(np.float32)
    gi1, gi2, gi3 = oracle.sepconv_backward(I, v, hz, g)
    for arr, gi, name in ((I, gi1, 0), (v, gi2, 1), (hz, gi3, 2)):
        for idx in _pick(r, arr.shape, 5):
            args = [I, v, hz]

            def f(x, name=name):
                a = list(args)
                a[name] = x
                return float((oracle.sepconv_forward(*a) * g).sum())
            fd = _fd(f, arr, idx)
            assert abs(fd - gi[idx]) <= 2e-3 * (1 + abs(fd))


def test_sepconvflow_gradients_match_finite_differences(oracle):
    r = U.rng(24)
    B, F, Ho, Wo = 1, 5, 4, 5
    v = (0.2 + r.random((B, F, Ho, Wo))).astype(np.float32)
    hz = (0.2 + r.random((B, F, Ho, Wo))).astype(np.float32)
    g = r.standard_normal((B, 2, Ho, Wo)).astype(np.float32)
    gi2, gi3 = oracle.sepconvflow_backward(v, hz, g)
    for idx in _pick(r, v.shape, 5):
        fd = _fd(lambda x: float((oracle.sepconvflow_forward(x, hz) * g).sum()), v, idx, eps=1e-3)
        assert abs(fd - gi2[idx]) <= 5e-3 * (1 + abs(fd))
        fd = _fd(lambda x: float((oracle.sepconvflow_forward(v, x) * g).sum()), hz, idx, eps=1e-3)
        assert abs(fd - gi3[idx]) <= 5e-3 * (1 + abs(fd))


@pytest.mark.parametrize("pad,k,md,s1,s2", [(4, 1, 4, 1, 1), (2, 1, 2, 1, 1), (5, 1, 4, 1, 2)])
def test_correlation_gradients_match_finite_differences(oracle, pad, k, md, s1, s2):
    r = U.rng(25)
    B, C, H, W = 1, 3, 9, 11
    f1, f2 = U.image(r, B, C, H, W, "normal"), U.image(r, B, C, H, W, "normal")
    out = oracle.correlation_forward(f1, f2, pad, k, md, s1, s2)
    g = r.standard_normal(out.shape).astype(np.float32)
    gi1, gi2 = oracle.correlation_backward(f1, f2, g, pad, k, md, s1, s2)
    for idx in _pick(r, f1.shape, 6):
        fd = _fd(lambda x: float((oracle.correlation_forward(x, f2, pad, k, md, s1, s2) * g).sum()), f1, idx)
        assert abs(fd - gi1[idx]) <= 2e-3 * (1 + abs(fd)), ("gi1", idx, fd, gi1[idx])
        fd = _fd(lambda x: float((oracle.correlation_forward(f1, x, pad, k, md, s1, s2) * g).sum()), f2, idx)
        assert abs(fd - gi2[idx]) <= 2e-3 * (1 + abs(fd)), ("gi2", idx, fd, gi2[idx])


# ------------------------------------------------------------------------------ PWCDCNet.warp
def _reference_pwc_warp(x, flo, align_corners):
    """PWCNet/PWCNet.py:159-199 as written (without the pre-allocated grid and .cuda()), on CPU tensors.
    align_corners=True is grid_sample of the reference's pinned torch 1.0.1, False the default since torch 1.3."""
    import torch
    B, C, H, W = x.size()
    xx = torch.arange(0, W).view(1, -1).repeat(H, 1).view(1, 1, H, W).repeat(B, 1, 1, 1)
    yy = torch.arange(0, H).view(-1, 1).repeat(1, W).view(1, 1, H, W).repeat(B, 1, 1, 1)
    vgrid = torch.cat((xx, yy), 1).to(x.dtype) + flo
    vgrid = torch.stack([2.0 * vgrid[:, 0] / max(W - 1, 1) - 1.0, 2.0 * vgrid[:, 1] / max(H - 1, 1) - 1.0], 1)
    vgrid = vgrid.permute(0, 2, 3, 1)
    output = torch.nn.functional.grid_sample(x, vgrid, align_corners=align_corners)
    mask = torch.nn.functional.grid_sample(torch.ones_like(x), vgrid, align_corners=align_corners).detach().clone()
    mask[mask < 0.9999] = 0
    mask[mask > 0] = 1
    return output * mask


@pytest.mark.parametrize("ac", [True, False])
def test_pwc_warp_oracle_is_pinned_to_torch_grid_sample(oracle, ac):
    """The numpy restatement of PWCDCNet.warp against the reference's own code path run on torch's CPU grid_sample
    (float32 for the values the reference computes, float64 autograd for the gradients)."""
    import torch
    r = U.rng(3100)
    for (B, C, H, W) in [(2, 3, 17, 23), (1, 5, 8, 40), (1, 2, 1, 9)]:
        x = r.standard_normal((B, C, H, W)).astype(np.float32)
        flo = (r.standard_normal((B, 2, H, W)) * 3).astype(np.float32)
        flo[:, :, 0, 0] = 0.0                      # exact-integer landing
        flo[:, 0, -1, -1] = 50.0                   # far outside
        ref = _reference_pwc_warp(torch.from_numpy(x), torch.from_numpy(flo), ac).numpy()
        got = oracle.pwc_warp_forward(x, flo, ac)
        assert np.abs(got - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max())
        # gradients: float64 torch autograd on float32-representable inputs (away from knife edges the index
        # arithmetic agrees between float32 and float64)
        flo_s = (np.round(flo * 8) / 8 + 0.0625).astype(np.float32)
        tx = torch.from_numpy(x).double().requires_grad_()
        tf = torch.from_numpy(flo_s).double().requires_grad_()
        g = r.standard_normal((B, C, H, W))
        _reference_pwc_warp(tx, tf, ac).backward(torch.from_numpy(g))
        # the oracle's float32 geometry must select the same corners: skip shapes where W - 1 makes 1/8-pixel offsets inexact
        gx, gf = oracle.pwc_warp_backward(x, flo_s, g.astype(np.float32), ac)
        g32 = g.astype(np.float32).astype(np.float64)
        tx2 = torch.from_numpy(x).double().requires_grad_()
        tf2 = torch.from_numpy(flo_s).double().requires_grad_()
        _reference_pwc_warp(tx2, tf2, ac).backward(torch.from_numpy(g32))
        assert np.abs(gx - tx2.grad.numpy()).max() <= 1e-4 * max(1.0, np.abs(gx).max())
        assert np.abs(gf - tf2.grad.numpy()).max() <= 1e-4 * max(1.0, np.abs(gf).max())


def test_pwc_warp_with_the_pinned_torch_rule_is_an_exact_warp(oracle):
    """torch 1.0.1's grid_sample (align_corners=True) on the reference's normalisation samples exactly at x + flow:
    zero flow is the identity with a full mask, an integer flow is a pure shift (the ADVICE finding of round 1: the
    align_corners=False rule instead shifts by up to half a pixel and masks row 0 / column 0 at zero flow)."""
    r = U.rng(3150)
    x = r.standard_normal((1, 3, 9, 17)).astype(np.float32)
    z = np.zeros((1, 2, 9, 17), np.float32)
    assert np.abs(oracle.pwc_warp_forward(x, z, True) - x).max() <= 1e-5
    modern = oracle.pwc_warp_forward(x, z, False)
    assert not modern[:, :, 0, :].any() and not modern[:, :, :, 0].any()          # the masked border of the modern rule
    f = z.copy()
    f[:, 0] = 2.0
    f[:, 1] = -1.0
    out = oracle.pwc_warp_forward(x, f, True)
    assert np.abs(out[:, :, 1:, :-2] - x[:, :, :-1, 2:]).max() <= 1e-5
    assert not out[:, :, 0, :].any() and not out[:, :, :, -2:].any()              # sampled outside the plane


def test_mindepth_oracle_known_answers(oracle):
    """MinDepthFlowProjection restatement: the closest surface wins its cell, ties go to the lowest pixel index,
    non-positive input2 never wins, zero flow is the identity."""
    H, W = 4, 6
    flo = np.zeros((1, 2, H, W), np.float32)
    dep = np.full((1, 1, H, W), 0.5, np.float32)
    out, cnt = oracle.mindepth_forward(flo, dep)
    assert not out.any() and (cnt == 0.5).all()
    # pixels (0,1) and (0,2) both land on cell (0,3): the larger input2 wins
    flo[0, 0, 0, 1], flo[0, 0, 0, 2] = 2.0, 1.0
    dep[0, 0, 0, 1], dep[0, 0, 0, 2], dep[0, 0, 0, 3] = 0.9, 0.7, 0.1
    out, cnt = oracle.mindepth_forward(flo, dep)
    assert cnt[0, 0, 0, 3] == np.float32(0.9) and out[0, 0, 0, 3] == -2.0
    assert cnt[0, 0, 0, 1] == 0 and cnt[0, 0, 0, 2] == 0          # their own cells became holes
    # a tie: the lowest pixel index wins
    dep[0, 0, 0, 2] = 0.9
    out, cnt = oracle.mindepth_forward(flo, dep)
    assert out[0, 0, 0, 3] == -2.0
    # non-positive input2 never wins
    dep[:] = 0.0
    out, cnt = oracle.mindepth_forward(flo, dep)
    assert not cnt.any() and not out.any()
    # hole filling: the hole at (0,1) takes the mean of its nearest non-hole neighbours
    dep[:] = 0.5
    dep[0, 0, 0, 1], dep[0, 0, 0, 2], dep[0, 0, 0, 3] = 0.9, 0.7, 0.1
    out, cnt = oracle.mindepth_forward(flo, dep, fillhole=1)
    assert cnt[0, 0, 0, 1] == 0 and out[0, 0, 0, 1] == (0.0 + -2.0 + 0.0) / 3     # left (0,0), right (0,3), down (1,1)
    # backward: the winner receives -gradoutput of the matching corners
    g = np.ones((1, 2, H, W), np.float32)
    gi1, gi2 = oracle.mindepth_backward(flo, dep, cnt.astype(np.float32), g)
    assert gi1[0, 0, 0, 1] == -1.0 and gi1[0, 0, 0, 2] == 0.0 and not gi2.any()


# ------------------------------------------------------------------------------ PyTorch-CPU fp32 baseline (bench.py leg)
def test_torch_cpu_baseline_matches_the_oracle(oracle):
    """oracle/torch_cpu.py -- the vectorised PyTorch-CPU fp32 path bench.py times as the second CPU baseline
     -- computes what the float64 oracle computes (fp32 tolerance; atomically ordered sums 1e-4)."""
    import torch
    from oracle import torch_cpu as T
    r = U.rng(3300)
    B, C, H, W = 2, 3, 23, 37
    I, ft = U.image(r, B, C, H, W), U.filt(r, B, 4, H, W)
    for fk in ("gauss", "stress"):
        fl = U.flow(r, B, H, W, fk)
        got = T.fi_ori_forward(torch.from_numpy(I), torch.from_numpy(fl), torch.from_numpy(ft)).numpy()
        U.assert_close(got, oracle.fi_forward("ori", I, fl, ft), 2e-5, f"torch-CPU FilterInterpolation ({fk})")
        d = U.depth_inv(r, B, H, W)
        for depth in (None, d):
            for fill in (0, 1):
                out, cnt = T.flowprojection_forward(torch.from_numpy(fl), None if depth is None else torch.from_numpy(depth), fill)
                ref, rc = oracle.flowprojection_forward(fl, depth, fill)
                U.assert_close(out.numpy(), ref, U.RTOL_ATOMIC, f"torch-CPU projection ({fk}, depth={depth is not None}, fill={fill})")
                U.assert_close(cnt.numpy(), rc, U.RTOL_ATOMIC, "torch-CPU projection count")
    # wide holes: constant shift leaves whole columns / rows empty
    fl = np.zeros((1, 2, 9, 14), np.float32)
    fl[:, 0] = 3.0
    fl[:, 1, 4:] = -2.0
    out, _ = T.flowprojection_forward(torch.from_numpy(fl), None, 1)
    U.assert_close(out.numpy(), oracle.flowprojection_forward(fl, None, 1)[0], U.RTOL_ATOMIC, "torch-CPU hole filling")
    f1, f2 = U.image(r, 2, 8, 12, 20, "normal"), U.image(r, 2, 8, 12, 20, "normal")
    U.assert_close(T.correlation_forward(torch.from_numpy(f1), torch.from_numpy(f2)).numpy(),
                   oracle.correlation_forward(f1, f2, 4, 1, 4, 1, 1), 2e-5, "torch-CPU correlation")
