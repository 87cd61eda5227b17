"""Parity of the sm_100a kernels (called through the reference-shaped Python API, which goes through the
C ABI) against the float64 CPU oracle on identical seeded inputs.

Tolerances (BASELINE.json north_star): 1e-5 for fp32 forward outputs and thread-private gradients,
1e-4 for atomically accumulated results; exact for counts.  The error measure is
|gpu - oracle| / (|oracle| + max|oracle|)  (tests/util.py)."""
import numpy as np
import pytest
import torch

import util as U

pytestmark = pytest.mark.gpu


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def host(t):
    return t.detach().cpu().numpy()


@pytest.fixture(scope="module", autouse=True)
def _native_library_is_loaded(lib):
    lib._lib.load()
    before = lib.launch_count()
    yield
    assert lib.launch_count() > before, "no kernel of libvfidkr_b200.so was launched by the GPU tests"


# ------------------------------------------------------------------------------ FilterInterpolation _ori
FI_SHAPES = [  # B, C, H, W, flow kind, filter kind
    (1, 3, 256, 448, "gauss", "softmax"),     # BASELINE config 1
    (1, 3, 256, 448, "unit", "uniform"),      # test_module.py recipe
    (2, 3, 37, 29, "stress", "uniform"),      # ragged, out-of-range pixels, border landings
    (1, 1, 1, 1, "unit", "uniform"),
    (1, 2, 3, 5, "unit", "softmax"),
    (2, 5, 40, 70, "smooth", "softmax"),
    (1, 196, 24, 40, "gauss", "softmax"),     # DAIN_slowmotion context width
    (3, 4, 33, 130, "gauss", "uniform"),
]


@pytest.mark.parametrize("B,C,H,W,fk,wk", FI_SHAPES)
def test_fi_ori_forward_backward(lib, oracle, B, C, H, W, fk, wk):
    r = U.rng(1001)
    I, fl, ft = U.image(r, B, C, H, W), U.flow(r, B, H, W, fk), U.filt(r, B, 4, H, W, wk)
    ti, tf, tw = cu(I).requires_grad_(), cu(fl).requires_grad_(), cu(ft).requires_grad_()
    out = lib.FilterInterpolationModule()(ti, tf, tw)
    U.assert_close(host(out), oracle.fi_forward("ori", I, fl, ft), U.RTOL_FWD, "fi_ori forward")
    g = U.image(r, B, C, H, W, "normal")
    out.backward(cu(g))
    gi1, gi2, gi3, _ = oracle.fi_backward("ori", I, fl, ft, None, g)
    U.assert_close(host(ti.grad), gi1, U.RTOL_ATOMIC, "fi_ori gradinput1 (atomic scatter)")
    U.assert_close(host(tf.grad), gi2, U.RTOL_FWD, "fi_ori gradinput2")
    U.assert_close(host(tw.grad), gi3, U.RTOL_FWD, "fi_ori gradinput3")


@pytest.mark.parametrize("F", [2, 5, 6])
def test_fi_ori_runtime_filter_sizes(lib, oracle, F):
    r = U.rng(1002 + F)
    B, C, H, W = 2, 3, 31, 45
    I, fl, ft = U.image(r, B, C, H, W), U.flow(r, B, H, W, "stress"), U.filt(r, B, F, H, W, "uniform")
    ti, tf, tw = cu(I).requires_grad_(), cu(fl).requires_grad_(), cu(ft).requires_grad_()
    out = lib.FilterInterpolationModule()(ti, tf, tw)
    U.assert_close(host(out), oracle.fi_forward("ori", I, fl, ft), U.RTOL_FWD, f"fi_ori F={F} forward")
    g = U.image(r, B, C, H, W, "normal")
    out.backward(cu(g))
    gi1, gi2, gi3, _ = oracle.fi_backward("ori", I, fl, ft, None, g)
    U.assert_close(host(ti.grad), gi1, U.RTOL_ATOMIC, "gradinput1")
    U.assert_close(host(tf.grad), gi2, U.RTOL_FWD, "gradinput2")
    U.assert_close(host(tw.grad), gi3, U.RTOL_FWD, "gradinput3")


def test_fi_ori_training_step_batch16(lib, oracle):
    """BASELINE config 3: fwd+bwd at B=16, 256x448, upstream gradient = the output (test_module.py:226)."""
    r = U.rng(1003)
    B, C, H, W = 16, 3, 256, 448
    I, fl, ft = U.image(r, B, C, H, W), U.flow(r, B, H, W, "gauss"), U.filt(r, B, 4, H, W)
    ti, tf, tw = cu(I).requires_grad_(), cu(fl).requires_grad_(), cu(ft).requires_grad_()
    out = lib.FilterInterpolationModule()(ti, tf, tw)
    ref = oracle.fi_forward("ori", I, fl, ft)
    U.assert_close(host(out), ref, U.RTOL_FWD, "forward")
    out.backward(out.detach())
    g = host(out)
    gi1, gi2, gi3, _ = oracle.fi_backward("ori", I, fl, ft, None, g)
    U.assert_close(host(ti.grad), gi1, U.RTOL_ATOMIC, "gradinput1")
    U.assert_close(host(tf.grad), gi2, U.RTOL_FWD, "gradinput2")
    U.assert_close(host(tw.grad), gi3, U.RTOL_FWD, "gradinput3")


def test_fi_known_answers_on_gpu(lib):
    r = U.rng(1004)
    I = U.image(r, 2, 3, 40, 52)
    ft = np.zeros((2, 16, 40, 52), np.float32)
    ft[:, 5] = 1
    zero = np.zeros((2, 2, 40, 52), np.float32)
    out = lib.FilterInterpolationModule()(cu(I), cu(zero), cu(ft))
    assert torch.equal(out.cpu(), torch.from_numpy(I))                      # KAT 1, bit exact
    far = zero.copy()
    far[:, 0] = 26.0                                                         # |fx| >= W/2
    out = lib.FilterInterpolationModule()(cu(I), cu(far), cu(U.filt(r, 2, 4, 40, 52)))
    assert torch.equal(out.cpu(), torch.from_numpy(I))                      # KAT 3: copy, not zero
    ft2 = np.zeros((2, 16, 40, 52), np.float32)
    ft2[:, [5, 6, 9, 10]] = 1
    fl = U.flow(r, 2, 40, 52, "gauss")
    a = lib.FilterInterpolationModule()(cu(I), cu(fl), cu(ft2))
    b = lib.InterpolationChModule(3)(cu(I), cu(fl))
    x2 = np.arange(52, dtype=np.float32)[None, None, :] + fl[:, 0]
    y2 = np.arange(40, dtype=np.float32)[None, :, None] + fl[:, 1]
    m = (x2 >= 0) & (y2 >= 0) & (x2 <= 51) & (y2 <= 39) & (np.abs(fl[:, 0]) < 26) & (np.abs(fl[:, 1]) < 20)
    m = np.broadcast_to(m[:, None], I.shape)
    assert np.abs(host(a) - host(b))[m].max() < 1e-5                        # KAT 2


# ------------------------------------------------------------------------------ DKR families
DKR_SHAPES = [(1, 3, 256, 448), (2, 3, 37, 29), (1, 5, 20, 33), (1, 1, 2, 3)]


@pytest.mark.parametrize("variant", ["dkr", "deforconv", "nofilterwithdeforconv"])
@pytest.mark.parametrize("B,C,H,W", DKR_SHAPES)
def test_fi_dkr_families(lib, oracle, variant, B, C, H, W):
    r = U.rng(1100 + len(variant))
    F = 4
    I, fl = U.image(r, B, C, H, W), U.flow(r, B, H, W, "stress" if H > 8 else "unit")
    ft, off = U.filt(r, B, F, H, W, "uniform"), U.offsets(r, B, F, H, W, 0.45)
    ti, tf = cu(I).requires_grad_(), cu(fl).requires_grad_()
    g = U.image(r, B, C, H, W, "normal")
    if variant == "nofilterwithdeforconv":
        to = cu(off).requires_grad_()
        out = lib.FilterInterpolationModule(variant)(ti, tf, to)
        U.assert_close(host(out), oracle.fi_forward(variant, I, fl, off), U.RTOL_FWD, f"{variant} forward")
        out.backward(cu(g))
        gi1, gi2, gi3, _ = oracle.fi_backward(variant, I, fl, off, None, g)
        U.assert_close(host(to.grad), gi3, U.RTOL_FWD, "offset gradient")
    else:
        tw, to = cu(ft).requires_grad_(), cu(off).requires_grad_()
        out = lib.FilterInterpolationModule(variant)(ti, tf, tw, to)
        U.assert_close(host(out), oracle.fi_forward(variant, I, fl, ft, off), U.RTOL_FWD, f"{variant} forward")
        out.backward(cu(g))
        gi1, gi2, gi3, gi4 = oracle.fi_backward(variant, I, fl, ft, off, g)
        U.assert_close(host(tw.grad), gi3, U.RTOL_FWD, "filter gradient")
        U.assert_close(host(to.grad), gi4, U.RTOL_FWD, "offset gradient")
    U.assert_close(host(ti.grad), gi1, U.RTOL_ATOMIC, "image gradient (atomic scatter)")
    U.assert_close(host(tf.grad), gi2, U.RTOL_FWD, "flow gradient")


@pytest.mark.parametrize("variant", ["dkr", "deforconv"])
def test_fi_dkr_zero_offsets_equal_ori_on_gpu(lib, variant):
    r = U.rng(1200)
    B, C, H, W = 2, 3, 64, 80
    I, fl, ft = U.image(r, B, C, H, W), U.flow(r, B, H, W, "gauss"), U.filt(r, B, 4, H, W)
    a = lib.FilterInterpolationModule()(cu(I), cu(fl), cu(ft))
    b = lib.FilterInterpolationModule(variant)(cu(I), cu(fl), cu(ft), cu(np.zeros((B, 32, H, W), np.float32)))
    assert U.max_err(host(b), host(a).astype(np.float64)) < 1e-6           # KAT 4


def test_fi_dkr_wild_offsets_are_memory_safe_and_defined(lib, oracle):
    """Outside the in-contract domain the reference reads out of the plane (undefined); the product and
    the oracle define those reads by clamping the read index.  Runs without faulting and matches."""
    r = U.rng(1201)
    B, C, H, W, F = 1, 3, 24, 31, 4
    I, fl, ft = U.image(r, B, C, H, W), U.flow(r, B, H, W, "unit"), U.filt(r, B, F, H, W)
    off = (r.standard_normal((B, 2 * F * F, H, W)) * 40).astype(np.float32)
    for variant in ("dkr", "deforconv"):
        out = lib.FilterInterpolationModule(variant)(cu(I), cu(fl), cu(ft), cu(off))
        torch.cuda.synchronize()
        ref = oracle.fi_forward(variant, I, fl, ft, off)
        U.assert_close(host(out), ref, U.RTOL_ATOMIC, f"{variant} with out-of-contract offsets")


def test_fi_dkr_forward_gate(lib, oracle):
    r = U.rng(1202)
    B, C, H, W, F = 1, 2, 16, 20, 5
    args = [U.image(r, B, C, H, W), U.flow(r, B, H, W, "unit"), U.filt(r, B, F, H, W), U.offsets(r, B, F, H, W)]
    out = lib.FilterInterpolationModule("dkr")(*map(cu, args))
    assert not out.any()                                                    # F not in {4,6}: zeros
    out = lib.FilterInterpolationModule("deforconv")(*map(cu, args))
    U.assert_close(host(out), oracle.fi_forward("deforconv", *args), U.RTOL_FWD, "deforconv F=5")
    args = [U.image(r, B, C, H, W), U.flow(r, B, H, W, "unit"), U.filt(r, B, 6, H, W), U.offsets(r, B, 6, H, W)]
    out = lib.FilterInterpolationModule("dkr")(*map(cu, args))
    U.assert_close(host(out), oracle.fi_forward("dkr", *args), U.RTOL_FWD, "dkr F=6")


# ------------------------------------------------------------------------------ projection
PROJ_SHAPES = [(1, 256, 448, "gauss"), (2, 37, 29, "stress"), (1, 1, 1, "unit"), (2, 64, 96, "smooth"), (1, 512, 704, "unit")]


@pytest.mark.parametrize("B,H,W,fk", PROJ_SHAPES)
@pytest.mark.parametrize("with_depth", [False, True])
def test_projection_forward_backward(lib, oracle, B, H, W, fk, with_depth):
    r = U.rng(1300)
    fl = U.flow(r, B, H, W, fk)
    d = U.depth_inv(r, B, H, W) if with_depth else None
    for requires_grad in (True, False):      # False -> hole filling (inference)
        tf = cu(fl).requires_grad_(requires_grad)
        if with_depth:
            td = cu(d).requires_grad_(requires_grad)
            out = lib.DepthFlowProjectionModule(tf.requires_grad)(tf, td)
        else:
            out = lib.FlowProjectionModule(tf.requires_grad)(tf)
        ref, cnt = oracle.flowprojection_forward(fl, d, 0 if requires_grad else 1)
        U.assert_close(host(out), ref, U.RTOL_ATOMIC, f"projection forward fillhole={not requires_grad}")
        if requires_grad:
            g = r.standard_normal((B, 2, H, W)).astype(np.float32)
            out.backward(cu(g))
            # the backward consumes the forward's own fp32 count / output, as the reference does
            gi1, gi2 = oracle.flowprojection_backward(fl, d, cnt.astype(np.float32), ref.astype(np.float32), g)
            U.assert_close(host(tf.grad), gi1, U.RTOL_ATOMIC, "gradinput1")
            if with_depth:
                U.assert_close(host(td.grad), gi2, U.RTOL_ATOMIC, "gradinput2 (depth)")


def test_flowprojection_count_is_bit_exact(lib, oracle):
    """Counts are small integers in fp32: the splat targets (index arithmetic) must match exactly."""
    from ctypes import c_void_p
    from vfidkr_b200 import _lib
    r = U.rng(1301)
    B, H, W = 2, 97, 131
    fl = U.flow(r, B, H, W, "stress")
    tf = cu(fl)
    count = torch.empty((B, 1, H, W), device="cuda")
    out = torch.empty((B, 2, H, W), device="cuda")
    _lib.call("vfidkr_flowprojection_forward", c_void_p(tf.data_ptr()), c_void_p(count.data_ptr()),
              c_void_p(out.data_ptr()), B, H, W, 0, c_void_p(torch.cuda.current_stream().cuda_stream))
    _, cnt = oracle.flowprojection_forward(fl, None, 0)
    assert np.array_equal(host(count).astype(np.float64), cnt)


PROJ_SHAPES = [(2, 37, 29, "stress"), (3, 97, 131, "stress"), (8, 256, 448, "gauss"), (5, 64, 200, "smooth"),
               (2, 1, 1, "unit"), (4, 9, 33, "unit"), (7, 40, 96, "gauss"), (3, 67, 132, "stress"), (2, 130, 260, "stress"), (1, 23, 4, "unit")]


@pytest.mark.parametrize("B,H,W,fk", PROJ_SHAPES)
@pytest.mark.parametrize("with_depth", [False, True])
def test_projection_forward_through_the_c_abi(lib, oracle, B, H, W, fk, with_depth):
    """Splat + box pass + bitmask-driven hole filling (64-row column words, dense lanes) through the C ABI, with and
    without hole filling: ragged shapes, one-pixel frames, heights that are not a multiple of 8 / 64, wide holes."""
    from vfidkr_b200 import _lib
    from vfidkr_b200._common import ptr, stream_ptr
    r = U.rng(1350 + B + H)
    fl = U.flow(r, B, H, W, fk)
    d = U.depth_inv(r, B, H, W) if with_depth else None
    tf, td = cu(fl), (cu(d) if with_depth else None)
    sp = stream_ptr(tf.device)
    for fill in (0, 1):
        cnt, out = torch.full((B, 1, H, W), -7.0, device="cuda"), torch.full((B, 2, H, W), -7.0, device="cuda")
        if with_depth:
            _lib.call("vfidkr_depthflowprojection_forward", ptr(tf), ptr(td), ptr(cnt), ptr(out), B, H, W, fill, sp)
        else:
            _lib.call("vfidkr_flowprojection_forward", ptr(tf), ptr(cnt), ptr(out), B, H, W, fill, sp)
        ref, rc = oracle.flowprojection_forward(fl, d, fill)
        U.assert_close(host(out), ref, U.RTOL_ATOMIC, f"projection fillhole={fill}")
        if with_depth:
            U.assert_close(host(cnt), rc, U.RTOL_ATOMIC, f"projection count fillhole={fill}")
        else:
            assert np.array_equal(host(cnt).astype(np.float64), rc), "FlowProjection counts are exact integers"


def test_projection_hole_filling_across_tall_and_wide_holes(lib, oracle):
    """Holes taller than one 64-row column word and wider than one 32-column row word, holes reaching the borders, and
    a frame that is one hole (nothing to fill from)."""
    B, H, W = 3, 200, 150
    fl = np.zeros((B, 2, H, W), np.float32)
    fl[0, 0, :, 40:] = 70.0                 # columns 40..109 empty over the full height
    fl[1, 1, 30:, :] = 140.0                # rows 30..169 empty over the full width
    fl[2, 0] = 3.0 * W                      # everything out of range
    out = lib.FlowProjectionModule(False)(cu(fl))
    ref, _ = oracle.flowprojection_forward(fl, None, 1)
    U.assert_close(host(out), ref, U.RTOL_ATOMIC, "hole filling, large holes")
    assert not host(out)[2].any()


def test_projection_is_repeatable_back_to_back(lib):
    """Back-to-back launches on one stream reuse the cached scratch block (dirty scratch image, stale bitmaps): every
    launch must clear what it needs itself."""
    r = U.rng(1399)
    B, H, W = 6, 120, 168
    fl, d = cu(U.flow(r, B, H, W, "stress")), cu(U.depth_inv(r, B, H, W))
    mod = lib.DepthFlowProjectionModule(False)
    first = mod(fl, d).clone()
    for _ in range(5):
        other = mod(cu(U.flow(r, B, H, W, "gauss")), d)      # different data through the same scratch
        again = mod(fl, d)
        assert torch.isfinite(other).all()
        assert (again - first).abs().max().item() <= 2e-5 * first.abs().max().item()


def test_projection_known_answers_on_gpu(lib):
    z = torch.zeros(1, 2, 6, 7, device="cuda")
    out = lib.FlowProjectionModule(False)(z)
    assert not out.any()
    z[:, 0] = 1
    out = lib.FlowProjectionModule(False)(z)
    assert (out[0, 0] == -1).all() and (out[0, 1] == 0).all()               # KAT 6 incl. hole filling
    fl = cu(U.flow(U.rng(5), 2, 30, 41, "unit"))
    a = lib.FlowProjectionModule(True)(fl)
    b = lib.DepthFlowProjectionModule(True)(fl, torch.ones(2, 1, 30, 41, device="cuda"))
    assert U.max_err(host(b), host(a).astype(np.float64)) < 1e-6           # KAT 5


# ------------------------------------------------------------------------------ Interpolation(Ch)
@pytest.mark.parametrize("B,C,H,W,fk", [(1, 3, 256, 448, "gauss"), (2, 3, 37, 29, "stress"), (1, 7, 20, 33, "unit"), (1, 1, 1, 1, "unit")])
def test_interpolation(lib, oracle, B, C, H, W, fk):
    r = U.rng(1400)
    I, fl = U.image(r, B, C, H, W), U.flow(r, B, H, W, fk)
    ti, tf = cu(I).requires_grad_(), cu(fl).requires_grad_()
    mod = lib.InterpolationModule() if C == 3 else lib.InterpolationChModule(C)
    out = mod(ti, tf)
    U.assert_close(host(out), oracle.interpolation_forward(I, fl), U.RTOL_FWD, "interpolation forward")
    g = U.image(r, B, C, H, W, "normal")
    out.backward(cu(g))
    gi1, gi2 = oracle.interpolation_backward(I, fl, g)
    U.assert_close(host(ti.grad), gi1, U.RTOL_ATOMIC, "gradinput1 (atomic scatter)")
    U.assert_close(host(tf.grad), gi2, U.RTOL_FWD, "gradinput2")


def test_interpolation_requires_three_channels(lib):
    with pytest.raises(lib.VfidkrError):
        lib.InterpolationModule()(torch.rand(1, 4, 8, 8, device="cuda"), torch.zeros(1, 2, 8, 8, device="cuda"))


# ------------------------------------------------------------------------------ SeparableConv(Flow)
@pytest.mark.parametrize("B,H,W,F", [(1, 128, 128, 51), (2, 20, 33, 5), (1, 9, 9, 9), (2, 70, 150, 11), (1, 40, 90, 4),
                                     (1, 150, 140, 120)])   # several ragged tiles; odd widths; F too large for the tiled kernels
def test_separableconv(lib, oracle, B, H, W, F):
    r = U.rng(1500)
    C, Ho, Wo = 3, H - F + 1, W - F + 1
    I = U.image(r, B, C, H, W)
    v = (1.0 / F + 0.1 / F * r.random((B, F, Ho, Wo))).astype(np.float32)      # test_module.py:903-907
    hz = (1.0 / F + 0.1 / F * r.random((B, F, Ho, Wo))).astype(np.float32)
    ti, tv, th = cu(I).requires_grad_(), cu(v).requires_grad_(), cu(hz).requires_grad_()
    out = lib.SeparableConvModule(F)(ti, tv, th)
    U.assert_close(host(out), oracle.sepconv_forward(I, v, hz), U.RTOL_FWD * 3, "sepconv forward")
    g = r.standard_normal((B, C, Ho, Wo)).astype(np.float32)
    out.backward(cu(g))
    gi1, gi2, gi3 = oracle.sepconv_backward(I, v, hz, g)
    U.assert_close(host(ti.grad), gi1, U.RTOL_ATOMIC, "gradinput1")
    U.assert_close(host(tv.grad), gi2, U.RTOL_ATOMIC, "gradinput2")
    U.assert_close(host(th.grad), gi3, U.RTOL_ATOMIC, "gradinput3")


@pytest.mark.parametrize("B,Ho,Wo,F", [(1, 78, 78, 51), (2, 16, 29, 5)])
def test_separableconvflow(lib, oracle, B, Ho, Wo, F):
    import warnings
    r = U.rng(1600)
    v = (1.0 / F + 0.1 / F * r.random((B, F, Ho, Wo))).astype(np.float32)
    hz = (1.0 / F + 0.1 / F * r.random((B, F, Ho, Wo))).astype(np.float32)
    v[0, :, 0, 0] = 0.0                                                      # sentinel pixel
    I = U.image(r, B, 3, Ho + F - 1, Wo + F - 1)
    ti, tv, th = cu(I).requires_grad_(), cu(v).requires_grad_(), cu(hz).requires_grad_()
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        mod = lib.SeparableConvFlowModule(F)
    out = mod(ti, tv, th)
    ref = oracle.sepconvflow_forward(v, hz)
    assert host(out)[0, 1, 0, 0] == -2000.0
    U.assert_close(host(out), ref, U.RTOL_FWD, "sepconvflow forward")
    g = r.standard_normal((B, 2, Ho, Wo)).astype(np.float32)
    out.backward(cu(g))
    gi2, gi3 = oracle.sepconvflow_backward(v, hz, g)
    U.assert_close(host(tv.grad), gi2, U.RTOL_FWD * 3, "gradinput2")
    U.assert_close(host(th.grad), gi3, U.RTOL_FWD * 3, "gradinput3")
    assert not ti.grad.any()


# ------------------------------------------------------------------------------ Correlation
@pytest.mark.parametrize("B,C,H,W", [(1, 32, 64, 112), (2, 196, 4, 7), (1, 64, 18, 31), (2, 5, 9, 13), (1, 128, 8, 14),
                                     (8, 128, 36, 62)])   # last: PWC level 5 at 1080p x 8 -- odd width, row-padded TMA path
def test_correlation_pwc_configuration(lib, oracle, B, C, H, W):
    r = U.rng(1700)
    f1, f2 = U.image(r, B, C, H, W, "normal"), U.image(r, B, C, H, W, "normal")
    t1, t2 = cu(f1).requires_grad_(), cu(f2).requires_grad_()
    out = lib.Correlation(pad_size=4, kernel_size=1, max_displacement=4, stride1=1, stride2=1, corr_multiply=1)(t1, t2)
    ref = oracle.correlation_forward(f1, f2, 4, 1, 4, 1, 1)
    assert tuple(out.shape) == ref.shape
    U.assert_close(host(out), ref, U.RTOL_FWD, "correlation forward")
    g = r.standard_normal(ref.shape).astype(np.float32)
    out.backward(cu(g))
    gi1, gi2 = oracle.correlation_backward(f1, f2, g, 4, 1, 4, 1, 1)
    U.assert_close(host(t1.grad), gi1, U.RTOL_FWD, "correlation gradinput1")
    U.assert_close(host(t2.grad), gi2, U.RTOL_FWD, "correlation gradinput2")


@pytest.mark.parametrize("pad,k,md,s1,s2", [(3, 3, 6, 1, 2), (2, 1, 2, 1, 1), (6, 1, 4, 1, 1), (4, 3, 4, 1, 1), (8, 1, 8, 2, 2)])
def test_correlation_generic_configurations(lib, oracle, pad, k, md, s1, s2):
    r = U.rng(1701)
    B, C, H, W = 2, 6, 20, 27
    f1, f2 = U.image(r, B, C, H, W, "normal"), U.image(r, B, C, H, W, "normal")
    t1, t2 = cu(f1).requires_grad_(), cu(f2).requires_grad_()
    out = lib.Correlation(pad, k, md, s1, s2, 1)(t1, t2)
    ref = oracle.correlation_forward(f1, f2, pad, k, md, s1, s2)
    assert tuple(out.shape) == ref.shape
    U.assert_close(host(out), ref, U.RTOL_FWD, "correlation forward")
    if s1 == 1 and pad >= md:      # in-contract domain of the reference backward
        g = r.standard_normal(ref.shape).astype(np.float32)
        out.backward(cu(g))
        gi1, gi2 = oracle.correlation_backward(f1, f2, g, pad, k, md, s1, s2)
        U.assert_close(host(t1.grad), gi1, U.RTOL_FWD, "gradinput1")
        U.assert_close(host(t2.grad), gi2, U.RTOL_FWD, "gradinput2")


@pytest.mark.parametrize("B,C,H,W,pad,k,md,s1,s2", [
    (2, 32, 40, 64, 4, 1, 4, 1, 1),      # TMA path
    (8, 196, 18, 31, 4, 1, 4, 1, 1),     # PWC level 6 at 1080p: split-K, odd width (cp.async staging)
    (8, 128, 36, 62, 4, 1, 4, 1, 1),     # PWC level 5: row-padded TMA path
    (1, 5, 9, 13, 4, 1, 4, 1, 1),
    (3, 16, 72, 124, 4, 1, 4, 1, 1),
    (1, 4, 14, 18, 7, 3, 6, 1, 2),       # generic configuration: two launches inside
])
def test_correlation_both_directions_in_one_launch(lib, oracle, B, C, H, W, pad, k, md, s1, s2):
    """vfidkr_correlation_forward_pair: equal to two single calls (bit-identical unless the channel split differs), equal to
    the oracle, differentiable."""
    r = U.rng(1750 + C + W)
    f1, f2 = U.image(r, B, C, H, W, "normal"), U.image(r, B, C, H, W, "normal")
    t1, t2 = cu(f1).requires_grad_(), cu(f2).requires_grad_()
    mod = lib.Correlation(pad, k, md, s1, s2, 1)
    before = lib.launch_count()
    o12, o21 = mod.both_directions(t1, t2)
    torch.cuda.synchronize()
    pair_launches = lib.launch_count() - before
    a, b = mod(t1.detach(), t2.detach()), mod(t2.detach(), t1.detach())
    # same arithmetic per output; where the channels are split over work items (coarse levels) the pair launch may
    # choose another slice count than the single one, i.e. another summation order: equal to rounding, else bit-identical
    split = B * ((H + 7) // 8) * ((W + 31) // 32) <= 148
    if split:
        assert U.max_err(host(o12), host(a).astype(np.float64)) <= 2e-6 and U.max_err(host(o21), host(b).astype(np.float64)) <= 2e-6
    else:
        assert torch.equal(o12, a) and torch.equal(o21, b)
    if k == 1:
        assert pair_launches <= 3      # (row padding) + kernel + (split-K reduce): never two of each
    U.assert_close(host(o12), oracle.correlation_forward(f1, f2, pad, k, md, s1, s2), U.RTOL_FWD, "pair: corr(f1, f2)")
    U.assert_close(host(o21), oracle.correlation_forward(f2, f1, pad, k, md, s1, s2), U.RTOL_FWD, "pair: corr(f2, f1)")
    g12, g21 = r.standard_normal(tuple(o12.shape)).astype(np.float32), r.standard_normal(tuple(o12.shape)).astype(np.float32)
    (o12 * cu(g12)).sum().backward(retain_graph=True)
    (o21 * cu(g21)).sum().backward()
    a1, a2 = oracle.correlation_backward(f1, f2, g12, pad, k, md, s1, s2)
    b2, b1 = oracle.correlation_backward(f2, f1, g21, pad, k, md, s1, s2)
    U.assert_close(host(t1.grad), a1 + b1, U.RTOL_FWD, "pair: gradinput1")
    U.assert_close(host(t2.grad), a2 + b2, U.RTOL_FWD, "pair: gradinput2")


@pytest.mark.parametrize("B,C,H,W", [(2, 32, 40, 64), (1, 8, 12, 20), (8, 196, 18, 31), (8, 128, 36, 62), (3, 16, 72, 124),
                                     (1, 5, 9, 13), (2, 64, 33, 50), (1, 32, 96, 256)])
def test_correlation_tensor_core_path(lib, oracle, B, C, H, W):
    """correlation_tc.cu (tcgen05.mma kind::tf32, 3 x TF32 split, accumulator in tensor memory) against the oracle at the
    forward tolerance and against the SIMT kernel: ragged patches, odd widths, channel counts that are not a multiple of
    the 32-channel stage (zero-filled), the PWC level shapes."""
    r = U.rng(1790 + C + W)
    f1, f2 = U.image(r, B, C, H, W, "normal"), U.image(r, B, C, H, W, "normal")
    mod = lib.Correlation(4, 1, 4, 1, 1, 1)
    lib.debug_force_correlation_path("tensor")
    tc = mod(cu(f1), cu(f2))
    lib.debug_force_correlation_path("simt")
    simt = mod(cu(f1), cu(f2))
    lib.debug_force_correlation_path(None)
    ref = oracle.correlation_forward(f1, f2, 4, 1, 4, 1, 1)
    U.assert_close(host(tc), ref, U.RTOL_FWD, "tensor-core correlation vs oracle")
    U.assert_close(host(simt), ref, U.RTOL_FWD, "SIMT correlation vs oracle")
    print(f"tensor vs SIMT: {U.max_err(host(tc), host(simt).astype(np.float64)):.2e}; tensor vs oracle: {U.max_err(host(tc), ref):.2e}; "
          f"SIMT vs oracle: {U.max_err(host(simt), ref):.2e}")


@pytest.mark.parametrize("B,C,H,W", [(2, 32, 40, 64), (8, 196, 18, 31), (1, 5, 9, 13), (2, 64, 33, 50)])
def test_correlation_tensor_core_pair_launch(lib, oracle, B, C, H, W):
    """Both directions in one launch of the tensor-core kernel: bit-identical to its two single launches, one launch."""
    r = U.rng(1795 + C + W)
    f1, f2 = U.image(r, B, C, H, W, "normal"), U.image(r, B, C, H, W, "normal")
    mod = lib.Correlation(4, 1, 4, 1, 1, 1)
    lib.debug_force_correlation_path("tensor")
    before = lib.launch_count()
    o12, o21 = mod.both_directions(cu(f1), cu(f2))
    torch.cuda.synchronize()
    launches = lib.launch_count() - before
    a, b = mod(cu(f1), cu(f2)), mod(cu(f2), cu(f1))
    lib.debug_force_correlation_path(None)
    assert launches == 1
    assert torch.equal(o12, a) and torch.equal(o21, b)
    U.assert_close(host(o12), oracle.correlation_forward(f1, f2, 4, 1, 4, 1, 1), U.RTOL_FWD, "tensor pair: corr(f1, f2)")
    U.assert_close(host(o21), oracle.correlation_forward(f2, f1, 4, 1, 4, 1, 1), U.RTOL_FWD, "tensor pair: corr(f2, f1)")


def test_correlation_automatic_path_choice(lib):
    """The automatic rule (correlation.cu: use_tensor_path): small deep maps -- PWC levels 6 and 5 -- go to the tensor-core kernel
    (one launch, no split-K reduction), larger ones to the FFMA kernel; both agree to rounding."""
    mod = lib.Correlation(4, 1, 4, 1, 1, 1)
    for (B, C, H, W), tensor in (((8, 196, 18, 31), True), ((8, 128, 36, 62), True), ((8, 96, 72, 124), False), ((2, 32, 40, 64), False)):
        a, b = torch.randn(B, C, H, W, device="cuda"), torch.randn(B, C, H, W, device="cuda")
        before = lib.launch_count()
        auto = mod(a, b)
        n = lib.launch_count() - before
        lib.debug_force_correlation_path("tensor" if tensor else "simt")
        forced = mod(a, b)
        lib.debug_force_correlation_path(None)
        assert torch.equal(auto, forced), (B, C, H, W)
        if tensor:
            assert n == 1


@pytest.mark.parametrize("name", ["corr_pwc", "corr_pwc_c196", "corr_wide_splitk", "corr_wide_tiled"])
def test_correlation_tensor_core_path_reproduces_reference_fixtures(lib, name):
    """The golden fixtures of the reference's own correlation kernel, through the tensor-core path."""
    import test_golden as TG
    ins, ref = TG.load(name)
    lib.debug_force_correlation_path("tensor")
    c = TG.G.CASES[name]
    out = lib.Correlation(c["pad"], c["k"], c["md"], c["s1"], c["s2"], 1)(cu(ins["input1"]), cu(ins["input2"]))
    lib.debug_force_correlation_path(None)
    TG.compare(name, dict(out=host(out)), {"out": ref["out"]}, ins)


def test_correlation_of_ones_on_gpu(lib):
    f = torch.ones(1, 8, 10, 12, device="cuda")
    out = lib.Correlation(4, 1, 4, 1, 1, 1)(f, f)
    assert (out[0, 40] == 1).all()                                          # KAT 7
    assert out[0, 0, 0, 0] == 0 and out[0, 0, 4, 4] == 1


# ------------------------------------------------------------------------------ full-size properties
def test_full_size_1080p_properties(lib, oracle):
    """BASELINE config 4 shape (1080p padded to 1152x1984, B=8): size-independent properties, plus the
    oracle on a crop-free sub-batch it can finish in seconds."""
    r = U.rng(1800)
    B, C, H, W = 8, 3, 1152, 1984
    I = torch.rand(B, C, H, W, device="cuda")
    fl = (torch.randn(B, 2, H, W, device="cuda") * 4).clamp_(-20, 20)
    ft = torch.softmax(torch.randn(B, 16, H, W, device="cuda"), dim=1)
    out = lib.FilterInterpolationModule()(I, fl, ft)
    # (1) linearity in the image: FI(a*I + J) = a*FI(I) + FI(J) for in-range and copied pixels alike
    J = torch.rand_like(I)
    lhs = lib.FilterInterpolationModule()(2.5 * I + J, fl, ft)
    rhs = 2.5 * out + lib.FilterInterpolationModule()(J, fl, ft)
    assert (lhs - rhs).abs().max().item() < 2e-5
    # (2) a convex filter (softmax) and convex bilinear blend keep the output inside the input range
    assert out.min().item() >= -1e-6 and out.max().item() <= 1 + 1e-6
    # (3) identity: zero flow + one-hot centre tap
    onehot = torch.zeros_like(ft)
    onehot[:, 5] = 1
    assert torch.equal(lib.FilterInterpolationModule()(I, torch.zeros_like(fl), onehot), I)
    # (4) oracle on the first batch item
    ref = oracle.fi_forward("ori", host(I[:1]), host(fl[:1]), host(ft[:1]))
    U.assert_close(host(out[:1]), ref, U.RTOL_FWD, "1080p forward, item 0")
    # (5) projection: zero flow -> zero output, count 4 in the interior; mass conservation of the splat
    z = torch.zeros(B, 2, H, W, device="cuda")
    assert not lib.FlowProjectionModule(False)(z).any()
    d = torch.rand(B, 1, H, W, device="cuda") * 0.9 + 0.1
    po = lib.DepthFlowProjectionModule(True)(fl, d)
    assert torch.isfinite(po).all()
    ref, _ = oracle.flowprojection_forward(host(fl[:1]), host(d[:1]), 0)
    U.assert_close(host(po[:1]), ref, U.RTOL_ATOMIC, "1080p depth projection, item 0")


# ------------------------------------------------------------------------------ TMA path vs direct path
def test_fi_ori_tma_and_direct_paths_agree(lib, oracle, monkeypatch):
    """The TMA-streamed persistent kernel (W % 4 == 0, aligned) and the direct kernel must agree; edge
    tiles (W not a multiple of the 64-wide tile, H not a multiple of 4) are covered."""
    r = U.rng(1900)
    for (B, C, H, W) in [(2, 3, 37, 132), (1, 3, 256, 448), (3, 5, 9, 68), (1, 196, 8, 64)]:
        I, fl, ft = U.image(r, B, C, H, W), U.flow(r, B, H, W, "stress"), U.filt(r, B, 4, H, W)
        lib.debug_force_forward_path("direct")
        a = lib.FilterInterpolationModule()(cu(I), cu(fl), cu(ft))
        lib.debug_force_forward_path(None)
        b = lib.FilterInterpolationModule()(cu(I), cu(fl), cu(ft))
        # same arithmetic, slightly different FMA grouping: agree to fp32 rounding, and each with the oracle
        assert U.max_err(host(a), host(b).astype(np.float64)) < 2e-6, (B, C, H, W)
        U.assert_close(host(b), oracle.fi_forward("ori", I, fl, ft), U.RTOL_FWD, "tma path vs oracle")
        U.assert_close(host(a), oracle.fi_forward("ori", I, fl, ft), U.RTOL_FWD, "direct path vs oracle")
    # an unaligned view (offset by one float) must silently take the direct path and still be right
    B, C, H, W = 1, 3, 16, 64
    I, fl, ft = U.image(r, B, C, H, W), U.flow(r, B, H, W, "unit"), U.filt(r, B, 4, H, W)
    buf = torch.empty(ft.size + 1, device="cuda")
    view = buf[1:].view(B, 16, H, W)
    view.copy_(cu(ft))
    out = lib.FilterInterpolationModule()(cu(I), cu(fl), view)
    U.assert_close(host(out), oracle.fi_forward("ori", I, fl, ft), U.RTOL_FWD, "unaligned filter tensor")


# ------------------------------------------------------------------------------ strip kernel (rolling window)
def big_flow(r, B, H, W, kind):
    """Flows that exercise the rolling shared-memory window of fi_strip.cu."""
    if kind == "uniform_motion":      # large constant motion: the window must be re-based away from the strip
        f = np.zeros((B, 2, H, W), np.float32)
        f[:, 0] = 37.3
        f[:, 1] = -21.6
        f += 0.3 * r.standard_normal((B, 2, H, W)).astype(np.float32)
    elif kind == "shear":             # x-displacement grows with y: periodic re-basing inside a strip
        yy = np.arange(H, dtype=np.float32)[None, :, None]
        f = np.zeros((B, 2, H, W), np.float32)
        f[:, 0] = 0.35 * yy - 20
        f[:, 1] = 6 * np.sin(yy / 9.0)
    elif kind == "wild":              # +-40 px i.i.d.: most tiles do not fit the window (global fallback)
        f = np.clip(r.standard_normal((B, 2, H, W)) * 25.0, -60, 60).astype(np.float32)
    elif kind == "jump_back":         # y-flow that jumps up by a block every 16 rows: window rows must be re-loaded
        yy = np.arange(H)[None, :, None]
        f = np.zeros((B, 2, H, W), np.float32)
        f[:, 1] = np.where((yy // 16) % 2 == 0, 18.0, -18.0)
        f[:, 0] = 2.5
    else:
        raise KeyError(kind)
    return np.ascontiguousarray(f, dtype=np.float32)


STRIP_CASES = [  # B, C, H, W, flow
    (1, 3, 256, 448, "gauss"),            # BASELINE config 1
    (2, 3, 131, 200, "stress"),           # ragged: W % 64 != 0, H % 4 != 0, out-of-range pixels
    (1, 3, 96, 256, "smooth"),
    (1, 3, 180, 192, "uniform_motion"),
    (1, 3, 200, 160, "shear"),
    (1, 3, 120, 224, "wild"),
    (1, 3, 160, 288, "jump_back"),
    (2, 1, 64, 160, "gauss"),             # narrowest width the path accepts (128-column instantiation)
    (1, 2, 70, 196, "unit"),
    (1, 4, 90, 164, "gauss"),
    (3, 3, 8, 640, "gauss"),              # very short strips: many strip changes per CTA
]


@pytest.mark.parametrize("B,C,H,W,fk", STRIP_CASES)
def test_fi_ori_strip_kernel(lib, oracle, monkeypatch, B, C, H, W, fk):
    """fi_strip.cu (rolling shared-memory window fed by TMA) against the oracle and against the direct kernel."""
    r = U.rng(2100 + H + W)
    I = U.image(r, B, C, H, W)
    fl = big_flow(r, B, H, W, fk) if fk in ("uniform_motion", "shear", "wild", "jump_back") else U.flow(r, B, H, W, fk)
    ft = U.filt(r, B, 4, H, W, "uniform")
    lib.debug_force_forward_path("strip")
    a = lib.FilterInterpolationModule()(cu(I), cu(fl), cu(ft))
    lib.debug_force_forward_path("direct")
    b = lib.FilterInterpolationModule()(cu(I), cu(fl), cu(ft))
    lib.debug_force_forward_path(None)
    ref = oracle.fi_forward("ori", I, fl, ft)
    U.assert_close(host(a), ref, U.RTOL_FWD, f"strip kernel vs oracle ({fk})")
    U.assert_close(host(b), ref, U.RTOL_FWD, f"direct kernel vs oracle ({fk})")
    assert U.max_err(host(a), host(b).astype(np.float64)) < 2e-6


def test_fi_ori_strip_kernel_repeated_launches_are_deterministic(lib):
    """The forward has no atomics: two launches must agree bit for bit (also shakes out pipeline races)."""
    B, C, H, W = 4, 3, 512, 960
    g = torch.Generator(device="cuda").manual_seed(5)
    I = torch.rand(B, C, H, W, device="cuda", generator=g)
    lo = (torch.randn(B, 2, H // 4, W // 4, device="cuda", generator=g) * 4).clamp_(-20, 20)
    fl = torch.nn.functional.interpolate(lo, scale_factor=4, mode="bilinear", align_corners=False).contiguous()
    ft = torch.softmax(torch.randn(B, 16, H, W, device="cuda", generator=g), 1)
    first = lib.FilterInterpolationModule()(I, fl, ft).clone()
    for _ in range(5):
        assert torch.equal(lib.FilterInterpolationModule()(I, fl, ft), first)


DKR_STRIP_CASES = [  # B, C, H, W, flow kind, offset amplitude
    (2, 3, 131, 200, "stress", 0.45),        # ragged, out-of-range pixels
    (1, 3, 96, 256, "smooth", 0.95),         # offsets up to the edge of the in-contract domain
    (1, 3, 180, 192, "uniform_motion", 0.45),
    (1, 3, 120, 224, "wild", 0.45),          # most tiles without a window (global path)
    (1, 3, 64, 160, "gauss", 3.0),           # wild offsets: per-tap global fallback next to window taps
    (1, 1, 40, 164, "gauss", 0.45),
    (1, 4, 50, 196, "unit", 0.45),
    (2, 2, 8, 320, "gauss", 1.5),
]


@pytest.mark.parametrize("variant", ["dkr", "deforconv", "nofilterwithdeforconv"])
@pytest.mark.parametrize("B,C,H,W,fk,amp", DKR_STRIP_CASES)
def test_fi_dkr_strip_kernel(lib, oracle, monkeypatch, variant, B, C, H, W, fk, amp):
    """fi_strip_dkr.cu (bilinear samples from the rolling window, tap-row sub-stages) against the oracle and the
    direct kernel; amplitudes > 1 mix window taps with the clamp-to-plane global fallback."""
    r = U.rng(2300 + H + W + len(variant))
    I = U.image(r, B, C, H, W)
    fl = big_flow(r, B, H, W, fk) if fk in ("uniform_motion", "shear", "wild", "jump_back") else U.flow(r, B, H, W, fk)
    ft, off = U.filt(r, B, 4, H, W, "uniform"), U.offsets(r, B, 4, H, W, amp)
    args = (I, fl, off) if variant == "nofilterwithdeforconv" else (I, fl, ft, off)
    run = lambda: lib.FilterInterpolationModule(variant)(*map(cu, args))
    a = run()
    lib.debug_force_forward_path("direct")
    b = run()
    lib.debug_force_forward_path(None)
    ref = oracle.fi_forward(variant, *args)
    # an offset of exactly +-amp can land a sample on a pixel boundary where the two paths round the fraction
    # identically (same fp32 index arithmetic), so both must match the oracle at the forward tolerance
    U.assert_close(host(a), ref, U.RTOL_FWD, f"{variant} strip kernel vs oracle")
    U.assert_close(host(b), ref, U.RTOL_FWD, f"{variant} direct kernel vs oracle")
    assert U.max_err(host(a), host(b).astype(np.float64)) < 3e-6


@pytest.mark.parametrize("variant", ["ori", "dkr"])
def test_fi_strip_window_is_not_reused_across_work_items(lib, monkeypatch, variant):
    """Regression (found as a 1-in-5 flake of the full-size linearity check): a CTA that starts a new work item
    (another frame / strip segment) must re-base its rolling window even when the item's first tiles never touch
    the window.  Here the first tile row of every 36-tile segment is out of range (copy path, no window) and the
    flow points 6 rows up, so without the re-base the next tile finds the previous item's rows "resident" whenever
    that item was the segment above in the same strip of a different frame.  Full config-4 size: every CTA handles
    several items; compared with the direct kernel (no shared-memory window)."""
    B, C, H, W = 8, 3, 1152, 1984
    g = torch.Generator(device="cuda").manual_seed(77)
    I = torch.rand(B, C, H, W, device="cuda", generator=g)
    fl = torch.zeros(B, 2, H, W, device="cuda")
    fl[:, 1] = -6.0
    rows = torch.arange(H, device="cuda")
    for segt in (36, 32, 24, 18, 16, 12, 9, 8):                  # segment lengths the launcher may choose at this size
        fl[:, 0, (rows // 4) % segt == 0, :] = float(W)           # |fx| >= W/2: out of range (:2735-2736)
    ft = torch.softmax(torch.randn(B, 16, H, W, device="cuda", generator=g), 1)
    off = (torch.rand(B, 32, H, W, device="cuda", generator=g) - 0.5) * 0.9
    args = (I, fl, ft) if variant == "ori" else (I, fl, ft, off)
    mod = lib.FilterInterpolationModule() if variant == "ori" else lib.FilterInterpolationModule(variant)
    lib.debug_force_forward_path("direct")
    ref = mod(*args)
    lib.debug_force_forward_path(None)
    for _ in range(4):   # the item-to-CTA assignment is dynamic: several draws
        out = mod(*args)
        assert (out - ref).abs().max().item() < 3e-6


# ------------------------------------------------------------------------------ backward on harder flows
@pytest.mark.parametrize("variant", ["ori", "dkr", "deforconv", "nofilterwithdeforconv"])
@pytest.mark.parametrize("B,C,H,W,fk", [(2, 3, 131, 200, "stress"), (1, 3, 96, 256, "smooth"), (1, 3, 120, 224, "wild"),
                                        (1, 5, 37, 70, "gauss"), (2, 1, 9, 33, "unit"), (1, 3, 64, 160, "compress")])
def test_fi_backward_all_families_hard_flows(lib, oracle, variant, B, C, H, W, fk):
    """Backward of every family (tap rows requested ahead of use, one pass over the taps, REDs for the image
    gradient only) against the oracle on ragged shapes, out-of-range pixels, channel chunks (C = 5) and
    "compress": a flow that maps runs of neighbouring pixels onto the same window origin (maximal RED collisions)."""
    r = U.rng(2500 + H + W + len(variant))
    I = U.image(r, B, C, H, W)
    if fk == "compress":
        xs = np.arange(W, dtype=np.float32)
        fl = np.zeros((B, 2, H, W), np.float32)
        fl[:, 0] = (np.floor(xs / 4) * 4 - xs + 0.25)[None, None, :]      # x + fx = 4*floor(x/4) + 0.25: 4 pixels per origin
        fl[:, 1] = 0.5
    elif fk in ("uniform_motion", "shear", "wild", "jump_back"):
        fl = big_flow(r, B, H, W, fk)
    else:
        fl = U.flow(r, B, H, W, fk)
    ft, off = U.filt(r, B, 4, H, W, "uniform"), U.offsets(r, B, 4, H, W, 0.45)
    g = r.standard_normal((B, C, H, W)).astype(np.float32)
    args = (I, fl, off) if variant == "nofilterwithdeforconv" else ((I, fl, ft) if variant == "ori" else (I, fl, ft, off))
    mod = lib.FilterInterpolationModule() if variant == "ori" else lib.FilterInterpolationModule(variant)
    ts = [cu(a).requires_grad_() for a in args]
    mod(*ts).backward(cu(g))
    got = [host(t.grad) for t in ts]
    if variant == "nofilterwithdeforconv":
        refs = list(oracle.fi_backward(variant, I, fl, off, None, g))[:3]
    elif variant == "ori":
        refs = list(oracle.fi_backward(variant, I, fl, ft, None, g))[:3]
    else:
        refs = list(oracle.fi_backward(variant, I, fl, ft, off, g))
    for k, (a, ref) in enumerate(zip(got, refs)):
        tol = U.RTOL_ATOMIC if k == 0 else U.RTOL_FWD   # gi1 is accumulated atomically, the rest is thread-private
        U.assert_close(a, ref, tol, f"{variant} backward, grad {k + 1}")


# ------------------------------------------------------------------------------ many-channel forward (fi_bigc.cu)
@pytest.mark.parametrize("B,C,H,W,fk", [(1, 196, 40, 128, "smooth"), (2, 9, 37, 132, "stress"), (1, 16, 64, 160, "wild"),
                                        (1, 7, 20, 64, "unit"), (1, 24, 50, 96, "uniform_motion"), (1, 12, 33, 68, "gauss")])
def test_fi_ori_many_channel_kernel(lib, oracle, monkeypatch, B, C, H, W, fk):
    """fi_bigc.cu (C > 4: image regions streamed channel group by channel group through shared memory, taps in
    registers) against the oracle and the direct kernel: staged tiles, oversized boxes (global fallback), channel
    counts that are not a multiple of the group size, ragged tiles, out-of-range pixels."""
    r = U.rng(2700 + C + H + W)
    I = U.image(r, B, C, H, W)
    fl = big_flow(r, B, H, W, fk) if fk in ("uniform_motion", "shear", "wild", "jump_back") else U.flow(r, B, H, W, fk)
    ft = U.filt(r, B, 4, H, W, "uniform")
    a = lib.FilterInterpolationModule()(cu(I), cu(fl), cu(ft))
    lib.debug_force_forward_path("direct")
    b = lib.FilterInterpolationModule()(cu(I), cu(fl), cu(ft))
    lib.debug_force_forward_path(None)
    ref = oracle.fi_forward("ori", I, fl, ft)
    U.assert_close(host(a), ref, U.RTOL_FWD, f"many-channel kernel vs oracle ({fk})")
    U.assert_close(host(b), ref, U.RTOL_FWD, f"direct kernel vs oracle ({fk})")
    assert U.max_err(host(a), host(b).astype(np.float64)) < 2e-6


# ------------------------------------------------------------------------------ BASELINE configs 2, 3 and 5 at full size
def test_config3_training_step_dkr_and_projections_batch16(lib, oracle):
    """BASELINE config 3 beyond the "_ori" warp: fwd+bwd of the 4-input DKR warp, FlowProjection (fillhole = 0)
    and DepthFlowProjection at B = 16, 256x448, against the oracle; upstream gradient = the output."""
    r = U.rng(1033)
    B, C, H, W = 16, 3, 256, 448
    I, fl, ft = U.image(r, B, C, H, W), U.flow(r, B, H, W, "gauss"), U.filt(r, B, 4, H, W)
    off, d = U.offsets(r, B, 4, H, W, 0.45), U.depth_inv(r, B, H, W)
    ts = [cu(a).requires_grad_() for a in (I, fl, ft, off)]
    out = lib.FilterInterpolationModule("dkr")(*ts)
    U.assert_close(host(out), oracle.fi_forward("dkr", I, fl, ft, off), U.RTOL_FWD, "DKR forward")
    out.backward(out.detach())
    refs = oracle.fi_backward("dkr", I, fl, ft, off, host(out))
    for k, (t, ref) in enumerate(zip(ts, refs)):
        U.assert_close(host(t.grad), ref, U.RTOL_ATOMIC if k == 0 else U.RTOL_FWD, f"DKR gradinput{k + 1}")
    for depth in (None, d):
        tf = cu(fl).requires_grad_()
        if depth is None:
            po = lib.FlowProjectionModule(True)(tf)
        else:
            td = cu(depth).requires_grad_()
            po = lib.DepthFlowProjectionModule(True)(tf, td)
        ref, cnt = oracle.flowprojection_forward(fl, depth, 0)
        U.assert_close(host(po), ref, U.RTOL_ATOMIC, "projection forward")
        po.backward(po.detach())
        gi1, gi2 = oracle.flowprojection_backward(fl, depth, cnt.astype(np.float32), ref.astype(np.float32), host(po))
        U.assert_close(host(tf.grad), gi1, U.RTOL_ATOMIC, "projection gradinput1")
        if depth is not None:
            U.assert_close(host(td.grad), gi2, U.RTOL_ATOMIC, "projection gradinput2 (depth)")


def test_config2_forward_chain_256x448(lib, oracle):
    """BASELINE config 2 (one forward at the Vimeo-90K shape, 256x448, batch 1) restricted to the path: the operators
    of one forward pass in network order -- PWC correlations at the five pyramid levels in both directions
    (PWCNet.py:72), the flow warp between levels (PWCNet.py:159-199), depth-aware projection of both flows with hole
    filling (networks/DAIN.py:306-330, inference), adaptive warping of both frames by the PROJECTED flows, in the "_ori"
    and the deformable-kernel-region form (DAIN.py:536-573), and the blend.  The convolutions between them are out of
    scope, so features, depths, filters and offsets are synthetic; every stage is checked against the oracle run on the
    same upstream tensors the GPU stage consumed."""
    r = U.rng(1002)
    B, H, W = 1, 256, 448
    corr = lib.Correlation(4, 1, 4, 1, 1, 1)
    for C, s in ((196, 64), (128, 32), (96, 16), (64, 8), (32, 4)):
        a, b = U.image(r, B, C, H // s, W // s, "normal"), U.image(r, B, C, H // s, W // s, "normal")
        for x, y in ((a, b), (b, a)):
            U.assert_close(host(corr(cu(x), cu(y))), oracle.correlation_forward(x, y, 4, 1, 4, 1, 1), U.RTOL_FWD, f"correlation C={C}")
    feat, flo = U.image(r, B, 32, H // 4, W // 4, "normal"), (U.flow(r, B, H // 4, W // 4, "gauss") * 0.25).astype(np.float32)
    U.assert_close(host(lib.pwc_warp(cu(feat), cu(flo))), oracle.pwc_warp_forward(feat, flo), U.RTOL_FWD, "PWC warp")

    frames = [U.image(r, B, 3, H, W) for _ in range(2)]
    flows = [(0.5 * U.flow(r, B, H, W, "gauss")).astype(np.float32) for _ in range(2)]   # time step 0.5
    depths = [U.depth_inv(r, B, H, W) for _ in range(2)]
    filts = [U.filt(r, B, 4, H, W) for _ in range(2)]
    offs = [U.offsets(r, B, 4, H, W, 0.45) for _ in range(2)]
    proj = []
    for fl, d in zip(flows, depths):
        p = lib.DepthFlowProjectionModule(False)(cu(fl), cu(d))      # requires_grad = False: holes are filled
        ref, _ = oracle.flowprojection_forward(fl, d, 1)
        U.assert_close(host(p), ref, U.RTOL_ATOMIC, "depth-aware projection with hole filling")
        proj.append(p)
    warped = []
    for I, p, ft, off in zip(frames, proj, filts, offs):
        w = lib.FilterInterpolationModule()(cu(I), p, cu(ft))
        U.assert_close(host(w), oracle.fi_forward("ori", I, host(p), ft), U.RTOL_FWD, "adaptive warp by the projected flow")
        wd = lib.FilterInterpolationModule("dkr")(cu(I), p, cu(ft), cu(off))
        U.assert_close(host(wd), oracle.fi_forward("dkr", I, host(p), ft, off), U.RTOL_FWD, "DKR warp by the projected flow")
        warped.append(w)
    out = lib.filter_interpolate_blend(cu(frames[0]), cu(frames[1]), proj[0], proj[1], cu(filts[0]), cu(filts[1]))
    U.assert_close(host(out), 0.5 * host(warped[0]).astype(np.float64) + 0.5 * host(warped[1]).astype(np.float64), U.RTOL_FWD, "blended frame")


def test_config5_4k_pair_properties(lib, oracle, monkeypatch):
    """BASELINE config 5 shape (4K padded to 2176x3904, one pair per step): the strip kernel against the direct kernel
    on the whole frame, the oracle on a band of rows, and size-independent projection properties."""
    B, C, H, W = 1, 3, 2176, 3904
    g = torch.Generator(device="cuda").manual_seed(55)
    I = torch.rand(B, C, H, W, device="cuda", generator=g)
    lo = (torch.randn(B, 2, H // 8, W // 8, device="cuda", generator=g) * 6).clamp_(-30, 30)
    fl = torch.nn.functional.interpolate(lo, scale_factor=8, mode="bilinear", align_corners=False).contiguous()
    ft = torch.softmax(torch.randn(B, 16, H, W, device="cuda", generator=g), 1)
    out = lib.FilterInterpolationModule()(I, fl, ft)
    lib.debug_force_forward_path("direct")
    ref = lib.FilterInterpolationModule()(I, fl, ft)
    lib.debug_force_forward_path(None)
    assert (out - ref).abs().max().item() < 3e-6
    assert out.min().item() >= -1e-6 and out.max().item() <= 1 + 1e-6        # convex filter, convex blend
    # oracle on the first 64 rows: a pixel's window may reach below the band, so only rows whose windows stay inside count
    band = 96
    o_ref = oracle.fi_forward("ori", host(I[:, :, :band]), host(fl[:, :, :band]), host(ft[:, :, :band]))
    safe = (torch.arange(band, device="cuda")[None, :, None] + fl[0, 1, :band] < band - 4).cpu().numpy()[0]
    got = host(out[:, :, :band])
    err = np.abs(got - o_ref) / (np.abs(o_ref) + np.abs(o_ref).max())
    assert err[:, :, safe].max() <= U.RTOL_FWD
    # projection: zero flow -> zero output and count 4 in the interior; count mass = 4 x in-range pixels
    z = torch.zeros(B, 2, H, W, device="cuda")
    assert not lib.FlowProjectionModule(False)(z).any()
    tf = fl.clone().requires_grad_()
    po = lib.FlowProjectionModule(True)(tf)
    assert torch.isfinite(po).all()
    ref_p, _ = oracle.flowprojection_forward(host(fl[:, :, :band]), None, 0)
    # rows of the band that no pixel from below the band can reach (|fy| <= 30)
    U.assert_close(host(po[:, :, :band - 32]), ref_p[:, :, :band - 32], U.RTOL_ATOMIC, "4K projection band")


# ------------------------------------------------------------------------------ host-side pair streaming
def test_pair_stream_matches_whole_batch(lib):
    """vfidkr_b200.PairStream (copies in / kernels / copies out of successive pairs on three streams) must return
    exactly what one whole-batch call returns, for any chunking, and leave the caller's stream ordered after it."""
    g = torch.Generator().manual_seed(9)
    n, C, H, W = 5, 3, 96, 224
    hd = {"frame": torch.rand(n, C, H, W, generator=g).pin_memory(),
          "flow": (torch.randn(n, 2, H, W, generator=g) * 3).pin_memory(),
          "filter": torch.softmax(torch.randn(n, 16, H, W, generator=g), 1).pin_memory(),
          "depth": (torch.rand(n, 1, H, W, generator=g) * 0.9 + 0.1).pin_memory()}
    fi, dp = lib.FilterInterpolationModule(), lib.DepthFlowProjectionModule(False)

    def step(d):
        return fi(d["frame"], d["flow"], d["filter"]), dp(d["flow"], d["depth"])
    with torch.no_grad():
        ref = [t.cpu() for t in step({k: v.cuda() for k, v in hd.items()})]
        for chunk in (1, 2, 5):
            outs = [torch.empty(n, C, H, W).pin_memory(), torch.empty(n, 2, H, W).pin_memory()]
            ps = lib.PairStream(torch.device("cuda", 0), step, pairs_per_chunk=chunk)
            for _ in range(3):                       # back-to-back runs reuse the streams
                ps.run(hd, outs)
            torch.cuda.current_stream().synchronize()   # the caller's stream was made to wait for the results
            assert torch.equal(outs[0], ref[0]), chunk
            # projection sums are atomically accumulated: order, hence rounding, may differ between calls
            assert (outs[1] - ref[1]).abs().max().item() <= 1e-4 * ref[1].abs().max().item(), chunk


# ------------------------------------------------------------------------------ two directions + blend
@pytest.mark.parametrize("B,C,H,W,w0,w2", [(2, 3, 96, 224, 0.5, 0.5),      # strip kernel (networks/DAIN.py:573)
                                            (1, 3, 37, 70, 0.7, 0.3),       # direct kernel, (1-t), t (DAIN_slowmotion.py:335)
                                            (1, 6, 20, 64, 0.25, 0.75)])    # C > 4: direct kernel in blend mode
def test_fi_blend_two_directions(lib, oracle, B, C, H, W, w0, w2):
    """filter_interpolate_blend = w0 * FI(ref0, ...) + w2 * FI(ref2, ...) in two launches with a blend epilogue
    (no intermediate frames), forward and backward against the oracle composed the same way."""
    r = U.rng(2900 + H + W)
    I0, I2 = U.image(r, B, C, H, W), U.image(r, B, C, H, W)
    f0, f2 = U.flow(r, B, H, W, "stress"), U.flow(r, B, H, W, "gauss")
    k0, k2 = U.filt(r, B, 4, H, W), U.filt(r, B, 4, H, W, "uniform")
    ts = [cu(a).requires_grad_() for a in (I0, I2, f0, f2, k0, k2)]
    out = lib.filter_interpolate_blend(*ts, w0, w2)
    ref = w0 * oracle.fi_forward("ori", I0, f0, k0) + w2 * oracle.fi_forward("ori", I2, f2, k2)
    U.assert_close(host(out), ref, U.RTOL_FWD, "blend forward")
    # the unfused composition gives the same numbers up to one rounding of the scaling
    with torch.no_grad():
        fi = lib.FilterInterpolationModule()
        unfused = w0 * fi(ts[0], ts[2], ts[4]) + w2 * fi(ts[1], ts[3], ts[5])
    assert (out.detach() - unfused).abs().max().item() < 1e-6
    g = r.standard_normal((B, C, H, W)).astype(np.float32)
    out.backward(cu(g))
    a = oracle.fi_backward("ori", I0, f0, k0, None, (g * np.float32(w0)).astype(np.float32))
    b = oracle.fi_backward("ori", I2, f2, k2, None, (g * np.float32(w2)).astype(np.float32))
    for t, refg, tol, name in ((ts[0], a[0], U.RTOL_ATOMIC, "ref0"), (ts[1], b[0], U.RTOL_ATOMIC, "ref2"),
                               (ts[2], a[1], U.RTOL_FWD, "offset0"), (ts[3], b[1], U.RTOL_FWD, "offset2"),
                               (ts[4], a[2], U.RTOL_FWD, "filter0"), (ts[5], b[2], U.RTOL_FWD, "filter2")):
        U.assert_close(host(t.grad), refg, tol, f"blend grad {name}")


def test_fi_blend_writes_into_a_channel_slice(lib):
    """out= : the blended frame goes straight into a channel slice of a wider tensor (the rectify-input concat of
    networks/DAIN.py:264-269), other channels untouched; strip kernel (W = 224) and direct kernel (W = 70)."""
    for (B, C, H, W) in [(2, 3, 96, 224), (2, 3, 37, 70)]:
        g = torch.Generator(device="cuda").manual_seed(H)
        I0, I2 = (torch.rand(B, C, H, W, device="cuda", generator=g) for _ in range(2))
        f0, f2 = ((torch.randn(B, 2, H, W, device="cuda", generator=g) * 3) for _ in range(2))
        k0, k2 = (torch.softmax(torch.randn(B, 16, H, W, device="cuda", generator=g), 1) for _ in range(2))
        with torch.no_grad():
            ref = lib.filter_interpolate_blend(I0, I2, f0, f2, k0, k2, 0.3, 0.7)
            cat = torch.full((B, 11, H, W), -5.0, device="cuda")
            ret = lib.filter_interpolate_blend(I0, I2, f0, f2, k0, k2, 0.3, 0.7, out=cat[:, 4:7])
        assert torch.equal(cat[:, 4:7], ref) and ret.data_ptr() == cat[:, 4:7].data_ptr()
        assert (cat[:, :4] == -5).all() and (cat[:, 7:] == -5).all()
    with pytest.raises(lib.VfidkrError):   # not differentiable through out=
        lib.filter_interpolate_blend(I0.requires_grad_(), I2, f0, f2, k0, k2, out=cat[:, 4:7])


@pytest.mark.parametrize("B,C,H,W", [(1, 196, 40, 96), (2, 12, 64, 192), (1, 7, 33, 70)])
def test_fi_many_channels_into_a_channel_slice(lib, oracle, B, C, H, W):
    """VERDICT r1 (missing 7): the many-channel kernel carries the scale / accumulate / batch-stride epilogue, so the warped
    context features (C = 196) go straight into their slices of the rectify-input concat (DAIN_slowmotion.py:167-181)."""
    r = U.rng(2950 + C)
    ctx0, ctx2 = U.image(r, B, C, H, W, "normal"), U.image(r, B, C, H, W, "normal")
    f0, f2 = U.flow(r, B, H, W, "gauss"), U.flow(r, B, H, W, "stress")
    k0, k2 = U.filt(r, B, 4, H, W), U.filt(r, B, 4, H, W, "uniform")
    cat = torch.full((B, 3 + 2 * C + 5, H, W), -5.0, device="cuda")
    n0 = lib.launch_count()
    with torch.no_grad():
        lib.filter_interpolate_into(cu(ctx0), cu(f0), cu(k0), cat[:, 3:3 + C])
        lib.filter_interpolate_into(cu(ctx2), cu(f2), cu(k2), cat[:, 3 + C:3 + 2 * C])
        # blend epilogue on the many-channel path: 0.25 * warp0 + 0.75 * warp2 accumulated in place
        acc = torch.empty((B, C, H, W), device="cuda")
        lib.filter_interpolate_into(cu(ctx0), cu(f0), cu(k0), acc, scale=0.25)
        lib.filter_interpolate_into(cu(ctx2), cu(f2), cu(k2), acc, scale=0.75, accumulate=True)
    assert lib.launch_count() - n0 == 4
    r0, r2 = oracle.fi_forward("ori", ctx0, f0, k0), oracle.fi_forward("ori", ctx2, f2, k2)
    U.assert_close(host(cat[:, 3:3 + C]), r0, U.RTOL_FWD, "ctx0 into its slice")
    U.assert_close(host(cat[:, 3 + C:3 + 2 * C]), r2, U.RTOL_FWD, "ctx2 into its slice")
    assert (cat[:, :3] == -5).all() and (cat[:, 3 + 2 * C:] == -5).all()
    U.assert_close(host(acc), 0.25 * r0 + 0.75 * r2, U.RTOL_FWD, "blend on the many-channel path")
    with pytest.raises(lib.VfidkrError):
        lib.filter_interpolate_into(cu(ctx0).requires_grad_(), cu(f0), cu(k0), cat[:, 3:3 + C])


# ------------------------------------------------------------------------------ PWCDCNet.warp
@pytest.mark.parametrize("ac", [True, False])
@pytest.mark.parametrize("B,C,H,W", [(2, 32, 64, 112), (1, 128, 18, 31), (3, 5, 17, 23), (1, 2, 1, 9), (1, 196, 8, 14)])
def test_pwc_warp(lib, oracle, B, C, H, W, ac):
    """pwc_warp (PWCNet/PWCNet.py:159-199: grid + flow, grid_sample with the reference's normalisation quirk, validity
    mask) forward and backward against the numpy restatement, and against PyTorch's own grid_sample on the GPU."""
    r = U.rng(3200 + H + W)
    x = r.standard_normal((B, C, H, W)).astype(np.float32)
    flo = (r.standard_normal((B, 2, H, W)) * 4).astype(np.float32)
    flo[:, :, 0, 0] = 0.0
    flo[:, 0, -1, -1] = 3.0 * W
    tx, tf = cu(x).requires_grad_(), cu(flo).requires_grad_()
    out = lib.pwc_warp(tx, tf, align_corners=ac)
    U.assert_close(host(out), oracle.pwc_warp_forward(x, flo, ac), U.RTOL_FWD, "pwc warp forward")
    g = r.standard_normal((B, C, H, W)).astype(np.float32)
    out.backward(cu(g))
    gx, gf = oracle.pwc_warp_backward(x, flo, g, ac)
    U.assert_close(host(tx.grad), gx, U.RTOL_ATOMIC, "pwc warp grad x")
    U.assert_close(host(tf.grad), gf, U.RTOL_ATOMIC, "pwc warp grad flow")
    # the reference's own code path on the same device
    with torch.no_grad():
        xx = torch.arange(W, device="cuda").view(1, 1, 1, W).expand(B, 1, H, W)
        yy = torch.arange(H, device="cuda").view(1, 1, H, 1).expand(B, 1, H, W)
        vg = torch.cat((xx, yy), 1).float() + cu(flo)
        vg = torch.stack([2.0 * vg[:, 0] / max(W - 1, 1) - 1.0, 2.0 * vg[:, 1] / max(H - 1, 1) - 1.0], 1).permute(0, 2, 3, 1)
        o = torch.nn.functional.grid_sample(cu(x), vg, align_corners=ac)
        m = torch.nn.functional.grid_sample(torch.ones_like(o), vg, align_corners=ac)
        m[m < 0.9999] = 0
        m[m > 0] = 1
        # torch's CUDA grid_sample contracts its coordinate arithmetic into FMAs, so samples near a knife edge of the
        # 0.9999 mask / the floor differ by a little more than the forward tolerance (seen: 1.02e-5); the oracle
        # comparison above is the parity criterion, this one only guards against a different FORMULA
        assert (out.detach() - o * m).abs().max().item() <= 5e-5 * max(1.0, float(np.abs(x).max()))


@pytest.mark.parametrize("B,h,w", [(2, 16, 28), (1, 9, 13), (1, 1, 1), (2, 64, 112)])
@pytest.mark.parametrize("with_depth", [False, True])
def test_flow_projection_from_lowres_flow_is_differentiable(lib, B, h, w, with_depth):
    """Gradients through flow_project_lowres (missing 7) equal those of the unfused chain
    scale -> nn.Upsample(x4, bilinear) -> (Depth)FlowProjection that networks/DAIN.py:306-308 + FlowProject build --
    PyTorch's own autograd through its own Upsample around this package's projection layer."""
    r = U.rng(3600 + h + w)
    lo = (r.standard_normal((B, 2, h, w)) * 0.4).astype(np.float32)
    d = U.depth_inv(r, B, 4 * h, 4 * w)
    g = r.standard_normal((B, 2, 4 * h, 4 * w)).astype(np.float32)
    t_lo, t_d = cu(lo).requires_grad_(), cu(d).requires_grad_()
    fused = lib.flow_project_lowres(t_lo, 20.0, 0.5, depth=t_d if with_depth else None, fillhole=False)
    fused.backward(cu(g))
    r_lo, r_d = cu(lo).requires_grad_(), cu(d).requires_grad_()
    up = torch.nn.Upsample(scale_factor=4, mode="bilinear", align_corners=False)(20.0 * r_lo * 0.5)
    ref = lib.DepthFlowProjectionModule(True)(up, r_d) if with_depth else lib.FlowProjectionModule(True)(up)
    ref.backward(cu(g))
    assert (fused - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item())
    U.assert_close(host(t_lo.grad), host(r_lo.grad).astype(np.float64), U.RTOL_ATOMIC, "grad wrt the low-resolution flow")
    if with_depth:
        U.assert_close(host(t_d.grad), host(r_d.grad).astype(np.float64), U.RTOL_ATOMIC, "grad wrt the depth")
    with pytest.raises(lib.VfidkrError):     # hole filling is inference-only, as in the reference
        lib.flow_project_lowres(t_lo, 20.0, 0.5, fillhole=True)


# ------------------------------------------------------------------------------ MinDepthFlowProjection
@pytest.mark.parametrize("B,H,W,fk", [(2, 37, 29, "stress"), (1, 256, 448, "gauss"), (2, 64, 96, "smooth"), (1, 1, 1, "unit")])
def test_mindepth_flow_projection(lib, oracle, B, H, W, fk):
    """Deterministic MinDepthFlowProjection against the restated rule: forward (with and without hole filling),
    backward, repeatability, and ties resolved towards the lowest pixel index."""
    r = U.rng(3300 + H + W)
    fl = U.flow(r, B, H, W, fk)
    d = U.depth_inv(r, B, H, W)
    d[:, :, ::3, ::5] = d[:, :, :1, :1]                         # plenty of exact ties
    d[:, :, 1::7, 2::9] = 0.0                                   # and sources that may not win
    for requires_grad in (True, False):
        tf, td = cu(fl).requires_grad_(requires_grad), cu(d).requires_grad_(requires_grad)
        out = lib.minDepthFlowProjectionModule(requires_grad)(tf, td)
        ref, cnt = oracle.mindepth_forward(fl, d, 0 if requires_grad else 1)
        U.assert_close(host(out), ref, U.RTOL_FWD, f"min-depth projection forward fillhole={not requires_grad}")
        again = lib.minDepthFlowProjectionModule(requires_grad)(tf, td)
        assert torch.equal(out, again)                           # no race: bit-identical on every launch
        if requires_grad:
            g = r.standard_normal((B, 2, H, W)).astype(np.float32)
            out.backward(cu(g))
            gi1, gi2 = oracle.mindepth_backward(fl, d, cnt.astype(np.float32), g)
            U.assert_close(host(tf.grad), gi1, U.RTOL_FWD, "min-depth gradinput1")
            assert not td.grad.any()


# ------------------------------------------------------------------------------ frame I/O boundary
@pytest.mark.parametrize("B,H,W", [(1, 480, 640), (2, 37, 129), (1, 128, 256), (1, 1080, 1920)])
def test_frame_io_is_bit_exact(lib, B, H, W):
    """uint8 HWC frames -> padded float CHW and back against the numpy / torch formulation of demo_MiddleBury.py:276-364."""
    r = U.rng(3400 + H)
    u8 = r.integers(0, 256, (B, H, W, 3), dtype=np.uint8)
    pad = lib.frames_to_padded(torch.from_numpy(u8).cuda())
    (top, Hp), (left, Wp) = lib.frame_padding(H), lib.frame_padding(W)
    x = torch.from_numpy(np.transpose(u8, (0, 3, 1, 2)).astype("float32") / 255.0)                    # :276
    ref = torch.nn.ReplicationPad2d([left, Wp - W - left, top, Hp - H - top])(x).numpy()               # :303-309
    assert tuple(pad.shape) == ref.shape and np.array_equal(host(pad), ref)
    # back: values beyond [0, 1], exact .5 cases and NaN-free noise
    y = (ref + r.standard_normal(ref.shape).astype(np.float32) * 0.3).astype(np.float32)
    y[0, 0, top, left:left + 4] = np.array([0.5 / 255, 1.5 / 255, 2.5 / 255, 254.5 / 255], np.float32)
    back = lib.padded_to_frames(torch.from_numpy(y).cuda(), H, W)
    yy = np.transpose(255.0 * y.clip(0, 1.0)[:, :, top:top + H, left:left + W], (0, 2, 3, 1))          # :350-351
    assert np.array_equal(host(back), np.round(yy).astype(np.uint8))                                  # :364
    # round trip of an untouched frame
    assert np.array_equal(host(lib.padded_to_frames(pad, H, W)), u8)


# ------------------------------------------------------------------------------ flow pre-processing fused
@pytest.mark.parametrize("B,h,w", [(2, 16, 24), (1, 64, 112), (1, 9, 7), (1, 1, 1)])
def test_flow_projection_from_lowres_flow(lib, oracle, B, h, w):
    """flow_upsample4 against nn.Upsample(x4, bilinear) of the scaled flow (networks/DAIN.py:306-308), and the fused
    projections against the unfused ones fed with that enlarged flow (FlowProjection and DepthFlowProjection)."""
    r = U.rng(3500 + h + w)
    lo = (r.standard_normal((B, 2, h, w)) * 0.4).astype(np.float32)
    t_lo = cu(lo)
    up = lib.flow_upsample4(t_lo, 20.0, 0.5)
    with torch.no_grad():
        ref = torch.nn.Upsample(scale_factor=4, mode="bilinear", align_corners=False)(20.0 * t_lo * 0.5)
    assert (up - ref).abs().max().item() <= 2e-6 * max(1.0, ref.abs().max().item())
    d = cu(U.depth_inv(r, B, 4 * h, 4 * w))
    with torch.no_grad():
        for depth in (None, d):
            for fill in (True, False):
                fused = lib.flow_project_lowres(t_lo, 20.0, 0.5, depth=depth, fillhole=fill)
                if depth is None:
                    unfused = lib.FlowProjectionModule(not fill)(up)
                else:
                    unfused = lib.DepthFlowProjectionModule(not fill)(up, depth)
                # same flow values bit for bit -> same cells; the sums are atomically accumulated in both
                assert (fused - unfused).abs().max().item() <= 1e-4 * max(1.0, unfused.abs().max().item())
