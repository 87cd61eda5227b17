"""Parity AT THE BENCHMARKED SIZE (BASELINE config 4: 1080p padded to 1152x1984, batch 8).

The kernels the bench and the operator table time -- the strip / rolling-window forwards, the many-channel kernel,
the persistent and split-K correlation kernels, every backward -- are held to the float64 oracle here on the very
shapes they are timed on: the GPU runs the whole batch (so the persistent kernels see the real work distribution) and
the oracle checks the first and the last batch item (the correlation: the whole batch at every PWC level).
Inputs are generated on the device from a seeded generator with the bench's own flow model (bench.scene_flow).
Tolerances: 1e-5 forward and thread-private gradients, 1e-4 atomically accumulated tensors, both criteria of
tests/util.py (normalised and per-element)."""
import numpy as np
import pytest
import torch

import util as U

pytestmark = pytest.mark.gpu

B, H, W = 8, 1152, 1984
ITEMS = (0, B - 1)


def host(t):
    return t.detach().cpu().numpy()


def gen(seed):
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    return g


def inputs(seed, C=3, batch=B, offsets=False, neg_offsets=False):
    import bench
    g = gen(seed)
    d = dict(I=torch.rand((batch, C, H, W), generator=g, device="cuda"),
             fl=bench.scene_flow(torch, g, torch.device("cuda"), batch, H, W),
             ft=torch.softmax(torch.randn((batch, 16, H, W), generator=g, device="cuda"), dim=1))
    if offsets:
        off = (torch.rand((batch, 32, H, W), generator=g, device="cuda") - 0.5) * 0.9
        d["off"] = -(off.abs() + 0.01) if neg_offsets else off
    return d


def sl(t, i):
    return host(t[i:i + 1].contiguous())


VARIANTS = {"dkr": "FilterInterpolationLayerDKR", "deforconv": "FilterInterpolationLayerDeforConv",
            "nofilterwithdeforconv": "FilterInterpolationLayerNoFilterWithDeforConv"}


@pytest.mark.parametrize("variant", sorted(VARIANTS))
def test_fullsize_dkr_forward(lib, oracle, variant):
    """The three deformable-kernel-region forwards (fi_strip_dkr.cu) at 8 x 3 x 1152 x 1984."""
    d = inputs(4100 + len(variant), offsets=True)
    layer = getattr(lib, VARIANTS[variant])
    with torch.no_grad():
        if variant == "nofilterwithdeforconv":
            out = layer.apply(d["I"], d["fl"], d["off"])
        else:
            out = layer.apply(d["I"], d["fl"], d["ft"], d["off"])
    for i in ITEMS:
        if variant == "nofilterwithdeforconv":
            ref = oracle.fi_forward(variant, sl(d["I"], i), sl(d["fl"], i), sl(d["off"], i), None)
        else:
            ref = oracle.fi_forward(variant, sl(d["I"], i), sl(d["fl"], i), sl(d["ft"], i), sl(d["off"], i))
        U.assert_close(sl(out, i), ref, U.RTOL_FWD, f"{variant} forward 1080p item {i}")


def test_fullsize_ori_forward_both_items_and_rough_flow(lib, oracle):
    """fi_strip.cu on the bench flow (first and last item) and on the operator table's rough quarter-resolution flow."""
    d = inputs(4200)
    with torch.no_grad():
        out = lib.FilterInterpolationLayer.apply(d["I"], d["fl"], d["ft"])
        g = gen(4201)
        up4 = torch.nn.functional.interpolate((torch.randn((B, 2, H // 4, W // 4), generator=g, device="cuda") * 4).clamp_(-20, 20),
                                              scale_factor=4, mode="bilinear", align_corners=False).contiguous()
        out_up4 = lib.FilterInterpolationLayer.apply(d["I"], up4, d["ft"])
    for i in ITEMS:
        U.assert_close(sl(out, i), oracle.fi_forward("ori", sl(d["I"], i), sl(d["fl"], i), sl(d["ft"], i)), U.RTOL_FWD,
                       f"ori forward 1080p item {i}")
    i = B // 2
    U.assert_close(sl(out_up4, i), oracle.fi_forward("ori", sl(d["I"], i), sl(up4, i), sl(d["ft"], i)), U.RTOL_FWD,
                   "ori forward 1080p, up4 flow")


@pytest.mark.timeout(120)
@pytest.mark.parametrize("C,Wc", [(4, W), (3, 1920), (2, 1152), (4, 448), (3, 176)])
def test_fullsize_ori_forward_other_channel_counts_and_widths(lib, oracle, C, Wc):
    """The strip kernel's other instantiations on many tiles per CTA: 4 channels run with fewer filter stages than the
    look-ahead distance (2 with 144-column tiles, 3 with 128) -- the configuration in which the producer's "tile done" ring
    aliased tiles t-4 and t-1 and the pipeline dead-locked at this size (round 2; small shapes never showed it) -- and
    other widths: ragged last strips (1920, 448) and one below the 144-column kernel's window (176: 128-column kernel)."""
    import bench
    Bc = 4
    g = gen(4300 + C + Wc)
    I = torch.rand((Bc, C, H, Wc), generator=g, device="cuda")
    fl = bench.scene_flow(torch, g, torch.device("cuda"), Bc, H, Wc)
    ft = torch.softmax(torch.randn((Bc, 16, H, Wc), generator=g, device="cuda"), dim=1)
    with torch.no_grad():
        lib.debug_force_forward_path("strip")
        out = lib.FilterInterpolationLayer.apply(I, fl, ft)
        lib.debug_force_forward_path("direct")
        direct = lib.FilterInterpolationLayer.apply(I, fl, ft)
        lib.debug_force_forward_path(None)
    torch.cuda.synchronize()
    assert U.max_err(host(out), host(direct).astype(np.float64)) < 2e-6
    i = Bc - 1
    U.assert_close(sl(out, i), oracle.fi_forward("ori", sl(I, i), sl(fl, i), sl(ft, i)), U.RTOL_FWD, f"ori forward C={C} W={Wc}")


@pytest.mark.parametrize("variant", ["ori", "dkr", "deforconv"])
def test_fullsize_fi_backward(lib, oracle, variant):
    """FilterInterpolation backward at 8 x 3 x 1152 x 1984 through the C ABI (what the operator table times)."""
    from vfidkr_b200 import _lib
    from vfidkr_b200._common import ptr, stream_ptr
    dkr = variant != "ori"
    # negative offsets keep every deformed tap of the border pixels inside the plane (tests/golden_cases.py)
    d = inputs(4300 + len(variant), offsets=dkr, neg_offsets=True)
    I, fl, ft = d["I"], d["fl"], d["ft"]
    g = torch.randn((B, 3, H, W), generator=gen(4310), device="cuda")
    gi1, gi2, gi3 = torch.empty_like(I), torch.empty_like(fl), torch.empty_like(ft)
    sp = stream_ptr(I.device)
    if dkr:
        off = d["off"]
        gi4 = torch.empty_like(off)
        _lib.call(f"vfidkr_filterinterpolation_backward_{variant}", ptr(I), ptr(fl), ptr(ft), ptr(off), ptr(g),
                  ptr(gi1), ptr(gi2), ptr(gi3), ptr(gi4), B, 3, H, W, 4, sp)
    else:
        _lib.call("vfidkr_filterinterpolation_backward_ori", ptr(I), ptr(fl), ptr(ft), ptr(g), ptr(gi1), ptr(gi2), ptr(gi3),
                  B, 3, H, W, 4, sp)
    torch.cuda.synchronize()
    for i in ITEMS:
        r1, r2, r3, r4 = oracle.fi_backward(variant, sl(I, i), sl(fl, i), sl(ft, i), sl(d["off"], i) if dkr else None, sl(g, i))
        U.assert_close(sl(gi1, i), r1, U.RTOL_ATOMIC, f"{variant} backward 1080p item {i}: gradinput1 (scatter)")
        U.assert_close(sl(gi2, i), r2, U.RTOL_FWD, f"{variant} backward 1080p item {i}: gradinput2")
        U.assert_close(sl(gi3, i), r3, U.RTOL_FWD, f"{variant} backward 1080p item {i}: gradinput3")
        if dkr:
            U.assert_close(sl(gi4, i), r4, U.RTOL_FWD, f"{variant} backward 1080p item {i}: gradinput4")


def test_fullsize_depthflowprojection_forward_fill_and_backward(lib, oracle):
    g = gen(4400)
    import bench
    fl = bench.scene_flow(torch, g, torch.device("cuda"), B, H, W).requires_grad_()
    dep = (torch.rand((B, 1, H, W), generator=g, device="cuda") * 0.9 + 0.1).requires_grad_()
    go = torch.randn((B, 2, H, W), generator=g, device="cuda")
    with torch.no_grad():
        filled = lib.DepthFlowProjectionModule(False)(fl.detach(), dep.detach())     # inference: hole filling
    out = lib.DepthFlowProjectionModule(True)(fl, dep)
    out.backward(go)
    for i in ITEMS:
        f, d_ = sl(fl, i), sl(dep, i)
        ref_fill, _ = oracle.flowprojection_forward(f, d_, 1)
        U.assert_close(sl(filled, i), ref_fill, U.RTOL_ATOMIC, f"depth projection + hole filling 1080p item {i}")
        ref_out, ref_cnt = oracle.flowprojection_forward(f, d_, 0)
        U.assert_close(sl(out, i), ref_out, U.RTOL_ATOMIC, f"depth projection 1080p item {i}")
        r1, r2 = oracle.flowprojection_backward(f, d_, ref_cnt, ref_out, sl(go, i))
        U.assert_close(sl(fl.grad, i), r1, U.RTOL_ATOMIC, f"depth projection backward 1080p item {i}: gradinput1")
        U.assert_close(sl(dep.grad, i), r2, U.RTOL_ATOMIC, f"depth projection backward 1080p item {i}: gradinput2")


@pytest.mark.parametrize("C,s", [(196, 64), (128, 32), (96, 16), (64, 8), (32, 4)])
def test_fullsize_correlation_levels(lib, oracle, C, s):
    """The five PWC-Net pyramid levels of a 1080p batch of 8 (18x31 ... 288x496): whole batch against the oracle --
    the split-K path (levels 6, 5), the row-padded TMA path (level 5) and the persistent kernel (levels 4-2)."""
    g = gen(4500 + C)
    a = torch.randn((B, C, H // s, W // s), generator=g, device="cuda")
    b = torch.randn((B, C, H // s, W // s), generator=g, device="cuda")
    with torch.no_grad():
        out = lib.Correlation(4, 1, 4, 1, 1, 1)(a, b)
    U.assert_close(host(out), oracle.correlation_forward(host(a), host(b), 4, 1, 4, 1, 1), U.RTOL_FWD,
                   f"correlation level C={C} {H // s}x{W // s} B=8")


@pytest.mark.parametrize("C,s", [(64, 8), (32, 4)])
def test_fullsize_correlation_backward(lib, oracle, C, s):
    g = gen(4600 + C)
    h, w = H // s, W // s
    a = torch.randn((B, C, h, w), generator=g, device="cuda").requires_grad_()
    b = torch.randn((B, C, h, w), generator=g, device="cuda").requires_grad_()
    go = torch.randn((B, 81, h, w), generator=g, device="cuda")
    lib.Correlation(4, 1, 4, 1, 1, 1)(a, b).backward(go)
    for i in ITEMS:
        r1, r2 = oracle.correlation_backward(sl(a, i), sl(b, i), sl(go, i), 4, 1, 4, 1, 1)
        U.assert_close(sl(a.grad, i), r1, U.RTOL_FWD, f"correlation backward C={C} item {i}: gradinput1")
        U.assert_close(sl(b.grad, i), r2, U.RTOL_FWD, f"correlation backward C={C} item {i}: gradinput2")


def test_fullsize_many_channel_forward_c196(lib, oracle):
    """fi_bigc.cu on the 196-channel context tensor of DAIN_slowmotion at 1152 x 1984 (one pair: 1.8 GB in, 1.8 GB out)."""
    d = inputs(4700, C=196, batch=1)
    with torch.no_grad():
        out = lib.FilterInterpolationLayer.apply(d["I"], d["fl"], d["ft"])
    I, fl, ft = host(d["I"]), host(d["fl"]), host(d["ft"])
    got = host(out)
    del d, out
    torch.cuda.empty_cache()
    # the oracle returns float64 (3.6 GB at this size): check it in four channel groups
    for c0 in range(0, 196, 49):
        ref = oracle.fi_forward("ori", np.ascontiguousarray(I[:, c0:c0 + 49]), fl, ft)
        U.assert_close(got[:, c0:c0 + 49], ref, U.RTOL_FWD, f"ori forward C=196 1080p, channels {c0}..{c0 + 48}")
