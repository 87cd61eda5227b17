#!/usr/bin/env python
"""Runs one operator a few times at the 1080p batch-8 shape -- the command line ncu profiles.

    python tools/run_op.py fi_ori_fwd [--iters 3] [--B 8] [--C 3]
"""
import argparse
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import vfidkr_b200 as V
from vfidkr_b200 import _lib
from vfidkr_b200._common import ptr, stream_ptr

ap = argparse.ArgumentParser()
ap.add_argument("op")
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--B", type=int, default=8)
ap.add_argument("--C", type=int, default=3)
ap.add_argument("--H", type=int, default=1152)
ap.add_argument("--W", type=int, default=1984)
ap.add_argument("--flow", default="up4", help="scene (bench flow) | up4 | gauss (iid per pixel) | smooth | zero")
a = ap.parse_args()
dev = torch.device("cuda", 0)
B, C, H, W = a.B, a.C, a.H, a.W
torch.manual_seed(0)
I = torch.rand(B, C, H, W, device=dev)
if a.flow == "up4":
    fl = torch.nn.functional.interpolate((torch.randn(B, 2, H // 4, W // 4, device=dev) * 4).clamp_(-20, 20),
                                         scale_factor=4, mode="bilinear", align_corners=False).contiguous()
elif a.flow == "scene":
    from bench import scene_flow
    gen = torch.Generator(device=dev)
    gen.manual_seed(77)
    fl = scene_flow(torch, gen, dev, B, H, W)
elif a.flow == "gauss":
    fl = (torch.randn(B, 2, H, W, device=dev) * 4).clamp_(-20, 20)
elif a.flow == "smooth":
    yy, xx = torch.meshgrid(torch.linspace(0, 6, H, device=dev), torch.linspace(0, 6, W, device=dev), indexing="ij")
    fl = torch.stack([6 * torch.sin(xx) + 3 * torch.cos(yy), 5 * torch.cos(0.7 * xx) - 3 * torch.sin(yy)], 0)[None].repeat(B, 1, 1, 1).contiguous()
else:
    fl = torch.zeros(B, 2, H, W, device=dev)
ft = torch.softmax(torch.randn(B, 16, H, W, device=dev), 1)
off = (torch.rand(B, 32, H, W, device=dev) - 0.5) * 0.9
dep = torch.rand(B, 1, H, W, device=dev) * 0.9 + 0.1
g = torch.randn(B, C, H, W, device=dev)
sp = stream_ptr(dev)
gi1, gi2, gi3, gi4 = torch.empty_like(I), torch.empty_like(fl), torch.empty_like(ft), torch.empty_like(off)

out = torch.empty_like(I)
ops = {
    "fi_ori_fwd": lambda: _lib.call("vfidkr_filterinterpolation_forward_ori", ptr(I), ptr(fl), ptr(ft), ptr(out), B, C, H, W, 4, sp),
    "fi_ori_blend": lambda: (_lib.call("vfidkr_filterinterpolation_forward_ori_blend", ptr(I), ptr(fl), ptr(ft), ptr(out), B, C, H, W, 4, 0.5, 0, 0, sp),
                             _lib.call("vfidkr_filterinterpolation_forward_ori_blend", ptr(I), ptr(fl), ptr(ft), ptr(out), B, C, H, W, 4, 0.5, 1, 0, sp)),
    "fi_ori_fwd_py": lambda: V.FilterInterpolationLayer.apply(I, fl, ft),
    "fi_dkr_fwd": lambda: V.FilterInterpolationLayerDKR.apply(I, fl, ft, off),
    "fi_deforconv_fwd": lambda: V.FilterInterpolationLayerDeforConv.apply(I, fl, ft, off),
    "fi_nofilter_fwd": lambda: V.FilterInterpolationLayerNoFilterWithDeforConv.apply(I, fl, off),
    "fi_ori_bwd": lambda: _lib.call("vfidkr_filterinterpolation_backward_ori", ptr(I), ptr(fl), ptr(ft), ptr(g), ptr(gi1),
                                    ptr(gi2), ptr(gi3), B, C, H, W, 4, sp),
    "fi_dkr_bwd": lambda: _lib.call("vfidkr_filterinterpolation_backward_dkr", ptr(I), ptr(fl), ptr(ft), ptr(off), ptr(g),
                                    ptr(gi1), ptr(gi2), ptr(gi3), ptr(gi4), B, C, H, W, 4, sp),
    "interp_fwd": lambda: V.InterpolationChLayer.apply(I, fl),
    "interp_bwd": lambda: _lib.call("vfidkr_interpolation_backward", ptr(I), ptr(fl), ptr(g), ptr(gi1), ptr(gi2), B, C, H, W, 0, sp),
    "proj_fwd": lambda: V.FlowProjectionLayer.apply(fl, False),
    "dproj_fwd": lambda: V.DepthFlowProjectionLayer.apply(fl, dep, False),
}
if a.op == "pwc_warp":
    feat = torch.randn(B, 32, H // 4, W // 4, device=dev)
    flo4 = torch.nn.functional.avg_pool2d(fl, 4) / 4
    ops["pwc_warp"] = lambda: V.pwc_warp(feat, flo4)
if a.op == "pwc_warp_torch":
    feat = torch.randn(B, 32, H // 4, W // 4, device=dev)
    flo4 = torch.nn.functional.avg_pool2d(fl, 4) / 4
    def _ref():
        Bq, Cq, Hq, Wq = feat.shape
        xx = torch.arange(Wq, device=dev).view(1, 1, 1, Wq).expand(Bq, 1, Hq, Wq)
        yy = torch.arange(Hq, device=dev).view(1, 1, Hq, 1).expand(Bq, 1, Hq, Wq)
        vg = torch.cat((xx, yy), 1).float() + flo4
        vg = torch.stack([2.0 * vg[:, 0] / max(Wq - 1, 1) - 1.0, 2.0 * vg[:, 1] / max(Hq - 1, 1) - 1.0], 1).permute(0, 2, 3, 1)
        o = torch.nn.functional.grid_sample(feat, vg, align_corners=False)
        m = torch.nn.functional.grid_sample(torch.ones_like(o), vg, align_corners=False)
        m[m < 0.9999] = 0
        m[m > 0] = 1
        return o * m
    ops["pwc_warp_torch"] = _ref
if a.op in ("dproj_bwd", "proj_bwd"):
    cnt, po = torch.empty(B, 1, H, W, device=dev), torch.empty(B, 2, H, W, device=dev)
    g2, gd = torch.randn(B, 2, H, W, device=dev), torch.empty(B, 1, H, W, device=dev)
    if a.op == "dproj_bwd":
        _lib.call("vfidkr_depthflowprojection_forward", ptr(fl), ptr(dep), ptr(cnt), ptr(po), B, H, W, 0, sp)
        ops[a.op] = lambda: _lib.call("vfidkr_depthflowprojection_backward", ptr(fl), ptr(dep), ptr(cnt), ptr(po), ptr(g2),
                                      ptr(gi2), ptr(gd), B, H, W, sp)
    else:
        _lib.call("vfidkr_flowprojection_forward", ptr(fl), ptr(cnt), ptr(po), B, H, W, 0, sp)
        ops[a.op] = lambda: _lib.call("vfidkr_flowprojection_backward", ptr(fl), ptr(cnt), ptr(g2), ptr(gi2), B, H, W, sp)
if a.op == "fi_bench":
    # the FilterInterpolation calls of bench.py, on bench.py's own inputs (both directions)
    import bench
    d = bench.build_inputs(torch, dev, seed=1004)
    fi = V.FilterInterpolationModule()
    ops["fi_bench"] = lambda: (fi(d["frame0"], d["flow0"], d["filter0"]), fi(d["frame1"], d["flow1"], d["filter1"]))
if a.op == "sepconv":
    # SeparableConv forward + backward at the op table's shape (8 x 3 x 256 x 448, F = 51)
    Bs, Hs, Ws, Fs = 8, 256, 448, 51
    Is = torch.rand(Bs, 3, Hs, Ws, device=dev)
    vs = torch.rand(Bs, Fs, Hs - Fs + 1, Ws - Fs + 1, device=dev) / Fs
    hs = torch.rand_like(vs) / Fs
    os_, gs = torch.empty(Bs, 3, Hs - Fs + 1, Ws - Fs + 1, device=dev), torch.randn(Bs, 3, Hs - Fs + 1, Ws - Fs + 1, device=dev)
    g1s, g2s, g3s = torch.empty_like(Is), torch.empty_like(vs), torch.empty_like(hs)
    ops["sepconv"] = lambda: (_lib.call("vfidkr_separableconv_forward", ptr(Is), ptr(vs), ptr(hs), ptr(os_), Bs, 3, Hs, Ws, Fs, sp),
                              _lib.call("vfidkr_separableconv_backward", ptr(Is), ptr(vs), ptr(hs), ptr(gs), ptr(g1s), ptr(g2s), ptr(g3s),
                                        Bs, 3, Hs, Ws, Fs, sp))
if a.op == "corr_bwd":
    f1 = torch.randn(B, 32, H // 4, W // 4, device=dev)
    f2 = torch.randn_like(f1)
    go = torch.randn(B, 81, H // 4, W // 4, device=dev)
    ga, gb = torch.empty_like(f1), torch.empty_like(f2)
    ops["corr_bwd"] = lambda: _lib.call("vfidkr_correlation_backward", ptr(f1), ptr(f2), ptr(go), ptr(ga), ptr(gb), B, 32, H // 4, W // 4,
                                        4, 1, 4, 1, 1, 1, sp)
    fn = ops["corr_bwd"]
elif a.op.startswith("corr"):
    Cc, s = {"corr_l2": (32, 4), "corr_l3": (64, 8), "corr_l4": (96, 16), "corr_l5": (128, 32), "corr_l6": (196, 64)}[a.op]
    f1 = torch.randn(B, Cc, H // s, W // s, device=dev)
    f2 = torch.randn_like(f1)
    corr = V.Correlation(4, 1, 4, 1, 1, 1)
    fn = lambda: corr(f1, f2)
else:
    fn = ops[a.op]
with torch.no_grad():
    for _ in range(a.iters):
        fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
with torch.no_grad():
    e0.record()
    for _ in range(a.iters):
        fn()
    e1.record()
torch.cuda.synchronize()
print(f"{a.op}: {e0.elapsed_time(e1) / a.iters * 1e3:.1f} us per call (B={B} C={C} {H}x{W}, flow={a.flow})")
