#!/usr/bin/env python
"""Times the reference's OWN CUDA kernels (oracle/_ref/*.so, unmodified sources built for sm_100a by
oracle/build_ref.py) on the same B200 and the same config-4 shapes as bench.py's operator table.
Measurement only -- this is the "recompiled reference" the B200-native kernels are meant to beat.

    python tools/time_ref.py [--out gpurun_out/ref_table.jsonl]
"""
import argparse
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
import torch

import build_ref

ap = argparse.ArgumentParser()
ap.add_argument("--out", default=None)
ap.add_argument("--B", type=int, default=8)
ap.add_argument("--H", type=int, default=1152)
ap.add_argument("--W", type=int, default=1984)
a = ap.parse_args()
dev = torch.device("cuda", 0)
B, H, W = a.B, a.H, a.W
px = B * H * W
peak = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"] if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
rows = []


def timeit(fn, iters=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def add(name, ms, bpp, n):
    gbs = bpp * n / (ms * 1e-3) / 1e9
    rows.append({"op": name, "impl": "reference kernels (unmodified, sm_100a build)", "ms": ms, "GBps": gbs, "frac_of_hbm": gbs / peak})
    print(f"[ref] {name:34s} {ms*1e3:10.1f} us  {gbs:8.1f} GB/s  {100*gbs/peak:5.1f} % of HBM peak", flush=True)


torch.manual_seed(0)
I = torch.rand(B, 3, H, W, device=dev)
fl = torch.nn.functional.interpolate((torch.randn(B, 2, H // 4, W // 4, device=dev) * 4).clamp_(-20, 20),
                                     scale_factor=4, mode="bilinear", align_corners=False).contiguous()
ft = torch.softmax(torch.randn(B, 16, H, W, device=dev), 1)
off = -(0.01 + 0.44 * torch.rand(B, 32, H, W, device=dev))   # negative offsets keep the reference in bounds
dep = torch.rand(B, 1, H, W, device=dev) * 0.9 + 0.1
g = torch.randn(B, 3, H, W, device=dev)
g2 = torch.randn(B, 2, H, W, device=dev)

fi = build_ref.load("filterinterpolation_cuda")
out = torch.zeros_like(I)


# the reference's Python layers zero-fill their outputs on every call (FilterInterpolationLayer.py:34,62-64);
# that cost is part of the reference path, so it is inside the timed lambda
def fi_fwd_ori():
    out.zero_()
    fi.FilterInterpolationLayer_gpu_forward_ori(I, fl, ft, out)


def fi_fwd_dkr():
    out.zero_()
    fi.FilterInterpolationLayer_gpu_forward(I, fl, ft, off, out)


gi1, gi2, gi3, gi4 = torch.zeros_like(I), torch.zeros_like(fl), torch.zeros_like(ft), torch.zeros_like(off)


def fi_bwd_ori():
    gi1.zero_(); gi2.zero_(); gi3.zero_()
    fi.FilterInterpolationLayer_gpu_backward_ori(I, fl, ft, g, gi1, gi2, gi3)


def fi_bwd_dkr():
    gi1.zero_(); gi2.zero_(); gi3.zero_(); gi4.zero_()
    fi.FilterInterpolationLayer_gpu_backward(I, fl, ft, off, g, gi1, gi2, gi3, gi4)


add("FI_ori_fwd_C3", timeit(fi_fwd_ori), 96, px)
add("FI_dkr_fwd_C3", timeit(fi_fwd_dkr), 224, px)
add("FI_ori_bwd_C3", timeit(fi_bwd_ori), 180, px)
add("FI_dkr_bwd_C3", timeit(fi_bwd_dkr), 436, px)

dp = build_ref.load("depthflowprojection_cuda")
fp = build_ref.load("flowprojection_cuda")
cnt, po = torch.zeros(B, 1, H, W, device=dev), torch.zeros(B, 2, H, W, device=dev)


def dproj(fill):
    cnt.zero_(); po.zero_()
    dp.DepthFlowProjectionLayer_gpu_forward(fl, dep, cnt, po, fill)


def fproj(fill):
    cnt.zero_(); po.zero_()
    fp.FlowProjectionLayer_gpu_forward(fl, cnt, po, fill)


add("DepthFlowProjection_fwd_fill", timeit(lambda: dproj(1)), 24, px)
add("FlowProjection_fwd_fill", timeit(lambda: fproj(1)), 20, px)
add("FlowProjection_fwd", timeit(lambda: fproj(0)), 20, px)
dproj(0)
gd = torch.zeros(B, 1, H, W, device=dev)


def dproj_bwd():
    gi2.zero_(); gd.zero_()
    dp.DepthFlowProjectionLayer_gpu_backward(fl, dep, cnt, po, g2, gi2, gd)


add("DepthFlowProjection_bwd", timeit(dproj_bwd), 44, px)

ip = build_ref.load("interpolation_cuda")


def interp_fwd():
    out.zero_()
    ip.InterpolationLayer_gpu_forward(I, fl, out)


add("Interpolation_fwd_C3", timeit(interp_fwd), 32, px)

del I, ft, off, g, gi1, gi3, gi4, out
torch.cuda.empty_cache()
cm = build_ref.load("correlation_cuda")
for C, s in [(196, 64), (128, 32), (96, 16), (64, 8), (32, 4)]:
    f1 = torch.randn(B, C, H // s, W // s, device=dev)
    f2 = torch.randn_like(f1)

    def corr():
        rb1, rb2, o = f1.new_empty(0), f2.new_empty(0), f1.new_empty(0)
        cm.forward(f1, f2, rb1, rb2, o, 4, 1, 4, 1, 1, 1)

    add(f"Correlation_fwd_C{C}_{H // s}x{W // s}", timeit(corr), 4 * (2 * C + 81), B * (H // s) * (W // s))

if a.out:
    with open(a.out, "w") as f:
        for r in rows:
            f.write(json.dumps(r) + "\n")
