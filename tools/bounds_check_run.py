#!/usr/bin/env python
"""Window-coherence / bounds check run (needs a -DVFIDKR_BOUNDS_CHECK build of the library, see csrc/common.cuh).

Every image value a kernel takes from a shared-memory window (the rolling window of the strip kernels, the region of the
many-channel kernel) is compared, bit for bit, with the global-memory value it stands for, and its index with the extent
of the window.  Prints the number of checks executed and failed per case; exits 1 on any failure."""
import ctypes
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch

import vfidkr_b200 as V
from vfidkr_b200 import _lib
from bench import scene_flow

dll = ctypes.CDLL(str(Path(_lib.__file__).parent / "libvfidkr_b200.so"))
if not hasattr(dll, "vfidkr_debug_bounds_counts"):
    sys.exit("this library was not built with -DVFIDKR_BOUNDS_CHECK")
dev = torch.device("cuda", 0)
buf = (ctypes.c_ulonglong * 2)()


def counts():
    assert dll.vfidkr_debug_bounds_counts(buf) == 0
    return int(buf[0]), int(buf[1])


def flows(B, H, W, gen):
    yield "scene", scene_flow(torch, gen, dev, B, H, W)
    yield "up4", torch.nn.functional.interpolate((torch.randn((B, 2, H // 4, W // 4), generator=gen, device=dev) * 4).clamp_(-20, 20),
                                                 scale_factor=4, mode="bilinear", align_corners=False).contiguous()
    yield "iid", (torch.randn((B, 2, H, W), generator=gen, device=dev) * 4).clamp_(-20, 20)
    yield "wild", (torch.randn((B, 2, H, W), generator=gen, device=dev) * 60)          # most tiles leave the window
    yy, xx = torch.meshgrid(torch.arange(H, device=dev, dtype=torch.float32), torch.arange(W, device=dev, dtype=torch.float32), indexing="ij")
    yield "shear", torch.stack([0.15 * (yy - H / 2), 0.1 * (xx - W / 2)], 0)[None].repeat(B, 1, 1, 1).contiguous()   # re-bases


total = [0, 0]
failed = False
shapes = [(8, 3, 1152, 1984), (2, 4, 1152, 1984), (4, 3, 256, 448), (2, 1, 132, 200), (3, 2, 64, 176), (1, 3, 2176, 3904)]
with torch.no_grad():
    for (B, C, H, W) in shapes:
        gen = torch.Generator(device=dev)
        gen.manual_seed(B * 1000 + W)
        I = torch.rand((B, C, H, W), generator=gen, device=dev)
        ft = torch.softmax(torch.randn((B, 16, H, W), generator=gen, device=dev), 1)
        off = (torch.rand((B, 32, H, W), generator=gen, device=dev) - 0.5) * 1.2          # some taps leave the in-contract domain
        for name, fl in flows(B, H, W, gen):
            for op, fn in (("ori", lambda: V.FilterInterpolationLayer.apply(I, fl, ft)),
                           ("ori blend", lambda: V.filter_interpolate_blend(I, I, fl, fl, ft, ft)),
                           ("dkr", lambda: V.FilterInterpolationLayerDKR.apply(I, fl, ft, off)),
                           ("deforconv", lambda: V.FilterInterpolationLayerDeforConv.apply(I, fl, ft, off)),
                           ("nofilter", lambda: V.FilterInterpolationLayerNoFilterWithDeforConv.apply(I, fl, off))):
                if (B, H) == (8, 1152) and op not in ("ori", "dkr") and name not in ("scene",):
                    continue                                                              # keep the full-size part short
                before = counts()
                for _ in range(2):                                                        # back to back: stale scratch / windows
                    fn()
                after = counts()
                n, bad = after[0] - before[0], after[1] - before[1]
                total[0] += n
                total[1] += bad
                failed |= bad != 0
                print(f"{op:10s} {B}x{C}x{H}x{W} flow={name:6s}: {n:>14,d} window checks, {bad} failed")
    # many-channel kernel (region per tile)
    for (B, C, H, W) in ((2, 196, 288, 496), (1, 64, 1152, 1984), (1, 9, 72, 132)):
        gen = torch.Generator(device=dev)
        gen.manual_seed(C)
        I = torch.rand((B, C, H, W), generator=gen, device=dev)
        ft = torch.softmax(torch.randn((B, 16, H, W), generator=gen, device=dev), 1)
        for name, fl in flows(B, H, W, gen):
            before = counts()
            V.FilterInterpolationLayer.apply(I, fl, ft)
            after = counts()
            n, bad = after[0] - before[0], after[1] - before[1]
            total[0] += n
            total[1] += bad
            failed |= bad != 0
            print(f"{'ori C>4':10s} {B}x{C}x{H}x{W} flow={name:6s}: {n:>14,d} window checks, {bad} failed")
print(f"TOTAL: {total[0]:,d} window checks executed, {total[1]} failed")
sys.exit(1 if failed else 0)
