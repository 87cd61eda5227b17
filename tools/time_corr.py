"""Correlation forward at the five PWC levels of a 1080p batch of 8: SIMT register-tile kernel vs tensor-core kernel."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import vfidkr_b200 as V

dev = torch.device("cuda", 0)
B, H, W = 8, 1152, 1984
corr = V.Correlation(4, 1, 4, 1, 1, 1)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3


with torch.no_grad():
    for C, s in ((196, 64), (128, 32), (96, 16), (64, 8), (32, 4)):
        a = torch.randn(B, C, H // s, W // s, device=dev)
        b = torch.randn_like(a)
        res = {}
        for path in ("simt", "tensor"):
            V.debug_force_correlation_path(path)
            res[path] = timeit(lambda: corr(a, b))
        pair = {}
        for path in ("simt", "tensor"):
            V.debug_force_correlation_path(path)
            pair[path] = timeit(lambda: corr.both_directions(a, b))
        V.debug_force_correlation_path("tensor")
        t = corr(a, b)
        V.debug_force_correlation_path("simt")
        sref = corr(a, b)
        err = ((t - sref).abs().max() / sref.abs().max()).item()
        gf = 2 * 81 * C * B * (H // s) * (W // s) / 1e9
        print(f"C={C:3d} {H // s}x{W // s}: SIMT {res['simt']:7.1f} us ({gf / res['simt'] * 1e3:6.1f} TFLOP/s useful), tensor {res['tensor']:7.1f} us "
              f"({gf / res['tensor'] * 1e3:6.1f} TFLOP/s useful); " f"max |tensor - SIMT| / max = {err:.2e}; both directions in one launch: SIMT {pair['simt']:.1f} us, "
              f"tensor {pair['tensor']:.1f} us")
V.debug_force_correlation_path(None)
