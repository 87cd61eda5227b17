"""Times the projection forwards at config 4 (8 x 1152 x 1984): three-kernel path vs fused pipeline, bench flow and up4 flow."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import bench
import vfidkr_b200 as V

dev = torch.device("cuda", 0)
B, H, W = 8, 1152, 1984
g = torch.Generator(device=dev); g.manual_seed(77)
fl = bench.scene_flow(torch, g, dev, B, H, W)
up4 = torch.nn.functional.interpolate((torch.randn(B, 2, H // 4, W // 4, device=dev) * 4).clamp_(-20, 20), scale_factor=4,
                                      mode="bilinear", align_corners=False).contiguous()
dep = torch.rand(B, 1, H, W, device=dev) * 0.9 + 0.1


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
    for path, kib in (("kernels", 0), ("chunks_pdl", 0), ("chunks_pdl", 700000), ("chunks_1px", 700000), ("chunks_pdl", 80000)):
        V.debug_force_projection_path(path)
        V.debug_projection_chunk_kib(kib)
        path = f"{path}/{kib // 1000}MB"
        for name, f in (("bench flow", fl), ("up4 flow", up4)):
            t_d = timeit(lambda: V.DepthFlowProjectionLayer.apply(f, dep, False))
            t_dn = timeit(lambda: V.DepthFlowProjectionLayer.apply(f, dep, True))
            t_f = timeit(lambda: V.FlowProjectionLayer.apply(f, False))
            print(f"{path:9s} {name:10s}: DepthFlowProjection fill {t_d:7.1f} us, no fill {t_dn:7.1f} us; FlowProjection fill {t_f:7.1f} us "
                  f"({24 * B * H * W / t_d / 1e3:.0f} GB/s = {24 * B * H * W / t_d / 1e3 / 6551.4 * 100:.1f} % of HBM peak)")
        for Bx, Hx, Wx in ((16, 256, 448), (2, 2176, 3904)):
            fx = bench.scene_flow(torch, g, dev, Bx, Hx, Wx)
            dx = torch.rand(Bx, 1, Hx, Wx, device=dev) * 0.9 + 0.1
            print(f"{path:9s} {Bx}x{Hx}x{Wx}: DepthFlowProjection fill {timeit(lambda: V.DepthFlowProjectionLayer.apply(fx, dx, False)):7.1f} us")
V.debug_force_projection_path(None)
V.debug_projection_chunk_kib(0)
