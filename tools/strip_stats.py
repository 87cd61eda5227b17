#!/usr/bin/env python
"""Pipeline statistics of the FI strip kernel (needs a debug build: VFIDKR_NVCC_EXTRA=-DVFIDKR_STRIP_STATS)."""
import ctypes, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
import vfidkr_b200 as V
from vfidkr_b200 import _lib
from bench import scene_flow

dll = ctypes.CDLL(str(Path(_lib.__file__).parent / "libvfidkr_b200.so"))
dev = torch.device("cuda", 0)
B, C, H, W = 8, 3, 1152, 1984
torch.manual_seed(0)
I = torch.rand(B, C, H, W, device=dev)
ft = torch.softmax(torch.randn(B, 16, H, W, device=dev), 1)
gen = torch.Generator(device=dev); gen.manual_seed(77)
flows = {"zero": torch.zeros(B, 2, H, W, device=dev), "scene": scene_flow(torch, gen, dev, B, H, W)}
fi = V.FilterInterpolationModule()
buf = (ctypes.c_ulonglong * 16)()
for name, fl in flows.items():
    for _ in range(2):
        fi(I, fl, ft)
    dll.vfidkr_debug_strip_stats(buf)
    n = 5
    for _ in range(n):
        fi(I, fl, ft)
    dll.vfidkr_debug_strip_stats(buf)
    s = [x / n for x in buf]
    ctas = 148
    clk = 1.965e3   # cycles per us
    print(f"== {name}: per CTA averages (us)")
    print(f"  producer: total {s[6]/ctas/clk:.1f}  box-wait {s[0]/ctas/clk:.1f}  rebase-drain {s[1]/ctas/clk:.1f}  slot-reuse-wait {s[2]/ctas/clk:.1f}"
          f"  rebases/CTA {s[3]/ctas:.1f}  global tiles/CTA {s[4]/ctas:.1f}  tiles/CTA {s[5]/ctas:.1f}")
    print(f"  compute warp 0: total {s[11]/ctas/clk:.1f}  image-full wait {s[8]/ctas/clk:.1f}  filter-full wait {s[9]/ctas/clk:.1f}  tiles/CTA {s[10]/ctas:.1f}")
